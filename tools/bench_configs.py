#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs 3, 4, 5 (parity-test shapes) timed on one GPU.

Not the contract bench (that is ../bench.py on configs[1]); this script reports device-resident
images/s for the other shapes so that DESIGN.md can quote them and ncu can be pointed at the
shading / backward kernels:

  python tools/bench_configs.py --cfg 3 [--batch N] [--steps K]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import jaxrenderer_b200 as jr  # noqa: E402
from jaxrenderer_b200 import _native, synthetic  # noqa: E402
from jaxrenderer_b200.shaders import GouraudTextureExtraInput, GouraudTextureShader  # noqa: E402


def timeit(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _native.launch_count()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (_native.launch_count() - l0) // steps


def merged_model(sc, n_caps, dev, atlas_tex=100):
    return synthetic.merged_model_from_batch(sc, n_caps, dev, atlas_tex)


NOTEBOOK_LIGHT = jr.LightParameters(direction=(0.57735, -0.57735, 0.57735), ambient=(0.8,) * 3,
                                    diffuse=(0.8,) * 3, specular=(0.6,) * 3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, required=True, choices=(3, 4, 5))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--size", type=str, default="", help="cfg 4/5: override WxH, e.g. 84x84")
    ap.add_argument("--caps", type=int, default=0, help="cfg 4/5: override the number of capsules")
    ap.add_argument("--vertex-grads", action="store_true",
                    help="cfg 5: also request gradients w.r.t. vertex positions and normals")
    ap.add_argument("--profile", action="store_true", help="cfg 4/5: torch-profiler kernel table on stderr")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    out = {"cfg": args.cfg}
    if args.cfg == 3:
        W = H = 32
        B = args.batch or 16384
        n_caps = 10  # T = 1932 (BASELINE: mixed envs, T in 576..3852)
        sc = synthetic.brax_like_batch(B, n_capsules=n_caps, with_attributes=True)
        cam = synthetic.brax_cameras(sc["eye"], sc["target"], W, H)
        cam = type(cam)(*[t.to(dev) for t in cam])
        tex = synthetic.checker_texture().to(dev)
        extra = GouraudTextureExtraInput(sc["position"].to(dev), sc["normal"].to(dev), (sc["uv"] * 100).to(dev),
                                         jr.LightSource(torch.tensor((0.57735, -0.57735, 0.57735), device=dev),
                                                        torch.ones(3, device=dev)), tex)
        faces = sc["faces"].to(dev)
        bufs = jr.Renderer.create_buffers(W, H, batch=B, device=dev)

        def step():
            jr.render(cam, GouraudTextureShader, bufs, faces, extra, inplace=True)
        ms, launches = timeit(step, args.steps)
        out.update(shader="gouraud_texture", W=W, H=H, B=B, T=synthetic.scene_sizes(n_caps)[1], ms_per_step=ms,
                   images_per_s=B / ms * 1e3, launches=launches)
    else:
        if args.cfg == 4:
            W, H, n_caps = 960, 540, 104  # T = 19980
            B = args.batch or 64
        else:
            W, H, n_caps = 480, 270, 17   # T = 3276
            B = args.batch or 128
        if args.size:
            W, H = (int(v) for v in args.size.split("x"))
        n_caps = args.caps or n_caps
        sc = synthetic.brax_like_batch(B, n_capsules=n_caps, with_attributes=True)
        cam = synthetic.brax_cameras(sc["eye"], sc["target"], W, H)
        cam = type(cam)(*[t.to(dev) for t in cam])
        model = merged_model(sc, n_caps, dev)
        sp = jr.ShadowParameters(centre=sc["target"].to(dev))
        bufs = jr.Renderer.create_buffers(W, H, batch=B, device=dev)
        if args.cfg == 4:
            def step():
                jr.Renderer.render(model, NOTEBOOK_LIGHT, cam, bufs, shadow_param=sp, inplace=True)
            ms, launches = timeit(step, args.steps)
            if args.profile:
                from torch.profiler import ProfilerActivity, profile
                with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
                    for _ in range(3):
                        step()
                    torch.cuda.synchronize()
                print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14), file=sys.stderr)
            out.update(shader="phong_reflection_shadow", W=W, H=H, B=B, T=synthetic.scene_sizes(n_caps)[1],
                       ms_per_step=ms, images_per_s=B / ms * 1e3, launches=launches)
        else:
            target = torch.rand(B, W, H, 3, device=dev)
            atlas = model.diffuse_map.clone().requires_grad_(True)
            ldir = torch.tensor(NOTEBOOK_LIGHT.direction, device=dev, requires_grad=True)
            amb = torch.tensor(NOTEBOOK_LIGHT.ambient, device=dev, requires_grad=True)
            w2c = cam.world_to_clip.clone().requires_grad_(True)
            verts = model.verts.clone().requires_grad_(args.vertex_grads)
            norms = model.norms.clone().requires_grad_(args.vertex_grads)

            def fwd():
                m = model._replace(diffuse_map=atlas, verts=verts, norms=norms)
                light = NOTEBOOK_LIGHT._replace(direction=ldir, ambient=amb)
                c = cam._replace(world_to_clip=w2c)
                b0 = jr.Renderer.create_buffers(W, H, batch=B, device=dev)
                return jr.Renderer.render(m, light, c, b0, shadow_param=sp)

            def step_fwd():
                with torch.no_grad():
                    fwd()

            def step():
                for p in (atlas, ldir, amb, w2c, verts, norms):
                    p.grad = None
                o = fwd()
                loss = ((o.targets[0] - target) ** 2).mean()
                loss.backward()
            ms_f, l_f = timeit(step_fwd, args.steps)
            ms, launches = timeit(step, args.steps)
            if args.profile:
                from torch.profiler import ProfilerActivity, profile
                with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
                    for _ in range(3):
                        step()
                    torch.cuda.synchronize()
                print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30), file=sys.stderr)
            out.update(shader="phong_reflection_shadow fwd+bwd (grads: light, world_to_clip, shared diffuse atlas)",
                       W=W, H=H, B=B, T=synthetic.scene_sizes(n_caps)[1], ms_per_step=ms, ms_forward_only=ms_f,
                       images_per_s=B / ms * 1e3, launches=launches, launches_forward=l_f,
                       grad_norms={"atlas": float(atlas.grad.norm()), "light_dir": float(ldir.grad.norm()),
                                   "ambient": float(amb.grad.norm()), "w2c": float(w2c.grad.norm())})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
