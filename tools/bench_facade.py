#!/usr/bin/env python
"""Secondary measurement: the whole ``Renderer.get_camera_image`` facade (merge_objects -> camera ->
shadow pass -> Phong pass -> uint8-ready canvas) on the real Brax ant scene of the reference's
``test_resources/pre-gen-brax`` fixture (tests/golden/brax_ant_frames.npz, 4 frames tiled to ``--batch``
environments), inputs resident on the device.

  python tools/bench_facade.py [--batch 1024] [--width 84 --height 84] [--steps 20] [--profile]

Prints one JSON line: total ms / step plus the split merge / camera / render, and the native launch count.
``--profile`` adds the torch-profiler kernel table (CPU vs GPU time of the step) on stderr.
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
from jaxrenderer_b200 import _native  # noqa: E402
from tests.helpers import load_brax_fixture  # noqa: E402


def tile(t: torch.Tensor, B: int, dev) -> torch.Tensor:
    reps = (B + t.shape[0] - 1) // t.shape[0]
    return t.repeat((reps,) + (1,) * (t.ndim - 1))[:B].contiguous().to(dev)


def timeit(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    host_ms = (time.perf_counter() - t0) * 1e3 / steps   # time the host needs to enqueue one step
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, host_ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--width", type=int, default=84)
    ap.add_argument("--height", type=int, default=84)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--merged", action="store_true", help="also time merge + materialise and the render of merged arrays")
    ap.add_argument("--graph", action="store_true",
                    help="also capture the whole facade call in a CUDA graph and time its replay")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, W, H = args.batch, args.width, args.height
    objs, cam = load_brax_fixture()
    n_frames = objs[0].transform.shape[0]
    objs = [jr.ModelObject(model=type(o.model)(*[t.to(dev) for t in o.model]),
                           local_scaling=tile(o.local_scaling, B, dev), transform=tile(o.transform, B, dev),
                           double_sided=tile(o.double_sided, B, dev)) for o in objs]
    cam = type(cam)(*[tile(v, B, dev) if isinstance(v, torch.Tensor) and v.ndim >= 1 and v.shape[0] == n_frames
                      else v for v in cam])
    cam = cam._replace(viewWidth=W, viewHeight=H)
    light = jr.LightParameters()
    sp = jr.ShadowParameters(centre=cam.target)

    def full():
        return jr.Renderer.get_camera_image(objs, light, cam, W, H, shadow_param=sp)

    def merge():
        return jr.merge_objects(objs)

    def camera():
        return jr.Renderer.create_camera_from_parameters(cam)

    model, camera_obj = merge(), camera()
    bufs = jr.Renderer.create_buffers(W, H, batch=B, device=dev)

    def render():
        return jr.Renderer.render(model, light, camera_obj, bufs, shadow_param=sp)

    def materialise():
        m = jr.merge_objects(objs)
        return m._replace(verts=m.verts.materialise(), norms=m.norms.materialise())

    model_merged = materialise()

    def render_merged():
        return jr.Renderer.render(model_merged, light, camera_obj, bufs, shadow_param=sp)

    out = {"workload": f"get_camera_image brax-ant fixture B={B} {W}x{H}", "T": int(model.faces.shape[-2])}
    if args.merged:   # A/B: world-space arrays written by the merge kernels, render kernels without instancing
        for name, fn in (("materialise", materialise), ("render_merged", render_merged)):
            out[f"ms_{name}"], out[f"ms_{name}_host"] = (round(v, 4) for v in timeit(fn, args.steps))
    l0 = _native.launch_count()
    full()
    out["launches"] = _native.launch_count() - l0
    for name, fn in (("total", full), ("merge", merge), ("camera", camera), ("render", render)):
        out[f"ms_{name}"], out[f"ms_{name}_host"] = (round(v, 4) for v in timeit(fn, args.steps))
    out["images_per_s"] = B / out["ms_total"] * 1e3
    if args.graph:
        # launch-bound at small batches: the call is allocation- and sync-free after warm-up, so it captures
        # into one CUDA graph (inputs are read from the same device tensors at every replay)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            full()
        torch.cuda.current_stream(dev).wait_stream(side)
        with torch.cuda.graph(graph):
            img = full()
        ref = full()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(img, ref), "graph replay differs from the eager call"
        out["ms_total_graph"], out["ms_total_graph_host"] = (round(v, 4) for v in timeit(graph.replay, args.steps))
        out["images_per_s_graph"] = B / out["ms_total_graph"] * 1e3
    print(json.dumps(out))
    if args.profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                full()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25), file=sys.stderr)
        print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25), file=sys.stderr)


if __name__ == "__main__":
    main()
