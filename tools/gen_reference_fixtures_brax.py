#!/usr/bin/env python
"""Frame 0 of the reference's pre-generated Brax ant scene (`tests/golden/brax_ant_frames.npz`, 18 objects, 3276
triangles) rendered by the UNMODIFIED reference's `Renderer.get_camera_image` (with the shadow pass) through the NumPy
stand-in for jax: at 20x20 -> `tests/golden/reference_run_brax.npz` (10-15 minutes, vmap is a Python loop), and at the
size BASELINE.json's configs[1] names, 84x84 -> `tests/golden/reference_run_brax84.npz` (46 M fragment evaluations:
about 40 minutes with the stand-in's outermost vmap split over 8 forked workers, JAX_SHIM_PROCS=8).

  python tools/gen_reference_fixtures_brax.py [/root/reference]
  JAX_SHIM_PROCS=8 python tools/gen_reference_fixtures_brax.py /root/reference 84
"""
from __future__ import annotations

import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
SIZE = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sys.path.insert(0, os.path.join(ROOT, "tools", "jax_numpy_shim"))
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import jax.numpy as jnp  # noqa: E402
import renderer as R  # noqa: E402

W = H = SIZE
FRAME = 0


def main():
    d = np.load(os.path.join(ROOT, "tests", "golden", "brax_ant_frames.npz"))
    J = lambda x: jnp.asarray(np.asarray(x))  # noqa: E731
    objs = []
    for i in range(int(d["n_objects"])):
        g = lambda k: d[f"o{i}_{k}"]  # noqa: E731
        m = R.Model(verts=J(g("verts")), norms=J(g("norms")), uvs=J(g("uvs")), faces=J(g("faces")),
                    faces_norm=J(g("faces_norm")), faces_uv=J(g("faces_uv")), diffuse_map=J(g("diffuse_map")),
                    specular_map=J(g("specular_map")))
        objs.append(R.ModelObject(model=m, local_scaling=J(g("local_scaling")[FRAME]), transform=J(g("transform")[FRAME]),
                                  double_sided=J(g("double_sided")[FRAME])))
    c = {k: np.asarray(d[f"cam_{k}"])[FRAME] for k in R.CameraParameters._fields}
    cp = R.CameraParameters(viewWidth=W, viewHeight=H, viewDepth=float(c["viewDepth"]), near=float(c["near"]),
                            far=float(c["far"]), hfov=float(c["hfov"]), vfov=float(c["hfov"]) * H / W,
                            position=J(c["position"]), target=J(c["target"]), up=J(c["up"]))
    light = R.LightParameters(direction=jnp.array((0.57735, -0.57735, 0.57735)), ambient=jnp.array((0.8, 0.8, 0.8)),
                              diffuse=jnp.array((0.8, 0.8, 0.8)), specular=jnp.array((0.6, 0.6, 0.6)))
    sp = R.ShadowParameters(centre=J(c["target"]))
    t = time.time()
    img = R.Renderer.get_camera_image(objs, light, cp, W, H, shadow_param=sp)
    print(f"rendered in {time.time() - t:.0f}s; background fraction {(np.asarray(img) == 1).all(-1).mean():.2f}")
    dst = os.path.join(ROOT, "tests", "golden", "reference_run_brax.npz" if SIZE == 20 else f"reference_run_brax{SIZE}.npz")
    np.savez_compressed(dst, canvas=np.asarray(img), W=W, H=H, frame=FRAME, vfov=np.float32(float(c["hfov"]) * H / W),
                        light_direction=np.asarray(light.direction), ambient=np.asarray(light.ambient),
                        diffuse=np.asarray(light.diffuse), specular=np.asarray(light.specular))
    print("wrote", dst)


if __name__ == "__main__":
    main()
