#!/usr/bin/env python
"""The reference's batching idiom, `examples/batch_rendering.py:83-95`, executed by the UNMODIFIED reference:

    jax.vmap(lambda model, buffer: Renderer.render(model=model, light=light, camera=camera, buffers=buffer,
                                                   shadow_param=shadow_param))(batch_models(merged_models), buffers)

on three poses of a small scene (ground box, cube, capsule) at 24x18 with the shadow pass, through the NumPy stand-in for
jax -> `tests/golden/reference_run_vmap.npz` (inputs per pose + the batched z-buffers and canvases).  The package's native
leading batch axis (and `torch.func.vmap` of the same lambda) must reproduce it.   ~3 minutes.

  python tools/gen_reference_fixtures_vmap.py [/root/reference]
"""
from __future__ import annotations

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_reference_fixtures as G  # noqa: E402  (puts the stand-in and the reference on sys.path)
import numpy as np  # noqa: E402
import jax  # noqa: E402
import jax.numpy as jnp  # noqa: E402
import renderer as R  # noqa: E402

J = G.J
W, H, POSES = 24, 18, 3


def objects(k: int, rng):
    """Pose k: the cube and the capsule move and turn, the ground stays."""
    ground = R.create_cube(half_extents=jnp.array((3.0, 3.0, 0.05)), texture_scaling=jnp.array(4.0),
                           diffuse_map=J(rng["ground"]), specular_map=jnp.ones((6, 5)) * 2.0)
    cube = R.create_cube(half_extents=jnp.array((0.5, 0.4, 0.3)), texture_scaling=jnp.array(1.0),
                         diffuse_map=J(rng["cube"]), specular_map=jnp.ones((2, 2)) * 3.0)
    cap = R.create_capsule(radius=jnp.array(0.3), half_height=jnp.array(0.4), up_axis=R.UpAxis.Z,
                           diffuse_map=J(rng["cap"]), specular_map=jnp.ones((1, 1)) * 2.0)
    return [
        R.ModelObject(model=ground),
        R.ModelObject(model=cube).replace_with_position(jnp.array((0.2 + 0.4 * k, -0.3, 0.8 + 0.1 * k)))
         .replace_with_orientation(R.quaternion(jnp.array((0.3, 0.4, 0.5)), jnp.array(40.0 + 25.0 * k))),
        R.ModelObject(model=cap, local_scaling=jnp.array((1.0, 1.2, 0.9)))
         .replace_with_position(jnp.array((-0.9, 0.6 - 0.5 * k, 0.9))),
    ]


def main():
    t0 = time.time()
    g = np.random.default_rng(17)
    maps = {"ground": g.random((6, 5, 3), dtype=np.float32), "cube": g.random((2, 2, 3), dtype=np.float32),
            "cap": g.random((1, 1, 3), dtype=np.float32)}
    poses = [objects(k, maps) for k in range(POSES)]
    cp = R.CameraParameters(viewWidth=W, viewHeight=H, position=jnp.array((3.0, -3.5, 2.5)),
                            target=jnp.array((0.0, 0.0, 0.5)), up=jnp.array((0.0, 0.0, 1.0)), hfov=58.0, vfov=58.0 * H / W)
    light = R.LightParameters()
    sp = R.ShadowParameters(centre=jnp.array((0.0, 0.0, 0.5)))
    merged = [R.merge_objects(objs) for objs in poses]
    buffers = R.Renderer.create_buffers(W, H, POSES)
    camera = R.Renderer.create_camera_from_parameters(cp)
    out = jax.vmap(lambda model, buffer: R.Renderer.render(model=model, light=light, camera=camera, buffers=buffer,
                                                           shadow_param=sp))(R.batch_models(merged), buffers)
    zbuffer, (canvas,) = out
    OUT = {"W": W, "H": H, "poses": POSES, "zbuffer": np.asarray(zbuffer), "canvas": np.asarray(canvas),
           "cam_position": np.asarray(cp.position), "cam_target": np.asarray(cp.target), "cam_up": np.asarray(cp.up),
           "hfov": np.float32(cp.hfov), "vfov": np.float32(cp.vfov), "shadow_centre": np.asarray(sp.centre)}
    for k, objs in enumerate(poses):
        for i, o in enumerate(objs):
            for f in ("verts", "norms", "uvs", "faces", "faces_norm", "faces_uv", "diffuse_map", "specular_map"):
                OUT[f"pose{k}/obj{i}/{f}"] = np.asarray(getattr(o.model, f))
            OUT[f"pose{k}/obj{i}/local_scaling"] = np.asarray(o.local_scaling)
            OUT[f"pose{k}/obj{i}/transform"] = np.asarray(o.transform)
    dst = os.path.join(G.ROOT, "tests", "golden", "reference_run_vmap.npz")
    np.savez_compressed(dst, **OUT)
    print(f"wrote {dst}: canvas {OUT['canvas'].shape}, covered {(OUT['zbuffer'] != 1).mean():.2f}, {time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
