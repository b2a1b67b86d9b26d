#!/usr/bin/env python
"""Gradient references from the reference's own forward code: central finite differences in FLOAT64.

The NumPy stand-in for jax has no autodiff, but it can run the unmodified reference in double precision
(`JAX_SHIM_X64=1`, the analogue of `jax_enable_x64`).  For scene `soup0` of `tests/golden/reference_run.npz` and the
loss `sum(wz * zbuffer) + sum(wc * canvas)` (fixed random weights), this script perturbs single entries of the
differentiable inputs by +-h (h = 1e-6) and stores `(loss(+h) - loss(-h)) / 2h`; entries whose second difference
is not O(h^2) -- a visibility, texel or shadow change inside the step -- are dropped.  Output:
`tests/golden/reference_grad.npz`.  About 20 minutes.

  python tools/gen_reference_grad_fixtures.py          # depth, gouraud, phong_reflection_shadow
  python tools/gen_reference_grad_fixtures.py more     # gouraud_texture, phong, phong_darboux, phong_reflection
"""
from __future__ import annotations

import os
import sys
import time

os.environ["JAX_SHIM_X64"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402

import gen_reference_fixtures as G  # noqa: E402  (puts the stand-in and the reference on sys.path)

R, jnp, J = G.R, G.jnp, G.J

D = np.load(os.path.join(ROOT, "tests", "golden", "reference_run.npz"))
P = "soup0"
W, H = int(D[P + "/W"]), int(D[P + "/H"])
rng = np.random.default_rng(123)
WZ = rng.random((W, H)) + 0.5
WC = rng.random((W, H, 3)) + 0.5


def inputs():
    keys = ["uv_texel", "texture", "normal_map", "position", "normal", "colour", "uv01", "light_direction", "light_colour", "world_to_clip", "viewport",
            "world_to_eye_norm", "view", "atlas", "specular_map", "light_dir_eye", "ambient", "diffuse", "specular",
            "shadow_strength", "shadow_map", "shadow_world_to_clip", "shadow_viewport", "texture_shape",
            "texture_index", "faces"]
    return {k: np.array(D[f"{P}/{k}"], dtype=np.float64 if D[f"{P}/{k}"].dtype.kind == "f" else None) for k in keys}


def camera(v):
    eye4 = jnp.identity(4)
    return R.Camera(view=J(v["view"]), projection=eye4, viewport=J(v["viewport"]), world_to_clip=J(v["world_to_clip"]),
                    world_to_eye_norm=J(v["world_to_eye_norm"]), world_to_screen=eye4, view_inv=eye4, screen_to_world=eye4)


def loss(shader_name, v):
    cam = camera(v)
    faces = J(v["faces"])
    z0, c0 = jnp.ones((W, H)), jnp.full((W, H, 3), 0.25)
    light = R.LightSource(direction=J(v["light_direction"]), colour=J(v["light_colour"]))
    if shader_name == "depth":
        out = R.render(cam, G.DepthShader, R.Buffers(zbuffer=z0, targets=()), faces,
                       G.DepthExtraInput(position=J(v["position"])))
        return float((np.asarray(out.zbuffer) * WZ).sum())
    n_tri = v["faces"].shape[0]
    if shader_name == "gouraud":
        extra = G.GouraudExtraInput(position=J(v["position"]), colour=J(v["colour"]), normal=J(v["normal"]), light=light)
        shader = G.GouraudShader
    elif shader_name == "gouraud_texture":
        extra = G.GouraudTextureExtraInput(position=J(v["position"]), normal=J(v["normal"]), uv=J(v["uv_texel"]),
                                           light=light, texture=J(v["texture"]))
        shader = G.GouraudTextureShader
    elif shader_name == "phong":
        extra = G.PhongTextureExtraInput(position=J(v["position"]), normal=J(v["normal"]), uv=J(v["uv_texel"]),
                                         light=light, texture=J(v["texture"]))
        shader = G.PhongTextureShader
    elif shader_name == "phong_darboux":
        extra = G.PhongTextureDarbouxExtraInput(
            position=J(v["position"]), normal=J(v["normal"]), uv=J(v["uv_texel"]), light=light, texture=J(v["texture"]),
            normal_map=J(v["normal_map"]), id_to_face=J(np.repeat(np.arange(n_tri, dtype=np.int32), 3)), faces_indices=faces)
        shader = G.PhongTextureDarbouxShader
    elif shader_name == "phong_reflection":
        extra = G.PhongReflectionTextureExtraInput(
            position=J(v["position"]), normal=J(v["normal"]), uv=J(v["uv01"]), light=light,
            light_dir_eye=J(v["light_dir_eye"]), texture_shape=J(v["texture_shape"]), texture_index=J(v["texture_index"]),
            texture_offset=J(np.int32(D[P + "/texture_offset"])), texture=J(v["atlas"]), specular_map=J(v["specular_map"]),
            ambient=J(v["ambient"]), diffuse=J(v["diffuse"]), specular=J(v["specular"]))
        shader = G.PhongReflectionTextureShader
    else:
        scam = R.Camera(view=jnp.identity(4), projection=jnp.identity(4), viewport=J(v["shadow_viewport"]),
                        world_to_clip=J(v["shadow_world_to_clip"]), world_to_eye_norm=jnp.identity(4),
                        world_to_screen=jnp.identity(4), view_inv=jnp.identity(4), screen_to_world=jnp.identity(4))
        shadow = G.Shadow(shadow_map=J(v["shadow_map"]), strength=J(v["shadow_strength"]), camera=scam)
        extra = G.PhongReflectionShadowTextureExtraInput(
            position=J(v["position"]), normal=J(v["normal"]), uv=J(v["uv01"]), light=light,
            light_dir_eye=J(v["light_dir_eye"]), texture_shape=J(v["texture_shape"]), texture_index=J(v["texture_index"]),
            texture_offset=J(np.int32(D[P + "/texture_offset"])), texture=J(v["atlas"]), specular_map=J(v["specular_map"]),
            shadow=shadow, camera=cam, ambient=J(v["ambient"]), diffuse=J(v["diffuse"]), specular=J(v["specular"]))
        shader = G.PhongReflectionShadowTextureShader
    out = R.render(cam, shader, R.Buffers(zbuffer=z0, targets=(c0,)), faces, extra)
    return float((np.asarray(out.zbuffer) * WZ).sum() + (np.asarray(out.targets[0]) * WC).sum())


def visible_triangles(base, top=8):
    """Triangles ranked by the number of pixel centres they win (plain float64 rasterisation of the fixture's
    geometry: only used to pick WHICH vertices to perturb, never as a reference value)."""
    pos = np.concatenate([base["position"], np.ones((base["position"].shape[0], 1))], axis=1)
    clip = pos @ base["world_to_clip"].T
    ndc = clip[:, :3] / clip[:, 3:4]
    scr = np.concatenate([ndc, np.ones((ndc.shape[0], 1))], axis=1) @ base["viewport"].T
    faces = base["faces"].astype(int)
    best = np.full((W, H), np.inf)
    owner = np.full((W, H), -1)
    for t, (a, b, c) in enumerate(faces):
        pa, pb, pc = scr[a], scr[b], scr[c]
        area = (pb[0] - pa[0]) * (pc[1] - pa[1]) - (pb[1] - pa[1]) * (pc[0] - pa[0])
        if abs(area) < 1e-12 or min(clip[a, 3], clip[b, 3], clip[c, 3]) <= 0:
            continue
        for x in range(W):
            for y in range(H):
                w0 = ((pb[0] - x) * (pc[1] - y) - (pb[1] - y) * (pc[0] - x)) / area
                w1 = ((pc[0] - x) * (pa[1] - y) - (pc[1] - y) * (pa[0] - x)) / area
                w2 = 1.0 - w0 - w1
                if w0 < 0 or w1 < 0 or w2 < 0:
                    continue
                z = w0 * pa[2] + w1 * pb[2] + w2 * pc[2]
                if z < best[x, y]:
                    best[x, y], owner[x, y] = z, t
    counts = np.bincount(owner[owner >= 0], minlength=len(faces))
    return [int(t) for t in np.argsort(-counts)[:top] if counts[t] > 0]


def main():
    t0 = time.time()
    base = inputs()
    plans = {
        "depth": [("world_to_clip", (0, 0)), ("world_to_clip", (1, 3)), ("world_to_clip", (2, 2)),
                  ("world_to_clip", (3, 2)), ("viewport", (2, 2)), ("viewport", (2, 3))],
        "gouraud": [("light_colour", (0,)), ("light_colour", (2,)), ("light_direction", (0,)),
                    ("light_direction", (1,)), ("light_direction", (2,)), ("world_to_clip", (0, 0)),
                    ("world_to_clip", (1, 3)), ("world_to_clip", (3, 1)), ("viewport", (0, 0)), ("viewport", (1, 3))],
        "phong_reflection_shadow": [("light_colour", (1,)), ("light_dir_eye", (0,)), ("light_dir_eye", (2,)),
                                    ("ambient", (0,)), ("diffuse", (1,)), ("specular", (2,)), ("shadow_strength", (1,)),
                                    ("world_to_clip", (0, 1)), ("world_to_clip", (3, 3)),
                                    ("world_to_eye_norm", (0, 0)), ("world_to_eye_norm", (1, 2)),
                                    ("world_to_eye_norm", (2, 1)), ("viewport", (0, 3))],
    }
    second = len(sys.argv) > 1 and sys.argv[-1] == "more"
    if second:   # the four remaining shaders -> tests/golden/reference_grad_more.npz
        common = [("light_colour", (0,)), ("light_direction", (1,)), ("light_direction", (2,)),
                  ("world_to_clip", (0, 0)), ("world_to_clip", (3, 1)), ("viewport", (1, 1))]
        plans = {
            "gouraud_texture": list(common),
            "phong": common + [("world_to_eye_norm", (0, 1)), ("world_to_eye_norm", (2, 2))],
            "phong_darboux": common + [("world_to_eye_norm", (1, 1)), ("world_to_eye_norm", (0, 2))],
            "phong_reflection": [("light_colour", (2,)), ("light_dir_eye", (1,)), ("ambient", (2,)), ("diffuse", (0,)),
                                 ("specular", (1,)), ("world_to_clip", (1, 1)), ("world_to_eye_norm", (2, 0))],
        }
    # vertex attributes of the triangles that cover most pixels (one coordinate per vertex, cycled)
    tris = visible_triangles(base)
    print("  most visible triangles:", tris, flush=True)
    for j, t in enumerate(tris):
        for k in range(3):
            vtx = int(base["faces"][t][k])
            c = (j + k) % 3
            if second:
                if j >= 4:
                    continue
                for sh in ("gouraud_texture", "phong", "phong_darboux", "phong_reflection"):
                    plans[sh].append(("position", (vtx, (c + 1) % 3)))
                    plans[sh].append(("normal", (vtx, c)))
                plans["phong_darboux"].append(("uv_texel", (vtx, k % 2)))
                continue
            plans["depth"].append(("position", (vtx, c)))
            plans["gouraud"].append(("position", (vtx, (c + 1) % 3)))
            plans["gouraud"].append(("normal", (vtx, (c + 2) % 3)))
            plans["gouraud"].append(("colour", (vtx, c)))
            plans["phong_reflection_shadow"].append(("position", (vtx, (c + 2) % 3)))
            plans["phong_reflection_shadow"].append(("normal", (vtx, c)))
    # texture / specular-map entries: all texels of the small atlas rows that are likely hit are too many; sample
    if second:
        tw, th = base["texture"].shape[:2]
        for u in range(tw):
            for sh in ("gouraud_texture", "phong", "phong_darboux"):
                plans[sh].append(("texture", (u, (3 * u + 1) % th, u % 3)))
            plans["phong_darboux"].append(("normal_map", (u, (5 * u + 2) % th, (u + 1) % 3)))
        for u in range(0, base["atlas"].shape[0], 4):
            plans["phong_reflection"].append(("atlas", (u, (u // 4) % base["atlas"].shape[1], u % 3)))
        for u in range(0, base["specular_map"].shape[0], 2):
            plans["phong_reflection"].append(("specular_map", (u, u % 2)))
    else:
        for u in range(0, base["atlas"].shape[0], 3):
            plans["phong_reflection_shadow"].append(("atlas", (u, (u // 3) % base["atlas"].shape[1], u % 3)))
        for u in range(base["specular_map"].shape[0]):
            plans["phong_reflection_shadow"].append(("specular_map", (u, u % 2)))
    out = {"wz": WZ.astype(np.float32), "wc": WC.astype(np.float32)}
    for shader_name, plan in plans.items():
        names, idxs, grads = [], [], []
        l0 = loss(shader_name, base)
        for name, idx in plan:
            h = 1e-6
            vp, vm = {k: a.copy() for k, a in base.items()}, {k: a.copy() for k, a in base.items()}
            vp[name][idx] += h
            vm[name][idx] -= h
            lp, lm = loss(shader_name, vp), loss(shader_name, vm)
            g = (lp - lm) / (2 * h)
            # a visibility / texel / shadow change inside the step shows up as a jump: the second difference is then
            # of the order of the first one instead of O(h^2)
            ok = abs(lp - 2 * l0 + lm) <= 1e-3 * abs(lp - lm) + 1e-13
            print(f"  {shader_name:24s} {name}{list(idx)}: {g: .8e} {'' if ok else 'DROPPED (not smooth inside the step)'}",
                  flush=True)
            if ok:
                names.append(name); idxs.append(tuple(idx) + (-1,) * (3 - len(idx))); grads.append(g)
        out[f"{shader_name}/names"] = np.array(names)
        out[f"{shader_name}/index"] = np.array(idxs, dtype=np.int32)
        out[f"{shader_name}/grad"] = np.array(grads, dtype=np.float64)
        out[f"{shader_name}/loss"] = np.float64(l0)
    dst = os.path.join(ROOT, "tests", "golden", "reference_grad_more.npz" if second else "reference_grad.npz")
    np.savez_compressed(dst, **out)
    print(f"wrote {dst} in {time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
