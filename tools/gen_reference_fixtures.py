#!/usr/bin/env python
"""Run the UNMODIFIED reference (`/root/reference/renderer`) on small scenes and store its inputs and outputs as
golden fixtures (`tests/golden/reference_run.npz`).

jax / jaxlib are not installable in the build image; the reference runs on the NumPy stand-in under
`tools/jax_numpy_shim/` (see its README): eager float32 NumPy, `vmap` as a Python loop.  That pins the reference's
algorithm -- conventions, clamping / wrapping, tie-breaking, the seven shaders' formulas, the shadow pass, merge_objects,
camera construction -- as executed by the reference's own code; last-bit XLA rounding is outside its reach.

  python tools/gen_reference_fixtures.py [/root/reference]        # ~6 minutes
"""
from __future__ import annotations

import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if (len(sys.argv) > 1 and os.path.isdir(sys.argv[1])) else "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tools", "jax_numpy_shim"))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import numpy as np  # noqa: E402
import jax.numpy as jnp  # noqa: E402  (the shim)
import renderer as R  # noqa: E402  (the reference)
from renderer.shaders.depth import DepthExtraInput, DepthShader  # noqa: E402
from renderer.shaders.gouraud import GouraudExtraInput, GouraudShader  # noqa: E402
from renderer.shaders.gouraud_texture import GouraudTextureExtraInput, GouraudTextureShader  # noqa: E402
from renderer.shaders.phong import PhongTextureExtraInput, PhongTextureShader  # noqa: E402
from renderer.shaders.phong_darboux import PhongTextureDarbouxExtraInput, PhongTextureDarbouxShader  # noqa: E402
from renderer.shaders.phong_reflection import (  # noqa: E402
    PhongReflectionTextureExtraInput, PhongReflectionTextureShader)
from renderer.shaders.phong_reflection_shadow import (  # noqa: E402
    PhongReflectionShadowTextureExtraInput, PhongReflectionShadowTextureShader)
from renderer.shadow import Shadow  # noqa: E402

OUT = {}


def J(x, dtype=None):
    return jnp.asarray(np.asarray(x), dtype=dtype)


def put(prefix, **arrays):
    for k, v in arrays.items():
        OUT[f"{prefix}/{k}"] = np.asarray(v)


def soup(seed, n_tri, W, H, tex=8):
    """Random triangle soup around the origin + all attributes (numpy twin of tests.helpers.random_mesh_scene)."""
    rng = np.random.default_rng(seed)
    V = 3 * n_tri
    centres = (rng.random((n_tri, 1, 3), dtype=np.float32) - 0.5) * 2.0
    pos = (centres + (rng.random((n_tri, 3, 3), dtype=np.float32) - 0.5) * 1.4).reshape(V, 3)
    s = dict(
        W=W, H=H, pos=pos, nrm=rng.standard_normal((V, 3)).astype(np.float32),
        uv_texel=(rng.random((V, 2), dtype=np.float32) * tex * 1.5 - 2.0),
        uv01=(rng.random((V, 2), dtype=np.float32) * 3.0 - 1.0),
        col=rng.random((V, 3), dtype=np.float32), faces=np.arange(V, dtype=np.int32).reshape(n_tri, 3),
        texture=rng.random((tex, tex + 3, 3), dtype=np.float32),
        normal_map=rng.standard_normal((tex, tex + 3, 3)).astype(np.float32),
        light_dir=np.array((0.3, 0.5, 0.8), np.float32), light_col=np.array((1.0, 0.9, 0.8), np.float32),
        eye=np.array((2.0, 2.5, 1.5), np.float32), rng=rng)
    return s


def camera_for(s):
    cp = R.CameraParameters(viewWidth=s["W"], viewHeight=s["H"], position=J(s["eye"]), target=jnp.zeros(3),
                            up=jnp.array((0.0, 0.0, 1.0)))
    return R.Renderer.create_camera_from_parameters(cp)


def buffers(s, with_canvas=True):
    z = jnp.ones((s["W"], s["H"]))
    c = jnp.full((s["W"], s["H"], 3), 0.25)
    return R.Buffers(zbuffer=z, targets=(c,) if with_canvas else ())


def run_soup(seed, n_tri=40, W=28, H=20):
    s = soup(seed, n_tri, W, H)
    cam = camera_for(s)
    light = R.LightSource(direction=J(s["light_dir"]), colour=J(s["light_col"]))
    pre = f"soup{seed}"
    put(pre, position=s["pos"], normal=s["nrm"], uv_texel=s["uv_texel"], uv01=s["uv01"], colour=s["col"],
        faces=s["faces"], texture=s["texture"], normal_map=s["normal_map"], light_direction=s["light_dir"],
        light_colour=s["light_col"], world_to_clip=cam.world_to_clip, viewport=cam.viewport,
        world_to_eye_norm=cam.world_to_eye_norm, view=cam.view, W=W, H=H)
    faces = J(s["faces"])
    pos, nrm = J(s["pos"]), J(s["nrm"])
    cases = [
        ("depth", DepthShader, DepthExtraInput(position=pos), False),
        ("gouraud", GouraudShader, GouraudExtraInput(position=pos, colour=J(s["col"]), normal=nrm, light=light), True),
        ("gouraud_texture", GouraudTextureShader,
         GouraudTextureExtraInput(position=pos, normal=nrm, uv=J(s["uv_texel"]), light=light, texture=J(s["texture"])),
         True),
        ("phong", PhongTextureShader,
         PhongTextureExtraInput(position=pos, normal=nrm, uv=J(s["uv_texel"]), light=light, texture=J(s["texture"])),
         True),
        ("phong_darboux", PhongTextureDarbouxShader,
         PhongTextureDarbouxExtraInput(position=pos, normal=nrm, uv=J(s["uv_texel"]), light=light,
                                       texture=J(s["texture"]), normal_map=J(s["normal_map"]),
                                       id_to_face=J(np.repeat(np.arange(n_tri, dtype=np.int32), 3)),
                                       faces_indices=faces), True),
    ]
    # atlas shaders
    rng = s["rng"]
    n_obj, tw, th = 3, 8, 6
    shapes = np.array([[8, 6], [5, 4], [8, 3]], np.int32)
    atlas = rng.random((n_obj * tw, th, 3), dtype=np.float32)
    spec = rng.random((n_obj * 2, 2), dtype=np.float32) * 6 + 0.5
    tix = np.repeat(rng.integers(0, n_obj, n_tri).astype(np.int32), 3)
    lde = np.array((0.2, 0.3, 0.9), np.float32)
    amb, dif, spe = (np.array(v, np.float32) for v in ((0.3, 0.2, 0.1), (0.5, 0.6, 0.7), (0.2, 0.3, 0.4)))
    put(pre, texture_shape=shapes, atlas=atlas, specular_map=spec, texture_index=tix, light_dir_eye=lde,
        ambient=amb, diffuse=dif, specular=spe, texture_offset=tw)
    base = dict(position=pos, normal=nrm, uv=J(s["uv01"]), light=light, light_dir_eye=J(lde), texture_shape=J(shapes),
                texture_index=J(tix), texture_offset=J(np.int32(tw)), texture=J(atlas), specular_map=J(spec), ambient=J(amb),
                diffuse=J(dif), specular=J(spe))
    cases.append(("phong_reflection", PhongReflectionTextureShader, PhongReflectionTextureExtraInput(**base), True))
    sm0 = jnp.full((W, H), float(np.finfo(np.float32).max))
    ldir_s, strength = np.array((0.4, 0.3, 0.9), np.float32), np.array((0.6, 0.5, 0.4), np.float32)
    shadow = Shadow.render_shadow_map(shadow_map=sm0, verts=pos, faces=faces, light_direction=J(ldir_s),
                                      viewport_matrix=cam.viewport, centre=jnp.zeros(3),
                                      up=jnp.array((0.0, 0.0, 1.0)), strength=J(strength), offset=0.05)
    put(pre, shadow_map=shadow.shadow_map, shadow_world_to_clip=shadow.camera.world_to_clip,
        shadow_viewport=shadow.camera.viewport, shadow_light_direction=ldir_s, shadow_strength=strength)
    cases.append(("phong_reflection_shadow", PhongReflectionShadowTextureShader,
                  PhongReflectionShadowTextureExtraInput(**base, shadow=shadow, camera=cam), True))
    for name, shader, extra, canvas in cases:
        t = time.time()
        out = R.render(cam, shader, buffers(s, canvas), faces, extra)
        put(f"{pre}/{name}", zbuffer=out.zbuffer, **({"canvas": out.targets[0]} if canvas else {}))
        print(f"  {pre}/{name}: {time.time() - t:.1f}s, covered {(np.asarray(out.zbuffer) != 1.0).mean():.2f}", flush=True)


def run_facade():
    """Renderer.get_camera_image (merge_objects -> camera -> shadow pass -> phong_reflection_shadow) and the
    intermediate merged model / camera, on a cube over a ground box with a capsule."""
    rng = np.random.default_rng(7)
    ground = R.create_cube(half_extents=jnp.array((3.0, 3.0, 0.05)), texture_scaling=jnp.array(4.0),
                           diffuse_map=J(rng.random((6, 5, 3), dtype=np.float32)), specular_map=jnp.ones((6, 5)) * 2.0)
    cube = R.create_cube(half_extents=jnp.array((0.5, 0.4, 0.3)), texture_scaling=jnp.array(1.0),
                         diffuse_map=J(rng.random((2, 2, 3), dtype=np.float32)), specular_map=jnp.ones((2, 2)) * 3.0)
    cap = R.create_capsule(radius=jnp.array(0.3), half_height=jnp.array(0.4), up_axis=R.UpAxis.Z,
                           diffuse_map=J(rng.random((1, 1, 3), dtype=np.float32)), specular_map=jnp.ones((1, 1)) * 2.0)
    objs = [
        R.ModelObject(model=ground),
        R.ModelObject(model=cube).replace_with_position(jnp.array((0.2, -0.3, 0.8)))
         .replace_with_orientation(R.quaternion(jnp.array((0.3, 0.4, 0.5)), jnp.array(40.0))),
        R.ModelObject(model=cap, local_scaling=jnp.array((1.0, 1.2, 0.9)))
         .replace_with_position(jnp.array((-0.9, 0.6, 0.9))),
    ]
    W, H = 24, 18
    cp = R.CameraParameters(viewWidth=W, viewHeight=H, position=jnp.array((3.0, -3.5, 2.5)),
                            target=jnp.array((0.0, 0.0, 0.5)), up=jnp.array((0.0, 0.0, 1.0)), hfov=58.0,
                            vfov=58.0 * H / W)
    light = R.LightParameters()
    sp = R.ShadowParameters(centre=jnp.array((0.0, 0.0, 0.5)))
    merged = R.merge_objects(objs)
    cam = R.Renderer.create_camera_from_parameters(cp)
    pre = "facade"
    for i, o in enumerate(objs):
        put(f"{pre}/obj{i}", verts=o.model.verts, norms=o.model.norms, uvs=o.model.uvs, faces=o.model.faces,
            faces_norm=o.model.faces_norm, faces_uv=o.model.faces_uv, diffuse_map=o.model.diffuse_map,
            specular_map=o.model.specular_map, local_scaling=o.local_scaling, transform=o.transform)
    put(f"{pre}/merged", **{k: getattr(merged, k) for k in merged._fields})
    put(f"{pre}/camera", **{k: getattr(cam, k) for k in cam._fields})
    put(f"{pre}/camera_parameters", position=cp.position, target=cp.target, up=cp.up, hfov=cp.hfov, vfov=cp.vfov,
        W=W, H=H)
    put(f"{pre}/shadow_parameters", centre=sp.centre, up=sp.up, strength=sp.strength, offset=sp.offset)
    for name, spar in (("with_shadow", sp), ("no_shadow", None)):
        t = time.time()
        img = R.Renderer.get_camera_image(objs, light, cp, W, H, shadow_param=spar)
        put(f"{pre}/{name}", canvas=img)
        print(f"  {pre}/{name}: {time.time() - t:.1f}s", flush=True)


def run_edge_cases():
    """Depth-shader scenes for the conventions that were restated from reading the reference: the back-facing
    triangle 0 quirk, geometry behind / straddling the camera plane, degenerate and duplicate triangles, incoming
    z-buffer values closer than the new fragments (no depth test against the old buffer), an empty face list."""
    W, H = 16, 12
    cp = R.CameraParameters(viewWidth=W, viewHeight=H, position=jnp.array((0.0, -3.0, 0.5)), target=jnp.zeros(3),
                            up=jnp.array((0.0, 0.0, 1.0)))
    cam = R.Renderer.create_camera_from_parameters(cp)
    put("edge", world_to_clip=cam.world_to_clip, viewport=cam.viewport, W=W, H=H)
    scenes = {}
    # triangle 0 back-facing (clockwise seen from the camera), triangle 1 front-facing, partly overlapping
    scenes["tri0_backfacing"] = (
        np.array([[-1.2, 0.0, -0.8], [0.0, 0.0, 1.2], [1.2, 0.0, -0.8],
                  [-0.2, -0.5, -0.6], [1.4, -0.5, -0.6], [0.6, -0.5, 0.9]], np.float32),
        np.array([[0, 1, 2], [3, 4, 5]], np.int32), 1.0)
    # same with the winding of triangle 0 flipped (front-facing): the control
    scenes["tri0_frontfacing"] = (scenes["tri0_backfacing"][0], np.array([[0, 2, 1], [3, 4, 5]], np.int32), 1.0)
    # behind the camera, straddling the camera plane, and a normal one
    scenes["behind_and_straddling"] = (
        np.array([[-1.0, -6.0, -0.5], [1.0, -6.0, -0.5], [0.0, -6.0, 1.0],       # entirely behind
                  [-0.8, -4.0, -0.4], [0.9, 1.0, -0.4], [0.0, 1.0, 0.9],          # crosses the camera plane
                  [-0.5, 0.5, -0.3], [0.7, 0.5, -0.3], [0.1, 0.5, 0.8]], np.float32),
        np.array([[0, 1, 2], [3, 4, 5], [6, 7, 8]], np.int32), 1.0)
    # zero-area, repeated-vertex and duplicate triangles around a regular one
    scenes["degenerate_and_duplicates"] = (
        np.array([[-0.9, 0.0, -0.6], [0.9, 0.0, -0.6], [0.0, 0.0, 0.9], [0.3, 0.0, 0.1]], np.float32),
        np.array([[3, 3, 3], [0, 1, 1], [0, 1, 2], [0, 1, 2], [0, 2, 1]], np.int32), 1.0)
    # incoming z-buffer already closer than every fragment
    scenes["old_z_closer"] = (scenes["tri0_frontfacing"][0], scenes["tri0_frontfacing"][1], 0.25)
    for name, (pos, faces, z_init) in scenes.items():
        out = R.render(cam, DepthShader, R.Buffers(zbuffer=jnp.full((W, H), z_init), targets=()), J(faces),
                       DepthExtraInput(position=J(pos)))
        put(f"edge/{name}", position=pos, faces=faces, z_init=np.float32(z_init), zbuffer=out.zbuffer)
        z = np.asarray(out.zbuffer)
        print(f"  edge/{name}: written {(z != z_init).sum()} of {z.size}", flush=True)


def run_host_helpers():
    """Outputs of the reference's host-side helpers (shapes, quaternions / rotations, camera matrix builders,
    display utilities) for fixed arguments."""
    rng = np.random.default_rng(11)
    tex = rng.random((3, 4, 3), dtype=np.float32)
    cube = R.create_cube(half_extents=jnp.array((0.5, 1.5, 2.0)), texture_scaling=jnp.array((2.0, 3.0)),
                         diffuse_map=J(tex), specular_map=jnp.ones((3, 4)) * 2.5)
    put("helpers/cube", **{k: getattr(cube, k) for k in cube._fields})
    for axis in (R.UpAxis.X, R.UpAxis.Y, R.UpAxis.Z):
        cap = R.create_capsule(radius=jnp.array(0.25), half_height=jnp.array(0.75), up_axis=axis, diffuse_map=J(tex),
                               specular_map=jnp.ones((3, 4)) * 2.5)
        put(f"helpers/capsule_{int(axis)}", verts=cap.verts, norms=cap.norms, uvs=cap.uvs, faces=cap.faces)
    axis, angle = jnp.array((0.3, -0.5, 0.8)), jnp.array(37.0)
    q1, q2 = R.quaternion(axis, angle), R.quaternion(jnp.array((1.0, 0.2, 0.1)), jnp.array(-110.0))
    put("helpers/rotation", axis=axis, angle=angle, quaternion=q1, quaternion2=q2, quaternion_mul=R.quaternion_mul(q1, q2),
        rotation_matrix=R.rotation_matrix(axis, angle), normalise=R.normalise(jnp.array((3.0, -4.0, 12.0))))
    mo = R.ModelObject(model=cube).replace_with_orientation(q1).replace_with_position(jnp.array((1.0, 2.0, 3.0)))
    put("helpers/model_object", transform=mo.transform)
    eye, centre, up = jnp.array((1.0, -2.0, 3.0)), jnp.array((0.2, 0.1, -0.3)), jnp.array((0.0, 0.0, 1.0))
    C = R.Camera
    put("helpers/camera", eye=eye, centre=centre, up=up, view=C.view_matrix(eye, centre, up),
        view_inv=C.view_matrix_inv(eye, centre, up),
        perspective=C.perspective_projection_matrix(jnp.array(40.0), jnp.array(1.6), jnp.array(0.1), jnp.array(50.0)),
        orthographic=C.orthographic_projection_matrix(jnp.array(-2.0), jnp.array(3.0), jnp.array(-1.0), jnp.array(1.5),
                                                       jnp.array(0.5), jnp.array(20.0)),
        viewport=C.viewport_matrix(jnp.array((1.0, 2.0)), jnp.array((640.0, 480.0)), jnp.array(2.0)),
        world_to_screen=C.world_to_screen_matrix(320, 200))
    canvas = rng.random((5, 7, 3), dtype=np.float32) * 1.4 - 0.2
    put("helpers/utils", canvas=canvas, transposed=R.transpose_for_display(J(canvas)),
        transposed_noflip=R.transpose_for_display(J(canvas), flip_vertical=False))
    raw = (rng.integers(0, 256, 4 * 3 * 3)).astype(np.float32)
    put("helpers/utils", pytiny_raw=raw, pytiny_texture=R.build_texture_from_PyTinyrenderer(J(raw), 4, 3))
    print("  helpers done", flush=True)


def main():
    t0 = time.time()
    run_host_helpers()
    run_edge_cases()
    for seed in (0, 1):
        run_soup(seed)
    run_facade()
    dst = os.path.join(ROOT, "tests", "golden", "reference_run.npz")
    np.savez_compressed(dst, **OUT)
    print(f"wrote {dst}: {len(OUT)} arrays, {os.path.getsize(dst) / 1024:.0f} KB, {time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
