"""JAX host binding of ``libjr_b200.so`` through the XLA FFI (``ffi/jr_ffi.cc``): what a maintainer of the reference
adds so that ``renderer.pipeline.render`` (``renderer/pipeline.py:470-537``) dispatches the seven built-in shaders to
the B200 kernels, with ``jax.grad`` (``custom_vjp``) and ``jax.vmap`` (``vmap_method="expand_dims"``) support.

NOT importable in the image this repository was developed in (no jax / jaxlib): it is source for the JAX side of the
boundary, kept next to the C++ handlers; ``tests/test_ffi_sources.py`` checks it parses and that its operand tables
agree with ``jr_ffi.cc``.  Needs ``jax >= 0.4.31`` (``jax.ffi``; the reference pins 0.4.13, ``poetry.lock:276-277``).

    import ffi.jax_binding as jr_jax
    jr_jax.register("ffi/libjr_ffi.so")
    renderer.pipeline.render = jr_jax.render        # drop-in
"""
from __future__ import annotations

import ctypes
from functools import partial
from typing import Any, Dict, Sequence, Tuple

# (name, un-batched rank, dtype, differentiable) in the order jr_ffi.cc expects -- keep in sync with kOperands there
COMMON = (("world_to_clip", 2, "f32", True), ("viewport", 2, "f32", True), ("position", 2, "f32", True),
          ("faces", 2, "i32", False))
PHONG_REFLECTION = (
    ("world_to_eye_norm", 2, "f32", True), ("normal", 2, "f32", True), ("uv", 2, "f32", False),
    ("light_colour", 1, "f32", True), ("light_dir_eye", 1, "f32", True), ("ambient", 1, "f32", True),
    ("diffuse", 1, "f32", True), ("specular", 1, "f32", True), ("texture", 3, "f32", True),
    ("specular_map", 2, "f32", True), ("texture_shape", 2, "i32", False), ("texture_index", 1, "i32", False))
OPERANDS: Dict[str, Tuple[Tuple[str, int, str, bool], ...]] = {
    "depth": COMMON,
    "gouraud": COMMON + (("normal", 2, "f32", True), ("colour", 2, "f32", True), ("light_direction", 1, "f32", True),
                         ("light_colour", 1, "f32", True)),
    "gouraud_texture": COMMON + (("normal", 2, "f32", True), ("uv", 2, "f32", False), ("light_direction", 1, "f32", True),
                                 ("light_colour", 1, "f32", True), ("texture", 3, "f32", True)),
    "phong": COMMON + (("world_to_eye_norm", 2, "f32", True), ("normal", 2, "f32", True), ("uv", 2, "f32", False),
                       ("light_direction", 1, "f32", True), ("light_colour", 1, "f32", True), ("texture", 3, "f32", True)),
    "phong_darboux": COMMON + (("world_to_eye_norm", 2, "f32", True), ("normal", 2, "f32", True), ("uv", 2, "f32", True),
                               ("light_direction", 1, "f32", True), ("light_colour", 1, "f32", True),
                               ("texture", 3, "f32", True), ("normal_map", 3, "f32", True),
                               ("id_to_face", 1, "i32", False), ("faces_indices", 2, "i32", False)),
    "phong_reflection": COMMON + PHONG_REFLECTION,
    "phong_reflection_shadow": COMMON + PHONG_REFLECTION + (
        ("shadow_map", 2, "f32", False), ("shadow_strength", 1, "f32", True),
        ("shadow_world_to_clip", 2, "f32", False), ("shadow_viewport", 2, "f32", False)),
}
SHADER_CLASS_TO_NAME = {
    "DepthShader": "depth", "GouraudShader": "gouraud", "GouraudTextureShader": "gouraud_texture",
    "PhongTextureShader": "phong", "PhongTextureDarbouxShader": "phong_darboux",
    "PhongReflectionTextureShader": "phong_reflection", "PhongReflectionShadowTextureShader": "phong_reflection_shadow",
}


def register(ffi_library: str, abi_library: str = "jaxrenderer_b200/lib/libjr_b200.so") -> None:
    """Load the handlers and register the 14 FFI targets for the CUDA platform."""
    import jax

    global _abi
    _abi = ctypes.CDLL(abi_library, mode=ctypes.RTLD_GLOBAL)    # the handlers link against it
    lib = ctypes.CDLL(ffi_library)
    for name in OPERANDS:
        for direction in ("forward", "backward"):
            target = f"jr_{name}_{direction}"
            jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(lib, target + "_ffi")), platform="CUDA")


def _collect(name: str, camera: Any, face_indices: Any, extra: Any) -> Dict[str, Any]:
    """The reference's ``extra`` NamedTuple of a built-in shader -> operand name -> array (cf.
    ``jaxrenderer_b200/pipeline.py::_collect``)."""
    a = {"world_to_clip": camera.world_to_clip, "viewport": camera.viewport, "position": extra.position,
         "faces": face_indices}
    if name == "depth":
        return a
    a["normal"], a["light_colour"] = extra.normal, extra.light.colour
    if name == "gouraud":
        a["colour"], a["light_direction"] = extra.colour, extra.light.direction
        return a
    a["uv"], a["texture"] = extra.uv, extra.texture
    if name in ("gouraud_texture", "phong", "phong_darboux"):
        a["light_direction"] = extra.light.direction
    if name != "gouraud_texture":
        a["world_to_eye_norm"] = camera.world_to_eye_norm
    if name == "phong_darboux":
        a["normal_map"], a["id_to_face"], a["faces_indices"] = extra.normal_map, extra.id_to_face, extra.faces_indices
    if name.startswith("phong_reflection"):
        a.update(light_dir_eye=extra.light_dir_eye, ambient=extra.ambient, diffuse=extra.diffuse,
                 specular=extra.specular, specular_map=extra.specular_map, texture_shape=extra.texture_shape,
                 texture_index=extra.texture_index)
    if name == "phong_reflection_shadow":
        sh = extra.shadow
        a.update(shadow_map=sh.shadow_map, shadow_strength=sh.strength, shadow_world_to_clip=sh.camera.world_to_clip,
                 shadow_viewport=sh.camera.viewport)
    return a


def _workspace(n_bytes: int):
    import jax.numpy as jnp

    return jnp.empty((max(int(n_bytes), 16),), dtype=jnp.uint8)


def _scratch_bytes(name: str, B: int, W: int, H: int, T: int, backward: bool) -> int:
    """Upper bound of ``jr_workspace_bytes`` / ``jr_backward_workspace_bytes`` from static shapes (the exact figure
    needs device pointers; XLA allocates scratch at trace time): binned-visibility records + bitmasks + the spill list,
    attribute records, and for backward the sort buffers (16 B x pixels x 4 entries x 3 passes)."""
    npix = W * H
    fwd = B * T * (64 + 4 + 4 + 4) + B * ((W + 63) // 64) * ((H + 63) // 64) * ((T + 31) // 32) * 4 + B * min(T, npix) * 176 + (1 << 16)
    return fwd + (B * npix * (16 * 4 * 3 + 52 * 4) + B * T * 320 if backward else 0)


def make_render_fn(name: str, texture_offset: int = 0):
    """``jax.custom_vjp`` function ``f(zbuffer, canvas_or_None, *operands) -> (zbuffer', canvas')`` of one shader."""
    import jax
    import jax.numpy as jnp

    ops = OPERANDS[name]
    has_canvas = name != "depth"
    diff_idx = [i for i, o in enumerate(ops) if o[3]]

    def _dims(zbuffer, operands):
        B = zbuffer.shape[0] if zbuffer.ndim == 3 else 1
        W, H = zbuffer.shape[-2:]
        T = operands[3].shape[-2]
        return B, W, H, T

    def _call_forward(zbuffer, canvas, *operands):
        z3 = zbuffer if zbuffer.ndim == 3 else zbuffer[None]
        B, W, H, T = _dims(z3, operands)
        outs = [jax.ShapeDtypeStruct(z3.shape, jnp.float32)]
        ins = list(operands) + [z3]
        aliases = {len(operands): 0}                       # the reference donates its buffers (pipeline.py:466)
        if has_canvas:
            c4 = canvas if canvas.ndim == 4 else canvas[None]
            outs.append(jax.ShapeDtypeStruct(c4.shape, jnp.float32))
            ins.append(c4)
            aliases[len(operands) + 1] = 1
        outs.append(jax.ShapeDtypeStruct(z3.shape, jnp.int32))      # tri_id G-buffer
        ins.append(_workspace(_scratch_bytes(name, B, W, H, T, False)))
        res = jax.ffi.ffi_call(f"jr_{name}_forward", tuple(outs), vmap_method="expand_dims",
                               input_output_aliases=aliases)(*ins, texture_offset=jnp.int32(texture_offset))
        return res

    @jax.custom_vjp
    def f(zbuffer, canvas, *operands):
        res = _call_forward(zbuffer, canvas, *operands)
        squeeze = (lambda t: t[0]) if zbuffer.ndim == 2 else (lambda t: t)
        return (squeeze(res[0]), squeeze(res[1]) if has_canvas else None)

    def fwd(zbuffer, canvas, *operands):
        res = _call_forward(zbuffer, canvas, *operands)
        squeeze = (lambda t: t[0]) if zbuffer.ndim == 2 else (lambda t: t)
        out = (squeeze(res[0]), squeeze(res[1]) if has_canvas else None)
        return out, (operands, res[-1], zbuffer.ndim == 2)          # save the inputs + the triangle-id G-buffer

    def bwd(saved, cts):
        operands, tri_id, squeezed = saved
        d_z, d_c = cts
        B, W, H, T = tri_id.shape[0], tri_id.shape[1], tri_id.shape[2], operands[3].shape[-2]
        d_z = jnp.zeros(tri_id.shape, jnp.float32) if d_z is None else d_z.reshape(tri_id.shape)
        ins = list(operands) + [tri_id, d_z]
        if has_canvas:
            ins.append(jnp.zeros(tri_id.shape + (3,), jnp.float32) if d_c is None else d_c.reshape(tri_id.shape + (3,)))
        ins.append(_workspace(_scratch_bytes(name, B, W, H, T, True)))
        outs = [jax.ShapeDtypeStruct(operands[i].shape, jnp.float32) for i in diff_idx]
        outs.append(jax.ShapeDtypeStruct(tri_id.shape, jnp.float32))
        if has_canvas:
            outs.append(jax.ShapeDtypeStruct(tri_id.shape + (3,), jnp.float32))
        res = jax.ffi.ffi_call(f"jr_{name}_backward", tuple(outs), vmap_method="expand_dims")(
            *ins, texture_offset=jnp.int32(texture_offset))
        grads = [None] * len(operands)
        for k, i in enumerate(diff_idx):
            grads[i] = res[k]
        unsq = (lambda t: t[0]) if squeezed else (lambda t: t)
        d_zin = unsq(res[len(diff_idx)])
        d_cin = unsq(res[len(diff_idx) + 1]) if has_canvas else None
        return (d_zin, d_cin, *grads)

    f.defvjp(fwd, bwd)
    return f


_FNS: Dict[Tuple[str, int], Any] = {}


def render(camera: Any, shader: type, buffers: Any, face_indices: Any, extra: Any, loop_unroll: int = 1) -> Any:
    """Drop-in for ``renderer.pipeline.render``: the seven built-in shaders go to the B200 kernels; anything else is
    rejected (BASELINE.json north_star: no fallback)."""
    del loop_unroll
    name = SHADER_CLASS_TO_NAME.get(getattr(shader, "__name__", ""))
    if name is None:
        raise NotImplementedError(f"shader {shader!r} is not a built-in shader: custom shaders are not supported")
    arrays = _collect(name, camera, face_indices, extra)
    offset = int(getattr(extra, "texture_offset", 0)) if name.startswith("phong_reflection") else 0
    fn = _FNS.setdefault((name, offset), make_render_fn(name, offset))
    canvas = buffers.targets[0] if name != "depth" else None
    z, c = fn(buffers.zbuffer, canvas, *[arrays[o[0]] for o in OPERANDS[name]])
    return type(buffers)(zbuffer=z, targets=() if c is None else (c,))
