"""Host-side evaluators of the five ``Shader`` stages (``renderer/shader.py:103-396`` and ``renderer/shaders/*.py``).

The reference exposes every stage of every shader as a static method that user code may call, compose or introspect
(``vertex``, ``primitive_chooser``, ``interpolate``, ``fragment``, ``mix``).  In this implementation the stages of the
seven built-in shaders run FUSED inside the CUDA kernels and ``pipeline.render`` NEVER calls anything in this module;
these functions only keep that part of the API surface callable with the reference's signatures and semantics (one
vertex / one fragment per call, as the reference defines them before ``vmap``): plain tensor code in the host framework,
on whatever device the arguments live.  They are not a fallback -- a custom ``Shader`` subclass is still rejected by
``render`` -- and they are checked stage by stage against the CPU oracle in ``tests/test_stage_methods.py``.

Conventions as in the reference: ``barycentric_clip`` are the perspective-correct weights, varyings of a triangle are
stacked on axis 0 (3, ...), a batch of primitives on axis 0 (P, ...).
"""
from __future__ import annotations

from typing import Any, Tuple

import torch

from .geometry import Camera, Interpolation, interpolate, normalise, to_homogeneous
from .shader import MixerOutput, PerFragment, PerVertex

Tensor = torch.Tensor


def _t(x: Any, like: Any = None, dtype: torch.dtype = torch.float32) -> Tensor:
    if isinstance(x, Tensor):
        return x.to(dtype) if x.dtype != dtype and dtype.is_floating_point == x.dtype.is_floating_point else x
    dev = like.device if isinstance(like, Tensor) else None
    return torch.as_tensor(x, dtype=dtype, device=dev)


def _tree_map(fn: Any, tree: Any) -> Any:
    """``jax.tree_util.tree_map`` over the (named) tuples the stages pass around."""
    if isinstance(tree, tuple):
        mapped = [_tree_map(fn, v) for v in tree]
        return type(tree)(*mapped) if hasattr(tree, "_fields") else tuple(mapped)
    if tree is None:
        return None
    return fn(tree)


# ------------------------------------------------------------------------------------------------ base class stages
def base_primitive_chooser(gl_FragCoord: Tensor, gl_FrontFacing: Tensor, gl_PointCoord: Tensor, keeps: Tensor,
                           values: Any, barycentric_screen: Tensor, barycentric_clip: Tensor) -> Tuple[Any, ...]:
    """``shader.py:159-251``: the closest primitive that is kept and front-facing -- ``argmin`` (first index on ties)
    over the window-space depths, ``inf`` for the others; with no candidate the index is 0.  One primitive is returned,
    with a leading axis of length 1."""
    depths = torch.where(keeps & gl_FrontFacing, gl_FragCoord[:, 2], torch.full_like(gl_FragCoord[:, 2], float("inf")))
    idx = int(torch.argmin(depths))   # torch.argmin returns the first minimal index, as jnp.argmin does
    pick = lambda x: x[idx:idx + 1]   # noqa: E731  (lax.dynamic_slice_in_dim(x, idx, 1, axis=0))
    return (pick(gl_FragCoord), pick(gl_FrontFacing), pick(gl_PointCoord), pick(keeps), _tree_map(pick, values),
            pick(barycentric_screen), pick(barycentric_clip))


def base_interpolate(values: Any, barycentric_screen: Tensor, barycentric_clip: Tensor) -> Any:
    """``shader.py:257-290``: every field ``smooth`` (perspective-correct)."""
    return _tree_map(lambda v: interpolate(_t(v), barycentric_screen, barycentric_clip, Interpolation.SMOOTH), values)


def base_fragment(gl_FragCoord: Tensor, gl_FrontFacing: Tensor, gl_PointCoord: Tensor, varying: Any,
                  extra: Any) -> Tuple[PerFragment, Any]:
    """``shader.py:296-337``: writes nothing; depth defaults to ``gl_FragCoord[2]`` downstream."""
    return PerFragment(use_default_depth=torch.tensor(True)), varying


def base_mix(gl_FragDepth: Tensor, keeps: Tensor, extra: Any) -> Tuple[MixerOutput, Any]:
    """``shader.py:343-396``: the kept fragment of minimal depth (``argmin`` over ``where(keeps, depth, inf)``)."""
    depths = torch.where(keeps, gl_FragDepth, torch.full_like(gl_FragDepth, float("inf")))
    idx = int(torch.argmin(depths))
    return MixerOutput(keep=keeps[idx], zbuffer=depths[idx]), _tree_map(lambda x: x[idx], extra)


def _keep_all(*flags: Any) -> Tensor:
    out = torch.as_tensor(True)
    for f in flags:
        out = out.to(f.device) & f if isinstance(f, Tensor) else out & torch.as_tensor(bool(f))
    return out


def _clip_position(camera: Camera, position: Tensor, vid: Any) -> Tensor:
    return camera.to_clip(to_homogeneous(_t(position)[vid]))


def _texel(uv: Tensor, texture: Tensor) -> Tensor:
    """``texture[floor(uv) % texture.shape[:2]]`` (``gouraud_texture.py:124-126``)."""
    u = torch.floor(uv[..., 0]).to(torch.int64) % texture.shape[0]
    v = torch.floor(uv[..., 1]).to(torch.int64) % texture.shape[1]
    return texture[u, v]


def _gather2(arr: Tensor, u: Tensor, v: Tensor) -> Tensor:
    """``arr[u, v]`` with jnp indexing semantics: one negative wrap, then clamp."""
    n0, n1 = arr.shape[0], arr.shape[1]
    u = torch.where(u < 0, u + n0, u).clamp(0, n0 - 1)
    v = torch.where(v < 0, v + n1, v).clamp(0, n1 - 1)
    return arr[u, v]


# ------------------------------------------------------------------------------------------------ depth
def depth_vertex(gl_VertexID: Any, gl_InstanceID: Any, camera: Camera, extra: Any):
    """``shaders/depth.py:47-61``."""
    from .shaders.depth import DepthExtraFragmentData

    return PerVertex(gl_Position=_clip_position(camera, extra.position, gl_VertexID)), DepthExtraFragmentData()


# ------------------------------------------------------------------------------------------------ gouraud
def gouraud_vertex(gl_VertexID: Any, gl_InstanceID: Any, camera: Camera, extra: Any):
    """``shaders/gouraud.py:58-88``: colour * light colour * (n . l) per vertex."""
    from .shaders.gouraud import GouraudExtraFragmentData

    n = normalise(_t(extra.normal)[gl_VertexID])
    intensity = torch.dot(n, normalise(_t(extra.light.direction, n)))
    colour = _t(extra.colour)[gl_VertexID] * _t(extra.light.colour, n) * intensity
    return (PerVertex(gl_Position=_clip_position(camera, extra.position, gl_VertexID)),
            GouraudExtraFragmentData(colour=colour))


def gouraud_fragment(gl_FragCoord: Tensor, gl_FrontFacing: Tensor, gl_PointCoord: Tensor, varying: Any, extra: Any):
    """``shaders/gouraud.py:94-124``: keep only front-facing fragments whose colour is non-negative."""
    built_in = base_fragment(gl_FragCoord, gl_FrontFacing, gl_PointCoord, varying, extra)[0]
    keeps = _keep_all(built_in.keeps, gl_FrontFacing, (varying.colour >= 0).all())
    return PerFragment(keeps=keeps, use_default_depth=built_in.use_default_depth), varying


def gouraud_mix(gl_FragDepth: Tensor, keeps: Tensor, extra: Any):
    from .shaders.gouraud import GouraudExtraMixerOutput

    out, picked = base_mix(gl_FragDepth, keeps, extra)
    return out, GouraudExtraMixerOutput(canvas=picked.colour)


# ------------------------------------------------------------------------------------------------ gouraud + texture
def gouraud_texture_vertex(gl_VertexID: Any, gl_InstanceID: Any, camera: Camera, extra: Any):
    """``shaders/gouraud_texture.py:68-101``: light colour * (n . l) and uv per vertex."""
    from .shaders.gouraud_texture import GouraudTextureExtraFragmentData

    n = normalise(_t(extra.normal)[gl_VertexID])
    intensity = torch.dot(n, normalise(_t(extra.light.direction, n)))
    return (PerVertex(gl_Position=_clip_position(camera, extra.position, gl_VertexID)),
            GouraudTextureExtraFragmentData(colour=_t(extra.light.colour, n) * intensity,
                                            uv=_t(extra.uv)[gl_VertexID]))


def gouraud_texture_fragment(gl_FragCoord: Tensor, gl_FrontFacing: Tensor, gl_PointCoord: Tensor, varying: Any,
                             extra: Any):
    """``shaders/gouraud_texture.py:107-144``."""
    built_in = base_fragment(gl_FragCoord, gl_FrontFacing, gl_PointCoord, varying, extra)[0]
    light_colour = varying.colour
    keeps = _keep_all(built_in.keeps, gl_FrontFacing, (light_colour >= 0).all())
    colour = _texel(varying.uv, _t(extra.texture)) * light_colour
    return (PerFragment(keeps=keeps, use_default_depth=built_in.use_default_depth),
            type(varying)(colour=colour, uv=varying.uv))


def gouraud_texture_mix(gl_FragDepth: Tensor, keeps: Tensor, extra: Any):
    from .shaders.gouraud_texture import GouraudTextureExtraMixerOutput

    out, picked = base_mix(gl_FragDepth, keeps, extra)
    return out, GouraudTextureExtraMixerOutput(canvas=picked.colour)


# ------------------------------------------------------------------------------------------------ phong
def _eye_normal(camera: Camera, normal: Tensor, vid: Any) -> Tensor:
    """``Camera.apply_vec(normalise(n), world_to_eye_norm)`` (``phong.py:92-96``)."""
    return Camera.apply_vec(normalise(_t(normal)[vid]), camera.world_to_eye_norm)


def phong_vertex(gl_VertexID: Any, gl_InstanceID: Any, camera: Camera, extra: Any):
    """``shaders/phong.py:78-104``: eye-space normal and uv per vertex."""
    from .shaders.phong import PhongTextureExtraFragmentData

    return (PerVertex(gl_Position=_clip_position(camera, extra.position, gl_VertexID)),
            PhongTextureExtraFragmentData(normal=_eye_normal(camera, extra.normal, gl_VertexID),
                                          uv=_t(extra.uv)[gl_VertexID], colour=torch.zeros(3, device=_t(extra.position).device)))


def _phong_colour(normal: Tensor, uv: Tensor, extra: Any) -> Tensor:
    light_colour = _t(extra.light.colour, normal) * torch.dot(normalise(normal),
                                                              normalise(_t(extra.light.direction, normal)))
    texture_colour = _texel(uv, _t(extra.texture))
    return torch.where((light_colour >= 0).all(), texture_colour * light_colour, torch.zeros_like(light_colour))


def phong_fragment(gl_FragCoord: Tensor, gl_FrontFacing: Tensor, gl_PointCoord: Tensor, varying: Any, extra: Any):
    """``shaders/phong.py:110-164``: per-fragment n . l; a negative light colour paints black (no discard)."""
    built_in = base_fragment(gl_FragCoord, gl_FrontFacing, gl_PointCoord, varying, extra)[0]
    keeps = _keep_all(built_in.keeps, gl_FrontFacing)
    return (PerFragment(keeps=keeps, use_default_depth=built_in.use_default_depth),
            varying._replace(colour=_phong_colour(varying.normal, varying.uv, extra)))


def phong_mix(gl_FragDepth: Tensor, keeps: Tensor, extra: Any):
    from .shaders.phong import PhongTextureExtraMixerOutput

    out, picked = base_mix(gl_FragDepth, keeps, extra)
    return out, PhongTextureExtraMixerOutput(canvas=picked.colour)


# ------------------------------------------------------------------------------------------------ phong + Darboux
def phong_darboux_vertex(gl_VertexID: Any, gl_InstanceID: Any, camera: Camera, extra: Any):
    """``shaders/phong_darboux.py:120-170``: besides normal and uv every vertex carries ITS triangle
    (``faces_indices[id_to_face[v]]``) in NDC and in uv space."""
    from .geometry import to_cartesian
    from .shaders.phong_darboux import PhongTextureDarbouxExtraFragmentData

    face = _t(extra.faces_indices, dtype=torch.int64)[_t(extra.id_to_face, dtype=torch.int64)[gl_VertexID]]
    tri_clip = camera.to_clip(to_homogeneous(_t(extra.position)[face]))
    return (PerVertex(gl_Position=_clip_position(camera, extra.position, gl_VertexID)),
            PhongTextureDarbouxExtraFragmentData(normal=_eye_normal(camera, extra.normal, gl_VertexID),
                                                 uv=_t(extra.uv)[gl_VertexID], triangle=to_cartesian(tri_clip),
                                                 triangle_uv=_t(extra.uv)[face], colour=torch.zeros(3, device=_t(extra.position).device)))


def phong_darboux_interpolate(values: Any, barycentric_screen: Tensor, barycentric_clip: Tensor) -> Any:
    """``shaders/phong_darboux.py:176-206``: normal and uv smooth, the triangle copies flat (first vertex)."""
    smooth = lambda v: interpolate(_t(v), barycentric_screen, barycentric_clip, Interpolation.SMOOTH)  # noqa: E731
    flat = lambda v: interpolate(_t(v), barycentric_screen, barycentric_clip, Interpolation.FLAT)      # noqa: E731
    return type(values)(normal=smooth(values.normal), uv=smooth(values.uv), triangle=flat(values.triangle),
                        triangle_uv=flat(values.triangle_uv), colour=smooth(values.colour))


def phong_darboux_fragment(gl_FragCoord: Tensor, gl_FrontFacing: Tensor, gl_PointCoord: Tensor, varying: Any,
                           extra: Any):
    """``shaders/phong_darboux.py:212-293``: the normal-map texel is taken from the tangent frame (Darboux basis) of
    the fragment's triangle to eye space: ``B = inv([p1 - p0; p2 - p0; n]) @ [du; dv; 0]`` column-normalised, ``n``
    as third column."""
    built_in = base_fragment(gl_FragCoord, gl_FrontFacing, gl_PointCoord, varying, extra)[0]
    keeps = _keep_all(built_in.keeps, gl_FrontFacing)
    nn = normalise(varying.normal)
    tri, tuv = varying.triangle, varying.triangle_uv
    A = torch.stack((tri[1] - tri[0], tri[2] - tri[0], nn))
    AI = torch.linalg.inv(A)
    ivec = AI[:, 0] * (tuv[1, 0] - tuv[0, 0]) + AI[:, 1] * (tuv[2, 0] - tuv[0, 0])
    jvec = AI[:, 0] * (tuv[1, 1] - tuv[0, 1]) + AI[:, 1] * (tuv[2, 1] - tuv[0, 1])
    B = torch.stack((normalise(ivec), normalise(jvec), nn), dim=1)
    nm = _texel(varying.uv, _t(extra.normal_map))
    normal = normalise(B @ nm)
    light_colour = _t(extra.light.colour, normal) * torch.dot(normal, normalise(_t(extra.light.direction, normal)))
    texture_colour = _texel(varying.uv, _t(extra.texture))
    colour = torch.where((light_colour >= 0).all(), texture_colour * light_colour, torch.zeros_like(light_colour))
    return PerFragment(keeps=keeps, use_default_depth=built_in.use_default_depth), varying._replace(colour=colour)


def phong_darboux_mix(gl_FragDepth: Tensor, keeps: Tensor, extra: Any):
    from .shaders.phong_darboux import PhongTextureDarbouxExtraMixerOutput

    out, picked = base_mix(gl_FragDepth, keeps, extra)
    return out, PhongTextureDarbouxExtraMixerOutput(canvas=picked.colour)


# ------------------------------------------------------------------------------------------------ phong reflection
def phong_reflection_vertex(gl_VertexID: Any, gl_InstanceID: Any, camera: Camera, extra: Any):
    """``shaders/phong_reflection.py:104-132``."""
    from .shaders.phong_reflection import PhongReflectionTextureExtraFragmentData

    return (PerVertex(gl_Position=_clip_position(camera, extra.position, gl_VertexID)),
            PhongReflectionTextureExtraFragmentData(
                normal=_eye_normal(camera, extra.normal, gl_VertexID), uv=_t(extra.uv)[gl_VertexID],
                texture_index=_t(extra.texture_index, dtype=torch.int64)[gl_VertexID], colour=torch.zeros(3, device=_t(extra.position).device)))


def phong_reflection_interpolate(values: Any, barycentric_screen: Tensor, barycentric_clip: Tensor) -> Any:
    """``shaders/phong_reflection.py:138-152``: the texture index is ``flat`` (first vertex), the rest smooth."""
    out = {}
    for name in values._fields:
        v = getattr(values, name)
        if name == "texture_index":
            out[name] = _t(v, dtype=torch.int64)[0]
        else:
            out[name] = interpolate(_t(v), barycentric_screen, barycentric_clip, Interpolation.SMOOTH)
    return type(values)(**out)


def _reflection_terms(varying: Any, extra: Any):
    """Texel, diffuse and specular terms of ``phong_reflection.py:175-220``."""
    from .model import MergedModel

    ti = varying.texture_index
    uv = MergedModel.uv_repeat(varying.uv.clone(), _t(extra.texture_shape)[ti], ti, extra.texture_offset)
    uvi = torch.floor(uv).to(torch.int64)
    texture_colour = _gather2(_t(extra.texture), uvi[0], uvi[1])
    nn = normalise(varying.normal)
    ld = normalise(_t(extra.light_dir_eye, nn))
    ndl = torch.dot(nn, ld)
    diffuse = torch.clamp_min(ndl, 0.0)
    refl = normalise(2 * ndl * nn - ld)
    specular = torch.pow(torch.clamp_min(refl[2], 0.0), _gather2(_t(extra.specular_map), uvi[0], uvi[1]))
    return texture_colour, diffuse, specular


def phong_reflection_fragment(gl_FragCoord: Tensor, gl_FrontFacing: Tensor, gl_PointCoord: Tensor, varying: Any,
                              extra: Any):
    """``shaders/phong_reflection.py:158-235``: ambient + (diffuse + specular) * light colour, all times the texel."""
    built_in = base_fragment(gl_FragCoord, gl_FrontFacing, gl_PointCoord, varying, extra)[0]
    keeps = _keep_all(built_in.keeps, gl_FrontFacing)
    tcol, diffuse, specular = _reflection_terms(varying, extra)
    amb, dif, spe = _t(extra.ambient, tcol), _t(extra.diffuse, tcol), _t(extra.specular, tcol)
    colour = amb * tcol + (dif * diffuse + spe * specular) * _t(extra.light.colour, tcol) * tcol
    return PerFragment(keeps=keeps, use_default_depth=built_in.use_default_depth), varying._replace(colour=colour)


def phong_reflection_mix(gl_FragDepth: Tensor, keeps: Tensor, extra: Any):
    from .shaders.phong_reflection import PhongReflectionTextureExtraMixerOutput

    out, picked = base_mix(gl_FragDepth, keeps, extra)
    return out, PhongReflectionTextureExtraMixerOutput(canvas=picked.colour)


# ------------------------------------------------------------------------------------------------ ... + shadow map
def phong_reflection_shadow_vertex(gl_VertexID: Any, gl_InstanceID: Any, camera: Camera, extra: Any):
    """``shaders/phong_reflection_shadow.py:115-151``: also the vertex in the light camera's normalised clip space."""
    from .geometry import normalise_homogeneous
    from .shaders.phong_reflection_shadow import PhongReflectionShadowTextureExtraFragmentData

    world = to_homogeneous(_t(extra.position)[gl_VertexID])
    return (PerVertex(gl_Position=camera.to_clip(world)),
            PhongReflectionShadowTextureExtraFragmentData(
                normal=_eye_normal(camera, extra.normal, gl_VertexID), uv=_t(extra.uv)[gl_VertexID],
                texture_index=_t(extra.texture_index, dtype=torch.int64)[gl_VertexID],
                shadow_coord=normalise_homogeneous(extra.shadow.camera.to_clip(world)), colour=torch.zeros(3, device=_t(extra.position).device)))


def phong_reflection_shadow_fragment(gl_FragCoord: Tensor, gl_FrontFacing: Tensor, gl_PointCoord: Tensor,
                                     varying: Any, extra: Any):
    """``shaders/phong_reflection_shadow.py:178-272``: the diffuse + specular part is scaled by ``1 - strength`` where
    the fragment lies behind the shadow map's depth."""
    from .geometry import normalise_homogeneous

    built_in = base_fragment(gl_FragCoord, gl_FrontFacing, gl_PointCoord, varying, extra)[0]
    keeps = _keep_all(built_in.keeps, gl_FrontFacing)
    tcol, diffuse, specular = _reflection_terms(varying, extra)
    sh = extra.shadow
    ss = normalise_homogeneous(_t(sh.camera.viewport) @ varying.shadow_coord)
    lit = ss[2] <= sh.get(ss[:2])
    strength = _t(sh.strength, tcol)
    shadow = torch.where(lit, torch.ones_like(strength), 1.0 - strength)
    amb, dif, spe = _t(extra.ambient, tcol), _t(extra.diffuse, tcol), _t(extra.specular, tcol)
    colour = amb * tcol + shadow * (dif * diffuse + spe * specular) * tcol * _t(extra.light.colour, tcol)
    return PerFragment(keeps=keeps, use_default_depth=built_in.use_default_depth), varying._replace(colour=colour)


def phong_reflection_shadow_mix(gl_FragDepth: Tensor, keeps: Tensor, extra: Any):
    from .shaders.phong_reflection_shadow import PhongReflectionShadowTextureExtraMixerOutput

    out, picked = base_mix(gl_FragDepth, keeps, extra)
    return out, PhongReflectionShadowTextureExtraMixerOutput(canvas=picked.colour)
