"""The five ``Shader`` stage methods (``vertex``, ``primitive_chooser``, ``interpolate``, ``fragment``, ``mix``) stay
callable with the reference's signatures (``renderer/shader.py:103-396``, ``renderer/shaders/*.py``).  They are host-side
tensor code (``jaxrenderer_b200/stages.py``) that ``render`` never calls; here they are COMPOSED per pixel the way the
reference's pipeline composes them (``pipeline.py:332-399``) and the result is compared with the CPU oracle's render of the
same scene, for all seven shaders.  CPU only."""
from types import SimpleNamespace as NS

import pytest
import torch

import jaxrenderer_b200 as jr
from jaxrenderer_b200.shaders import (
    DepthExtraInput, DepthShader, GouraudExtraInput, GouraudShader, GouraudTextureExtraInput, GouraudTextureShader,
    PhongReflectionShadowTextureExtraInput, PhongReflectionShadowTextureShader, PhongReflectionTextureExtraInput,
    PhongReflectionTextureShader, PhongTextureDarbouxExtraInput, PhongTextureDarbouxShader, PhongTextureExtraInput,
    PhongTextureShader,
)
from oracle import jr_oracle as O
from tests.helpers import random_mesh_scene


def _stack(trees):
    """Three per-vertex varyings -> one tree with the triangle's values on axis 0 (what ``vmap(vertex)`` yields)."""
    first = trees[0]
    if isinstance(first, tuple):
        fields = [_stack([t[i] for t in trees]) for i in range(len(first))]
        return type(first)(*fields) if hasattr(first, "_fields") else tuple(fields)
    if first is None:
        return None
    return torch.stack([torch.as_tensor(t) for t in trees])


def _lead(tree):
    if isinstance(tree, tuple):
        fields = [_lead(t) for t in tree]
        return type(tree)(*fields) if hasattr(tree, "_fields") else tuple(fields)
    return None if tree is None else torch.as_tensor(tree)[None]


def _to(tree, dev):
    if isinstance(tree, torch.Tensor):
        return tree.to(dev)
    if isinstance(tree, tuple):
        fields = [_to(t, dev) for t in tree]
        return type(tree)(*fields) if hasattr(tree, "_fields") else tuple(fields)
    return tree


def _compose(shader, cam, faces, extra, z0, c0, dev="cpu", max_pixels=None):
    """``pipeline.render`` out of the stage methods, for the triangle the visibility stage chose at each pixel.  The
    visibility stage itself (not a shader stage) is the oracle's, on the CPU; the stage methods run on ``dev``."""
    W, H = z0.shape
    pos = extra.position
    clip_v = O.mat4_apply(pos, cam.world_to_clip, w_one=True)
    setup = O.primitive_setup(clip_v, faces.long())
    idx, has, kc, _ = O.visibility(setup, cam.viewport, W, H)
    f_idx = faces.long()[idx]
    fr = _to(O.chosen_fragments(clip_v, f_idx, cam.viewport), dev)
    cam, extra, clip_v, kc = _to(cam, dev), _to(extra, dev), clip_v.to(dev), kc.to(dev)
    z, canvas = z0.clone().to(dev), (None if c0 is None else c0.clone().to(dev))
    n_checked = 0
    for x in range(W):
        for y in range(H):
            if not bool(kc[x, y]) or (max_pixels is not None and n_checked >= max_pixels):
                continue
            vids = [int(v) for v in f_idx[x, y]]
            per_vertex, varyings = zip(*[shader.vertex(v, 0, cam, extra) for v in vids])
            for k, pv in enumerate(per_vertex):                      # the vertex stage's clip position
                assert torch.allclose(pv.gl_Position, clip_v[vids[k]], rtol=0, atol=1e-6)
            bc = fr.tc[x, y]
            varying = shader.interpolate(_stack(list(varyings)), bc, bc)
            frag_coord = torch.stack((torch.tensor(float(x), device=dev), torch.tensor(float(y), device=dev), fr.zw[x, y],
                                      fr.w_rec[x, y]))
            per_frag, varying = shader.fragment(frag_coord, fr.front[x, y], torch.zeros(2, device=dev), varying, extra)
            depth = frag_coord[2] if bool(per_frag.use_default_depth) else per_frag.gl_FragDepth
            keeps = torch.as_tensor(per_frag.keeps) & kc[x, y]
            out, mixed = shader.mix(depth[None], keeps[None], _lead(varying))
            if bool(out.keep):
                z[x, y] = out.zbuffer
                if canvas is not None:
                    canvas[x, y] = mixed.canvas
            n_checked += 1
    return z.cpu(), (None if canvas is None else canvas.cpu()), n_checked


def _scene(seed):
    s = random_mesh_scene(seed, n_tri=40, W=26, H=22)
    g = torch.Generator().manual_seed(100 + seed)
    shapes = torch.tensor([[8, 6], [5, 4], [8, 3]], dtype=torch.int32)
    atlas = torch.rand(3 * 8, 6, 3, generator=g)
    spec = torch.rand(3 * 2, 2, generator=g) * 6 + 0.5
    tix = torch.randint(0, 3, (s.pos.shape[0] // 3,), generator=g).repeat_interleave(3).to(torch.int32)
    refl = dict(position=s.pos, normal=s.nrm, uv=s.uv01, light=s.light, light_dir_eye=torch.tensor((0.2, 0.3, 0.9)),
                texture_shape=shapes, texture_index=tix, texture_offset=8, texture=atlas, specular_map=spec,
                ambient=torch.tensor((0.3, 0.2, 0.1)), diffuse=torch.tensor((0.5, 0.6, 0.7)),
                specular=torch.tensor((0.2, 0.3, 0.4)))
    return s, refl


def _cases(s, refl):
    n_tri = s.faces.shape[0]
    i2f = torch.arange(n_tri, dtype=torch.int32).repeat_interleave(3)
    yield "depth", DepthShader, DepthExtraInput(s.pos), None
    yield "gouraud", GouraudShader, GouraudExtraInput(s.pos, s.col, s.nrm, s.light), None
    yield ("gouraud_texture", GouraudTextureShader,
           GouraudTextureExtraInput(s.pos, s.nrm, s.uv_texel, s.light, s.texture), None)
    yield "phong", PhongTextureShader, PhongTextureExtraInput(s.pos, s.nrm, s.uv_texel, s.light, s.texture), None
    yield ("phong_darboux", PhongTextureDarbouxShader,
           PhongTextureDarbouxExtraInput(s.pos, s.nrm, s.uv_texel, s.light, s.texture, s.normal_map, i2f, s.faces), None)
    yield "phong_reflection", PhongReflectionTextureShader, PhongReflectionTextureExtraInput(**refl), None
    # the shadow pass on the CPU: the oracle's light camera and shadow map inside the product's `Shadow` carrier
    ocam = O.shadow_camera(torch.tensor((0.4, 0.3, 0.9)), s.cam.viewport, torch.zeros(3), torch.tensor((0.0, 0.0, 1.0)))
    sm = O.render_shadow_map(torch.full((s.W, s.H), torch.finfo(torch.float32).max), s.pos, s.faces, ocam, 0.05)
    eye = torch.eye(4)
    light_cam = jr.Camera(view=ocam.view, projection=ocam.projection, viewport=ocam.viewport,
                          world_to_clip=ocam.world_to_clip, world_to_eye_norm=ocam.world_to_eye_norm,
                          view_inv=eye, screen_to_world=eye, world_to_screen=eye)
    shadow = jr.Shadow(shadow_map=sm, strength=torch.tensor((0.6, 0.5, 0.4)), camera=light_cam)
    oshadow = NS(shadow_map=sm, strength=shadow.strength, camera=ocam)
    yield ("phong_reflection_shadow", PhongReflectionShadowTextureShader,
           PhongReflectionShadowTextureExtraInput(**refl, shadow=shadow, camera=s.cam),
           NS(**refl, shadow=oshadow, camera=s.cam))


@pytest.mark.parametrize("seed", [0, 5])
def test_composed_stage_methods_reproduce_the_oracle_render(seed):
    s, refl = _scene(seed)
    for name, shader, extra, oracle_extra in _cases(s, refl):
        z0 = torch.full((s.W, s.H), 1.0)
        c0 = None if name == "depth" else torch.full((s.W, s.H, 3), 0.25)
        z, c, n = _compose(shader, s.cam, s.faces, extra, z0, c0)
        ref = O.render(s.cam, name, z0, () if c0 is None else (c0,), s.faces, oracle_extra or extra)
        assert n > 20, (name, n)
        assert torch.allclose(z, ref.zbuffer, rtol=0, atol=1e-6), name
        if c0 is not None:
            err = float((c - ref.targets[0]).abs().max())
            assert err <= 2e-6, (name, err)
            assert int((c != c0).any(-1).sum()) > 5, name


def test_primitive_chooser_and_mix_semantics():
    """``shader.py:207-217``, ``:383-396``: closest kept front-facing primitive, first index on ties, index 0 (not kept)
    when there is none; one primitive comes back, on a leading axis of length 1."""
    depth = torch.tensor([0.7, 0.3, 0.3, 0.1, 0.5])
    coord = torch.stack((torch.zeros(5), torch.zeros(5), depth, torch.ones(5)), dim=1)
    front = torch.tensor([True, True, True, False, True])
    keeps = torch.tensor([True, True, True, True, False])
    values = (torch.arange(5.0)[:, None].repeat(1, 3), torch.arange(10.0).reshape(5, 2))
    bary = torch.rand(5, 3)
    out = jr.Shader.primitive_chooser(coord, front, torch.zeros(5, 2), keeps, values, bary, bary)
    assert out[0].shape == (1, 4) and float(out[0][0, 2]) == pytest.approx(0.3)
    assert torch.equal(out[4][0], values[0][1:2]) and torch.equal(out[5], bary[1:2])     # first of the two ties
    none = jr.Shader.primitive_chooser(coord, front & False, torch.zeros(5, 2), keeps, values, bary, bary)
    assert torch.equal(none[4][1], values[1][0:1]) and not bool(none[1][0])
    mo, picked = jr.Shader.mix(depth, keeps, values)
    assert bool(mo.keep) and float(mo.zbuffer) == pytest.approx(0.1) and torch.equal(picked[0], values[0][3])
    mo, _ = jr.Shader.mix(depth, keeps & False, values)
    assert not bool(mo.keep) and float(mo.zbuffer) == float("inf")
    with pytest.raises(NotImplementedError):
        jr.Shader.vertex(0, 0, None, None)


@pytest.mark.gpu
def test_stage_methods_are_device_agnostic():
    """The same composition with every tensor on the GPU (a handful of pixels: the stage methods are per-element host
    code, not a rendering path)."""
    s, refl = _scene(0)
    for name, shader, extra, oracle_extra in _cases(s, refl):
        if name not in ("gouraud_texture", "phong_darboux", "phong_reflection_shadow"):
            continue
        z0, c0 = torch.full((s.W, s.H), 1.0), torch.full((s.W, s.H, 3), 0.25)
        z, c, n = _compose(shader, s.cam, s.faces, extra, z0, c0, dev="cuda", max_pixels=12)
        ref = O.render(s.cam, name, z0, (c0,), s.faces, oracle_extra or extra)
        written = (c != c0).any(-1)
        assert n == 12 and int(written.sum()) > 0
        assert float((c - ref.targets[0])[written].abs().max()) <= 1e-5, name
