"""GPU parity tests (forward): CUDA path through the C ABI vs the CPU oracle.

Tolerances are BASELINE.json's: triangle-id buffer bit-exact except pixels
whose competing depths differ by < 1e-6 (counted in the report), z bit-exact
on agreeing pixels, colours within 1e-5 relative.
"""
from types import SimpleNamespace as NS

import pytest
import torch

import jaxrenderer_b200 as jr
from jaxrenderer_b200 import synthetic
from jaxrenderer_b200.shaders import (
    DepthExtraInput, DepthShader, GouraudExtraInput, GouraudShader, GouraudTextureExtraInput,
    GouraudTextureShader, PhongReflectionShadowTextureExtraInput, PhongReflectionShadowTextureShader,
    PhongReflectionTextureExtraInput, PhongReflectionTextureShader, PhongTextureDarbouxExtraInput,
    PhongTextureDarbouxShader, PhongTextureExtraInput, PhongTextureShader,
)
from oracle import jr_oracle as O
from tests.helpers import cam_at, assert_parity, compare, load_brax_fixture, random_mesh_scene, smoke_scene

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cuda(x):
    if isinstance(x, torch.Tensor):
        return x.to(DEV)
    if isinstance(x, tuple) and hasattr(x, "_fields"):
        return type(x)(*[_cuda(v) for v in x])
    return x


def _run(cam, shader, z0, c0, faces, extra):
    bufs = jr.Buffers(zbuffer=z0.to(DEV), targets=() if c0 is None else (c0.to(DEV),))
    out, tri = jr.render(_cuda(cam), shader, bufs, faces.to(DEV), _cuda(extra), return_tri_id=True)
    torch.cuda.synchronize()
    return out.zbuffer, (out.targets[0] if out.targets else None), tri


def test_reference_smoke_scene_1920x1080():
    """reference tests/smoke_test.py:27-132, every assertion + oracle parity (tiled path)."""
    W, H = 1920, 1080
    cam, faces, extra = smoke_scene(W, H)
    z0, c0 = torch.full((W, H), 1.0), torch.zeros(W, H, 3)
    z, c, tri = _run(cam, GouraudShader, z0, c0, faces, extra)
    zd, cd = jr.transpose_for_display(z.cpu()), jr.transpose_for_display(c.cpu())
    assert zd.shape == (H, W) and cd.shape == (H, W, 3)
    assert torch.unique(zd[293:528, 964:1423].to(torch.uint8)).shape == (1,)
    assert bool((zd[590:1049, 964:1423] == 1.0).all())
    assert zd[551, 914] < zd[1026, 92]
    empty = int((cd == 0).all(dim=2).sum())
    assert W * H // 2 < empty < W * H
    ref = O.render(cam, "gouraud", z0, (c0,), faces, extra)
    rep = compare("smoke1920", z, c, tri, ref)
    print(rep)
    assert_parity(rep)


@pytest.mark.parametrize("wh", [(84, 84), (32, 32), (200, 120)])
def test_depth_brax_like(wh):
    """BASELINE config 2 shape (depth shader, ant-like scenes), small batch vs oracle."""
    W, H = wh
    B = 3
    sc = synthetic.brax_like_batch(B, n_capsules=10)
    cam = synthetic.brax_cameras(sc["eye"], sc["target"], W, H)
    z0 = torch.full((B, W, H), 1.0)
    bufs = jr.Buffers(zbuffer=z0.to(DEV), targets=())
    out, tri = jr.render(_cuda(cam), DepthShader, bufs, sc["faces"].to(DEV),
                         DepthExtraInput(position=sc["position"].to(DEV)), return_tri_id=True)
    for b in range(B):
        camb = cam_at(cam, b)
        ref = O.render(camb, "depth", z0[b], (), sc["faces"][b], NS(position=sc["position"][b]))
        rep = compare(f"depth{W}x{H}[{b}]", out.zbuffer[b], None, tri[b], ref)
        print(rep)
        assert_parity(rep)
        assert int((ref.tri_id >= 0).sum()) > W * H // 2  # ground plane covers most of the view


@pytest.mark.parametrize("wh", [(40, 36), (300, 200)])
def test_depth_triangle0_backfacing_leak(wh):
    """a kept back-facing triangle 0 leaks into the depth buffer where
    no candidate exists (DepthShader has no front-face term).  Single-tile and binned paths."""
    W, H = wh
    cam, _, extra = smoke_scene(W, H, depth=1.0)
    pos = torch.cat((extra.position, torch.tensor(((-2.0, -1.0, 0.0), (-1.0, -1.0, 0.0), (-1.5, -0.2, 0.3)))))
    faces = torch.tensor(((0, 2, 1), (6, 7, 8)), dtype=torch.int32)  # tri 0 = flipped (0,1,2)
    z0 = torch.full((W, H), 7.0)
    z, _, tri = _run(cam, DepthShader, z0, None, faces, DepthExtraInput(position=pos))
    ref = O.render(cam, "depth", z0, (), faces, NS(position=pos))
    assert int(ref.has.sum()) > 0, "scene must also contain a regular front-facing triangle"
    assert int(((ref.tri_id == 0) & ~ref.has).sum()) > 0, "scene must exercise the leak"
    rep = compare("tri0leak", z, None, tri, ref)
    print(rep)
    assert_parity(rep)


def _shader_cases(s):
    light = s.light
    yield "gouraud", GouraudShader, GouraudExtraInput(s.pos, s.col, s.nrm, light)
    yield "gouraud_texture", GouraudTextureShader, GouraudTextureExtraInput(s.pos, s.nrm, s.uv_texel, light, s.texture)
    yield "phong", PhongTextureShader, PhongTextureExtraInput(s.pos, s.nrm, s.uv_texel, light, s.texture)
    n_tri = s.faces.shape[0]
    id_to_face = torch.arange(n_tri, dtype=torch.int32).repeat_interleave(3)
    yield "phong_darboux", PhongTextureDarbouxShader, PhongTextureDarbouxExtraInput(
        s.pos, s.nrm, s.uv_texel, light, s.texture, s.normal_map, id_to_face, s.faces)


@pytest.mark.parametrize("seed", [0, 1])
def test_simple_shaders_random_soup(seed):
    s = random_mesh_scene(seed)
    z0, c0 = torch.full((s.W, s.H), 1.0), torch.full((s.W, s.H, 3), 0.25)
    for name, shader, extra in _shader_cases(s):
        z, c, tri = _run(s.cam, shader, z0, c0, s.faces, extra)
        ref = O.render(s.cam, name, z0, (c0,), s.faces, extra)
        rep = compare(name, z, c, tri, ref)
        print(rep)
        assert_parity(rep)
        assert rep["pixels"] - int((ref.tri_id < 0).sum()) > 50


@pytest.mark.parametrize("wh", [(84, 84), (50, 37)])
def test_clustered_small_batch_equals_one_cta_per_image(wh):
    """Depth passes with z-only keys of batches that leave SMs idle split every image over a 2-CTA thread-block cluster
    (`k_vis3<..., CLUSTER>`: two key tiles merged through distributed shared memory); larger batches run one CTA per
    image.  The same images rendered both ways must be bit-identical (vector resolve at 84x84, scalar resolve at 50x37)."""
    W, H = wh
    B = 640                                   # 2 * 640 CTAs > 4 per SM on 148 SMs: one CTA per image
    sc = synthetic.brax_like_batch(B, n_capsules=3, env0=901)
    cam = synthetic.brax_cameras(sc["eye"], sc["target"], W, H)
    camd = _cuda(cam)
    pos, faces = sc["position"].to(DEV), sc["faces"].to(DEV)
    z_all = jr.render(camd, DepthShader, jr.Buffers(torch.full((B, W, H), 1.0, device=DEV), ()), faces,
                      DepthExtraInput(position=pos)).zbuffer
    n = 12                                    # clustered
    cam_n = type(camd)(*[(v[:n] if v.ndim == 3 and v.shape[0] == B else v) for v in camd])
    z_few = jr.render(cam_n, DepthShader, jr.Buffers(torch.full((n, W, H), 1.0, device=DEV), ()), faces[:n],
                      DepthExtraInput(position=pos[:n])).zbuffer
    assert int((z_few != 1.0).sum()) > 0.2 * z_few.numel()
    assert torch.equal(z_few, z_all[:n])


def test_visible_triangle_lists_fused_resolve_equals_separate_scan():
    """Non-depth shaders on single-tile canvases: the visibility kernel's resolve builds the visible-triangle lists
    itself (V3Vis, jr_vis3.cuh) while meshes of more than 49 152 triangles -- beyond its shared-memory flag array --
    still go through `k_mark_visible`.  Padding the index buffer with degenerate triangles switches between the two
    without changing the image: both must agree bit for bit (z, colours, triangle ids), compact records (T > W*H)
    and slot-per-triangle records alike."""
    for (n_tri, W, H) in ((60, 48, 40), (3000, 48, 40)):
        s = random_mesh_scene(7, n_tri=n_tri, W=W, H=H)
        z0, c0 = torch.full((W, H), 1.0), torch.full((W, H, 3), 0.25)
        pad = torch.zeros((49160 - n_tri, 3), dtype=torch.int32)          # degenerate: never rasterised
        for name, shader, extra in _shader_cases(s):
            if name == "phong_darboux":
                continue                                                    # no attribute records for this shader
            z, c, tri = _run(s.cam, shader, z0, c0, s.faces, extra)
            zp, cp, trip = _run(s.cam, shader, z0, c0, torch.cat((s.faces, pad)), extra)
            assert int((tri >= 0).sum()) > 50
            assert torch.equal(z, zp) and torch.equal(c, cp) and torch.equal(tri, trip), (name, n_tri)


def _atlas_inputs(s, n_obj=3):
    g = s.gen
    tw, th = 8, 6
    shapes = torch.tensor([[8, 6], [5, 4], [8, 3]], dtype=torch.int32)
    atlas = torch.rand(n_obj * tw, th, 3, generator=g)
    spec = torch.rand(n_obj * 2, 2, generator=g) * 6 + 0.5
    tix = torch.randint(0, n_obj, (s.pos.shape[0] // 3,), generator=g).repeat_interleave(3).to(torch.int32)
    return shapes, atlas, spec, tix, tw


@pytest.mark.parametrize("seed", [0, 3])
def test_phong_reflection_and_shadow_random_soup(seed):
    s = random_mesh_scene(seed, n_tri=80)
    shapes, atlas, spec, tix, off = _atlas_inputs(s)
    lde = torch.tensor((0.2, 0.3, 0.9))
    base = dict(position=s.pos, normal=s.nrm, uv=s.uv01, light=s.light, light_dir_eye=lde,
                texture_shape=shapes, texture_index=tix, texture_offset=off, texture=atlas,
                specular_map=spec, ambient=torch.tensor((0.3, 0.2, 0.1)),
                diffuse=torch.tensor((0.5, 0.6, 0.7)), specular=torch.tensor((0.2, 0.3, 0.4)))
    z0, c0 = torch.full((s.W, s.H), 1.0), torch.full((s.W, s.H, 3), 0.25)
    extra = PhongReflectionTextureExtraInput(**base)
    z, c, tri = _run(s.cam, PhongReflectionTextureShader, z0, c0, s.faces, extra)
    ref = O.render(s.cam, "phong_reflection", z0, (c0,), s.faces, extra)
    rep = compare("phong_reflection", z, c, tri, ref)
    print(rep)
    assert_parity(rep)
    # shadow pass through the product API, then S7
    sm0 = torch.full((s.W, s.H), torch.finfo(torch.float32).max)
    shadow = jr.Shadow.render_shadow_map(sm0.to(DEV), s.pos.to(DEV), s.faces.to(DEV),
                                         torch.tensor((0.4, 0.3, 0.9)), s.cam.viewport.to(DEV),
                                         torch.zeros(3), torch.tensor((0.0, 0.0, 1.0)),
                                         torch.tensor((0.6, 0.5, 0.4)), offset=0.05)
    scam = NS(world_to_clip=shadow.camera.world_to_clip.cpu(), viewport=shadow.camera.viewport.cpu())
    ref_sm = O.render_shadow_map(sm0, s.pos, s.faces, scam, 0.05)
    assert torch.equal(shadow.shadow_map.cpu(), ref_sm), "shadow map must be bit-equal"
    assert int((ref_sm < 1e30).sum()) > 20
    extra7 = PhongReflectionShadowTextureExtraInput(**base, shadow=shadow, camera=s.cam)
    z, c, tri = _run(s.cam, PhongReflectionShadowTextureShader, z0, c0, s.faces, extra7)
    oshadow = NS(shadow_map=ref_sm, strength=shadow.strength, camera=scam)
    ref = O.render(s.cam, "phong_reflection_shadow", z0, (c0,), s.faces,
                   NS(**base, shadow=oshadow, camera=s.cam))
    rep = compare("phong_reflection_shadow", z, c, tri, ref)
    print(rep)
    assert_parity(rep)


def _cube_objects():
    cube = jr.create_cube(torch.ones(3), torch.ones(2),
                          torch.zeros(2, 2, 3).index_fill_(2, torch.tensor([2]), 1.0), torch.ones(2, 2) * 2.0)
    return [jr.ModelObject(model=cube)]


@pytest.mark.parametrize("shadow", [False, True])
def test_simple_cube_example(shadow):
    """BASELINE config 1: examples/simple_cube.py (640x480), default and shadow modes,
    plus the analytic check of SURVEY 8c(3)."""
    W, H = 640, 480
    objs = _cube_objects()
    camp = jr.CameraParameters(viewWidth=W, viewHeight=H, position=torch.tensor([2.0, 4.0, 1.0]))
    light = jr.LightParameters()
    sp = jr.ShadowParameters() if shadow else None
    objs_d = [o._replace(model=_cuda(o.model)) for o in objs]
    # host-side matrices are computed once on the CPU and shared with the oracle
    model = jr.merge_objects(objs)
    cam = jr.Renderer.create_camera_from_parameters(camp)
    img = jr.Renderer.get_camera_image(objs_d, light, _cuda(cam), W, H, shadow_param=sp)
    assert img.shape == (W, H, 3) and img.is_cuda
    img_p = jr.Renderer.get_camera_image(objs_d, light, camp, W, H, shadow_param=sp)
    assert float((img_p - img).abs().max()) < 1e-3      # CameraParameters route, device-built camera
    scam = None
    if shadow:
        sh = jr.Shadow.render_shadow_map(
            torch.full((W, H), torch.finfo(torch.float32).max, device=DEV), model.verts.to(DEV),
            model.faces.to(DEV), torch.tensor(light.direction), cam.viewport.to(DEV), sp.centre, sp.up,
            sp.strength, offset=sp.offset)
        scam = NS(world_to_clip=sh.camera.world_to_clip.cpu(), viewport=sh.camera.viewport.cpu())
    lightp = NS(**{k: torch.tensor(v) for k, v in light._asdict().items()})
    spo = NS(centre=torch.tensor(sp.centre), up=torch.tensor(sp.up), strength=torch.tensor(sp.strength),
             offset=sp.offset) if shadow else None
    res = O.renderer_render(model, lightp, cam, torch.ones(W, H), torch.ones(W, H, 3), spo, scam)
    ref = res["out"]
    ok = torch.ones(W, H, dtype=torch.bool)
    err = ((img.cpu() - ref.targets[0]).abs() / ref.targets[0].abs().clamp_min(1e-3))
    print("simple_cube shadow=%s max rel err %.3g, covered %d" % (shadow, float(err.max()), int((ref.tri_id >= 0).sum())))
    assert float(err.max()) <= 1e-5
    covered = ref.tri_id >= 0
    assert 10000 < int(covered.sum()) < W * H // 2
    px = img.cpu()[covered]
    assert float(px[:, :2].abs().max()) == 0.0          # pure-blue texture: R = G = 0
    assert bool((img.cpu()[~covered] == 1.0).all())      # background


def test_brax_fixture_frame_with_shadow_84():
    """Real Brax ant fixture (3276 triangles), full view 84x84, shadow pass on."""
    W, H = 84, 84
    objs, camp = load_brax_fixture()
    f = 1
    objs1 = [o._replace(local_scaling=o.local_scaling[f], transform=o.transform[f]) for o in objs]
    camp1 = jr.CameraParameters(**{k: v[f] for k, v in camp._asdict().items()})._replace(
        viewWidth=W, viewHeight=H, vfov=58.0 * H / W)
    light = jr.LightParameters(direction=(0.57735, -0.57735, 0.57735), ambient=(0.8,) * 3,
                               diffuse=(0.8,) * 3, specular=(0.6,) * 3)
    sp = jr.ShadowParameters(centre=camp1.target)
    model = jr.merge_objects(objs1)
    assert model.faces.shape[0] == 3276
    cam = jr.Renderer.create_camera_from_parameters(camp1)
    bufs = jr.Renderer.create_buffers(W, H, device=DEV)
    img = jr.Renderer.render(_cuda(model), light, _cuda(cam), bufs, shadow_param=sp).targets[0]
    sh = jr.Shadow.render_shadow_map(
        torch.full((W, H), torch.finfo(torch.float32).max, device=DEV), model.verts.to(DEV),
        model.faces.to(DEV), torch.tensor(light.direction), cam.viewport.to(DEV), sp.centre, sp.up,
        sp.strength, offset=sp.offset)
    scam = NS(world_to_clip=sh.camera.world_to_clip.cpu(), viewport=sh.camera.viewport.cpu())
    lightp = NS(**{k: torch.tensor(v) for k, v in light._asdict().items()})
    spo = NS(centre=sp.centre, up=torch.tensor(sp.up), strength=torch.tensor(sp.strength), offset=sp.offset)
    res = O.renderer_render(model, lightp, cam, torch.ones(W, H), torch.ones(W, H, 3), spo, scam)
    assert torch.equal(sh.shadow_map.cpu(), res["shadow_map"])
    ref = res["out"]
    err = ((img.cpu() - ref.targets[0]).abs() / ref.targets[0].abs().clamp_min(1e-3))
    bad = int((err > 1e-5).sum())
    print("brax frame: max rel err %.3g, pixels over tol %d, shadowed texels %d"
          % (float(err.max()), bad, int((res["shadow_map"] < 1e30).sum())))
    assert bad == 0


def test_batch_broadcast_and_inplace():
    """vmap semantics: batched positions, shared camera/faces; donated buffers."""
    W, H, B = 32, 24, 5
    sc = synthetic.brax_like_batch(B, n_capsules=2)
    cam = synthetic.brax_cameras(sc["eye"][0], sc["target"][0], W, H)          # un-batched camera
    z0 = torch.full((B, W, H), 1.0, device=DEV)
    out = jr.render(_cuda(cam), DepthShader, jr.Buffers(z0, ()), sc["faces"][0].to(DEV),
                    DepthExtraInput(position=sc["position"].to(DEV)), inplace=True)
    assert out.zbuffer.data_ptr() == z0.data_ptr()
    for b in (0, B - 1):
        one = jr.render(_cuda(cam), DepthShader, jr.Buffers(torch.full((W, H), 1.0, device=DEV), ()),
                        sc["faces"][0].to(DEV), DepthExtraInput(position=sc["position"][b].to(DEV)))
        assert one.zbuffer.shape == (W, H)
        assert torch.equal(one.zbuffer, out.zbuffer[b])
    # host tensors in -> host tensors out (implicit device_put)
    host = jr.render(cam, DepthShader, jr.Buffers(torch.full((W, H), 1.0), ()), sc["faces"][0],
                     DepthExtraInput(position=sc["position"][0]))
    assert not host.zbuffer.is_cuda and torch.equal(host.zbuffer, out.zbuffer[0].cpu())


def test_uint8_display_epilogue():
    W, H, B = 37, 21, 3
    c = torch.rand(B, W, H, 3, device=DEV) * 1.4 - 0.2
    got = jr.canvas_to_uint8_display(c)
    want = (c.clamp(0, 1) * 255).to(torch.uint8).transpose(1, 2).flip(1)
    assert torch.equal(got, want)


def test_fused_merge_objects_matches_host_merge():
    """jr_merge_objects (CUDA) vs the torch host implementation of merge_objects on the real Brax
    fixture (18 objects, batched transforms), and the faces / atlas bookkeeping."""
    objs, _ = load_brax_fixture()
    ref = jr.merge_objects(objs)                                  # CPU: torch ops
    objs_d = [jr.ModelObject(model=_cuda(o.model), local_scaling=o.local_scaling.to(DEV),
                             transform=o.transform.to(DEV), double_sided=o.double_sided) for o in objs]
    got = jr.merge_objects(objs_d)                                # CUDA: fused kernels
    assert got.verts.shape == ref.verts.shape == (4, 9816, 3)
    assert torch.equal(got.faces.cpu(), ref.faces) and torch.equal(got.faces_norm.cpu(), ref.faces_norm)
    assert torch.equal(got.faces_uv.cpu(), ref.faces_uv) and torch.equal(got.texture_index.cpu(), ref.texture_index)
    scale = ref.verts.abs().max()
    assert float((got.verts.cpu() - ref.verts).abs().max()) <= 2e-6 * float(scale)
    assert float((got.norms.cpu() - ref.norms).abs().max()) <= 1e-6
    # un-batched objects
    one = [o._replace(local_scaling=o.local_scaling[1], transform=o.transform[1]) for o in objs_d]
    g1 = jr.merge_objects(one)
    assert g1.verts.shape == (9816, 3)
    assert float((g1.verts - got.verts[1]).abs().max()) == 0.0 and float((g1.norms - got.norms[1]).abs().max()) == 0.0


def test_instanced_geometry_equals_merged_path_bit_for_bit():
    """SURVEY 8f-1 in full: `merge_objects` on CUDA hands the render kernels the FACTORED geometry (shared local
    meshes + per-image object transforms); they instance it on the fly.  Results must equal the materialised
    (merged world-space arrays) path bit for bit: depth via `pipeline.render`, phong_reflection and
    phong_reflection_shadow (with its shadow pass) via `Renderer.render`, on the real Brax ant fixture (18 objects,
    batched transforms), single-tile (84x84) and binned (200x150) canvases."""
    from jaxrenderer_b200 import _native
    from jaxrenderer_b200.model import InstancedArray

    objs, camp = load_brax_fixture()
    objs_d = [jr.ModelObject(model=_cuda(o.model), local_scaling=o.local_scaling.to(DEV),
                             transform=o.transform.to(DEV), double_sided=o.double_sided) for o in objs]
    B = objs[0].transform.shape[0]
    light = jr.LightParameters(direction=(0.57735, -0.57735, 0.57735), ambient=(0.8,) * 3,
                               diffuse=(0.8,) * 3, specular=(0.6,) * 3)
    for (W, H) in ((84, 84), (200, 150)):
        cp = jr.CameraParameters(**{k: v.to(DEV) for k, v in camp._asdict().items()})._replace(
            viewWidth=W, viewHeight=H, vfov=58.0 * H / W)
        cam = jr.Renderer.create_camera_from_parameters(cp)
        sp = jr.ShadowParameters(centre=cp.target)
        lazy = jr.merge_objects(objs_d)
        assert isinstance(lazy.verts, InstancedArray) and isinstance(lazy.norms, InstancedArray)
        assert lazy.verts.shape == (B, 9816, 3)
        merged = lazy._replace(verts=lazy.verts.materialise(), norms=lazy.norms.materialise())
        n0 = _native.launch_count()
        for shadow in (None, sp):
            a = jr.Renderer.render(lazy, light, cam, jr.Renderer.create_buffers(W, H, batch=B, device=DEV), shadow_param=shadow)
            m = jr.Renderer.render(merged, light, cam, jr.Renderer.create_buffers(W, H, batch=B, device=DEV), shadow_param=shadow)
            assert torch.equal(a.zbuffer, m.zbuffer), (W, H, shadow is not None)
            assert torch.equal(a.targets[0], m.targets[0]), (W, H, shadow is not None)
            assert int((a.zbuffer != 1.0).sum()) > 0.5 * a.zbuffer.numel()
        # depth shader through the generic boundary, lazy positions handed over as they are
        za, ta = jr.render(cam, DepthShader, jr.Buffers(torch.full((B, W, H), 1.0, device=DEV), ()), lazy.faces,
                           DepthExtraInput(position=lazy.verts), return_tri_id=True)
        zm, tm = jr.render(cam, DepthShader, jr.Buffers(torch.full((B, W, H), 1.0, device=DEV), ()), merged.faces,
                           DepthExtraInput(position=merged.verts), return_tri_id=True)
        assert torch.equal(za.zbuffer, zm.zbuffer) and torch.equal(ta, tm)
        # z-only-key variant
        zk = jr.render(cam, DepthShader, jr.Buffers(torch.full((B, W, H), 1.0, device=DEV), ()), lazy.faces,
                       DepthExtraInput(position=lazy.verts))
        assert torch.equal(zk.zbuffer, zm.zbuffer)
    # gradients requested -> the materialised path is used transparently
    atlas = lazy.diffuse_map.clone().requires_grad_(True)
    out = jr.Renderer.render(lazy._replace(diffuse_map=atlas), light, cam,
                             jr.Renderer.create_buffers(W, H, batch=B, device=DEV))
    out.targets[0].sum().backward()
    assert atlas.grad is not None and float(atlas.grad.abs().sum()) > 0


@pytest.mark.parametrize("wh", [(84, 84), (50, 37), (200, 150)])
def test_shadow_pass_in_one_launch_equals_fill_render_add(wh):
    """The depth epilogue (JrRenderArgs.depth_offset / depth_fill): `Renderer.render`'s shadow map -- fill with the
    largest float, depth render from the light, `+ offset` (renderer.py:349-354, shadow.py:106-116) -- in ONE launch,
    bit-equal to the three-pass form, on single-tile (vector and scalar resolve) and binned canvases."""
    from jaxrenderer_b200 import _native
    from jaxrenderer_b200.shadow import ConstantFill

    W, H = wh
    B = 3
    sc = synthetic.brax_like_batch(B, n_capsules=3, env0=77)
    cam = synthetic.brax_cameras(sc["eye"], sc["target"], W, H)
    fill = torch.finfo(torch.float32).max
    args = (sc["position"].to(DEV), sc["faces"].to(DEV), torch.tensor((0.57735, -0.57735, 0.57735)),
            cam.viewport.to(DEV), sc["target"].to(DEV), (0.0, 0.0, 1.0), (0.6, 0.6, 0.6))
    three = jr.Shadow.render_shadow_map(torch.full((B, W, H), fill, device=DEV), *args, offset=0.05)
    n0 = _native.launch_count()
    one = jr.Shadow.render_shadow_map(ConstantFill((B, W, H), fill, torch.device(DEV)), *args, offset=0.05)
    n1 = _native.launch_count()
    assert torch.equal(one.shadow_map, three.shadow_map)
    assert int((one.shadow_map < 1e30).sum()) > 0 and bool((one.shadow_map[one.shadow_map > 1e30] == fill).all())
    # light camera (1) + depth kernel(s): no fill pass, no add pass
    assert n1 - n0 <= (2 if W * H * 8 <= 96 * 1024 else 3), n1 - n0


def test_torch_func_vmap_is_the_native_batch():
    """Batch rendering "via vmap" (BASELINE north_star; examples/batch_rendering.py:87-95:
    `jax.vmap(lambda m, b: Renderer.render(m, light, camera, b))`): `torch.func.vmap` around `Renderer.render` /
    `pipeline.render` gives exactly the natively batched result (mapped inputs -> batched arrays, un-mapped ones ->
    shared), and a gradient taken outside the vmap equals the native one."""
    from torch.func import vmap

    W, H, B = 84, 84, 5
    sc = synthetic.brax_like_batch(B, n_capsules=3, env0=555, with_attributes=True)
    cam1 = synthetic.brax_cameras(sc["eye"][0], sc["target"][0], W, H)          # one shared camera
    cam = _cuda(cam1)
    model = synthetic.merged_model_from_batch(sc, 3, DEV)                        # verts / norms batched, rest shared
    light = jr.LightParameters(direction=(0.57735, -0.57735, 0.57735), ambient=(0.8,) * 3, diffuse=(0.8,) * 3,
                               specular=(0.6,) * 3)
    sp = jr.ShadowParameters(centre=sc["target"][0].to(DEV))
    native = jr.Renderer.render(model, light, cam, jr.Renderer.create_buffers(W, H, batch=B, device=DEV), shadow_param=sp)
    one = jr.Renderer.create_buffers(W, H, device=DEV)

    def per_image(v, n, z, c):
        out = jr.Renderer.render(model._replace(verts=v, norms=n), light, cam, jr.Buffers(z, (c,)), shadow_param=sp)
        return out.zbuffer, out.targets[0]
    # model mapped, buffers shared (in_axes None) ...
    z1, c1 = vmap(per_image, in_dims=(0, 0, None, None))(model.verts, model.norms, one.zbuffer, one.targets[0])
    assert torch.equal(z1, native.zbuffer) and torch.equal(c1, native.targets[0])
    # ... and both mapped, as in the reference's example
    bufs = jr.Renderer.create_buffers(W, H, batch=B, device=DEV)
    z2, c2 = vmap(per_image)(model.verts, model.norms, bufs.zbuffer, bufs.targets[0])
    assert torch.equal(z2, native.zbuffer) and torch.equal(c2, native.targets[0])
    # the generic boundary, position mapped along axis 1
    pos_t = model.verts.transpose(0, 1).contiguous()                              # (Nv, B, 3)
    zd = vmap(lambda p: jr.render(cam, DepthShader, jr.Buffers(torch.full((W, H), 1.0, device=DEV), ()), model.faces,
                                  DepthExtraInput(position=p)).zbuffer, in_dims=1)(pos_t)
    want = jr.render(cam, DepthShader, jr.Buffers(torch.full((B, W, H), 1.0, device=DEV), ()), model.faces,
                     DepthExtraInput(position=model.verts)).zbuffer
    assert torch.equal(zd, want)
    # gradient outside the vmap (shared atlas)
    atlas = model.diffuse_map.clone().requires_grad_(True)
    cv = vmap(lambda v, n: jr.Renderer.render(model._replace(verts=v, norms=n, diffuse_map=atlas), light, cam,
                                              jr.Buffers(one.zbuffer, one.targets)).targets[0])(model.verts, model.norms)
    cv.sum().backward()
    atlas_n = model.diffuse_map.clone().requires_grad_(True)
    jr.Renderer.render(model._replace(diffuse_map=atlas_n), light, cam,
                       jr.Renderer.create_buffers(W, H, batch=B, device=DEV)).targets[0].sum().backward()
    assert torch.equal(atlas.grad, atlas_n.grad)
    # nested batching is refused with a clear message
    with pytest.raises(NotImplementedError, match="nested"):
        vmap(lambda z: jr.render(cam, DepthShader, jr.Buffers(z, ()), model.faces,
                                 DepthExtraInput(position=model.verts)).zbuffer)(bufs.zbuffer)


@pytest.mark.parametrize("wh", [(84, 84), (50, 37), (200, 150)])
def test_display_uint8_fused_into_the_shading_store(wh):
    """`get_camera_image(..., display_uint8=True)` -- the uint8 / transposed / flipped display image
    written by the shading kernel itself -- equals the fp32 render followed by the reference's epilogue
    (`transpose_for_display((clip(c, 0, 1) * 255).astype(uint8))`, utils.py:79-98), with and without the shadow
    pass, constant background and incoming canvas."""
    W, H = wh
    B = 3
    objs, eye, tgt = synthetic.brax_like_objects(B, n_capsules=3, env0=99, device=DEV)
    cam = _cuda(synthetic.brax_cameras(eye, tgt, W, H))
    light = jr.LightParameters(direction=(0.57735, -0.57735, 0.57735), ambient=(0.8,) * 3, diffuse=(0.9,) * 3,
                               specular=(0.9,) * 3)          # bright enough to exercise the clamp
    for sp in (None, jr.ShadowParameters(centre=tgt.to(DEV))):
        ref = jr.Renderer.get_camera_image(objs, light, cam, W, H, colour_default=(0.2, 0.5, 1.3), shadow_param=sp)
        want = (ref.clamp(0, 1) * 255).to(torch.uint8).transpose(1, 2).flip(1)
        got = jr.Renderer.get_camera_image(objs, light, cam, W, H, colour_default=(0.2, 0.5, 1.3), shadow_param=sp,
                                           display_uint8=True)
        assert got.dtype == torch.uint8 and got.shape == (B, H, W, 3)
        assert torch.equal(got, want), (wh, sp is not None)
        assert torch.equal(got, jr.canvas_to_uint8_display(ref))
    # incoming canvas as background, through Renderer.render
    model = jr.merge_objects(objs)
    g = torch.Generator(device="cpu").manual_seed(1)
    bg = torch.rand(B, W, H, 3, generator=g).to(DEV)
    a = jr.Renderer.render(model, light, cam, jr.Buffers(torch.ones(B, W, H, device=DEV), (bg.clone(),)))
    u = jr.Renderer.render(model, light, cam, jr.Buffers(torch.ones(B, W, H, device=DEV), (bg.clone(),)),
                           display_uint8=(0.0, 0.0, 0.0))
    assert torch.equal(u.targets[0], jr.canvas_to_uint8_display(a.targets[0])) and torch.equal(u.zbuffer, a.zbuffer)
