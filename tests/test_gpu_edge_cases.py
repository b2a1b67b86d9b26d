"""GPU parity on edge cases: empty mesh, degenerate triangles, geometry straddling / behind the
camera (reference examples/behind_camera.py), odd canvas sizes, canvases around the single-tile /
binned switch, batch broadcast of the camera, buffers with non-default contents."""
from types import SimpleNamespace as NS

import pytest
import torch

import jaxrenderer_b200 as jr
from jaxrenderer_b200.shaders import DepthExtraInput, DepthShader, GouraudExtraInput, GouraudShader
from oracle import jr_oracle as O
from tests.helpers import assert_parity, compare, random_mesh_scene

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cam_d(cam):
    return type(cam)(*[t.to(DEV) for t in cam])


def test_empty_mesh_leaves_buffers_untouched():
    cam = jr.Renderer.create_camera_from_parameters(jr.CameraParameters(viewWidth=20, viewHeight=12))
    z0 = torch.rand(20, 12)
    c0 = torch.rand(20, 12, 3)
    out, tri = jr.render(_cam_d(cam), DepthShader, jr.Buffers(z0.to(DEV), ()),
                         torch.zeros(0, 3, dtype=torch.int32, device=DEV),
                         DepthExtraInput(position=torch.zeros(3, 3, device=DEV)), return_tri_id=True)
    assert torch.equal(out.zbuffer.cpu(), z0) and bool((tri == -1).all())
    light = jr.LightSource(torch.tensor((0.0, 0.0, -1.0)), torch.ones(3))
    out = jr.render(_cam_d(cam), GouraudShader, jr.Buffers(z0.to(DEV), (c0.to(DEV),)),
                    torch.zeros(0, 3, dtype=torch.int32, device=DEV),
                    GouraudExtraInput(torch.zeros(3, 3, device=DEV), torch.zeros(3, 3, device=DEV),
                                      torch.ones(3, 3, device=DEV), light))
    assert torch.equal(out.zbuffer.cpu(), z0) and torch.equal(out.targets[0].cpu(), c0)


def test_degenerate_and_duplicate_triangles():
    """Zero-area triangles are dropped by |det| > 1e-6; exact duplicates tie on depth and the
    lowest index wins (shader.py:217)."""
    s = random_mesh_scene(4, n_tri=30)
    pos = s.pos.clone()
    pos[3:6] = pos[3:4]                      # triangle 1 collapses to a point
    pos[9] = pos[10]                         # triangle 3 has two equal vertices
    faces = torch.cat((s.faces, s.faces[5:8]))   # duplicates of triangles 5..7 with higher ids
    z0 = torch.full((s.W, s.H), 1.0)
    out, tri = jr.render(_cam_d(s.cam), DepthShader, jr.Buffers(z0.to(DEV), ()), faces.to(DEV),
                         DepthExtraInput(position=pos.to(DEV)), return_tri_id=True)
    ref = O.render(s.cam, "depth", z0, (), faces, NS(position=pos))
    rep = compare("degenerate", out.zbuffer, None, tri, ref)
    print(rep)
    assert rep["tri_mismatch"] == 0 and rep["z_not_bit_equal"] == 0
    assert not bool(((tri.cpu() >= 30)).any()), "duplicates must lose the tie to the lower index"
    assert not bool(((tri.cpu() == 1) | (tri.cpu() == 3)).any())


@pytest.mark.parametrize("wh", [(64, 48), (640, 480)])
def test_geometry_behind_and_straddling_the_camera(wh):
    """examples/behind_camera.py: a 20 x 20 slab around the eye -- triangles with some or all w <= 0
    (no clipping in the reference: README.md:115)."""
    W, H = wh
    tex = torch.tensor([[[1.0, 0, 0], [0, 1.0, 0]], [[0, 0, 1.0], [1.0, 1.0, 0]]])
    slab = jr.create_cube(torch.tensor((10.0, 10.0, 0.03)), torch.tensor((160.0, 160.0)), tex, torch.ones(2, 2) * 2)
    small = jr.create_cube(torch.tensor((1.0, 1.0, 0.03)), torch.tensor((16.0, 16.0)), tex, torch.ones(2, 2) * 2)
    t = torch.eye(4); t[2, 3] = 0.5
    model = jr.merge_objects([jr.ModelObject(model=slab), jr.ModelObject(model=small, transform=t)])
    camp = jr.CameraParameters(viewWidth=W, viewHeight=H, position=torch.tensor([2.5894797, -2.5876467, 1.9174135]),
                               hfov=58.0, vfov=32.625)
    cam = jr.Renderer.create_camera_from_parameters(camp)
    clip_w = (torch.cat((model.verts, torch.ones(len(model.verts), 1)), 1) @ cam.world_to_clip.T)[:, 3]
    assert bool((clip_w < 0).any()) and bool((clip_w > 0).any()), "scene must straddle w = 0"
    z0 = torch.full((W, H), 1.0)
    out, tri = jr.render(_cam_d(cam), DepthShader, jr.Buffers(z0.to(DEV), ()), model.faces.to(DEV),
                         DepthExtraInput(position=model.verts.to(DEV)), return_tri_id=True)
    ref = O.render(cam, "depth", z0, (), model.faces, NS(position=model.verts))
    rep = compare(f"behind{W}", out.zbuffer, None, tri, ref)
    print(rep)
    assert_parity(rep)
    assert int((ref.tri_id >= 0).sum()) > W * H // 3
    # and through the full renderer (phong_reflection + shadow)
    img = jr.Renderer.render(type(model)(*[v.to(DEV) if isinstance(v, torch.Tensor) else v for v in model]),
                             jr.LightParameters(), _cam_d(cam), jr.Renderer.create_buffers(W, H, device=DEV),
                             shadow_param=jr.ShadowParameters()).targets[0]
    assert bool(torch.isfinite(img).all())


@pytest.mark.parametrize("wh", [(1, 1), (7, 3), (3, 97), (110, 110), (111, 111), (255, 48), (256, 48), (129, 65)])
def test_odd_canvas_sizes(wh):
    """Canvas shapes around every internal switch: 1x1, non-square, single tile <-> binned (96 KB of
    keys, 255-pixel side limit), partial edge tiles."""
    W, H = wh
    s = random_mesh_scene(7, n_tri=50, W=W, H=H)
    z0 = torch.full((W, H), 2.0)
    c0 = torch.full((W, H, 3), 0.5)
    ex = GouraudExtraInput(s.pos, s.col, s.nrm, s.light)
    out, tri = jr.render(_cam_d(s.cam), GouraudShader, jr.Buffers(z0.to(DEV), (c0.to(DEV),)), s.faces.to(DEV),
                         GouraudExtraInput(s.pos.to(DEV), s.col.to(DEV), s.nrm.to(DEV),
                                           jr.LightSource(s.light.direction.to(DEV), s.light.colour.to(DEV))),
                         return_tri_id=True)
    ref = O.render(s.cam, "gouraud", z0, (c0,), s.faces, ex)
    rep = compare(f"{W}x{H}", out.zbuffer, out.targets[0], tri, ref)
    print(rep)
    assert_parity(rep)


def test_many_large_overlapping_triangles():
    """Stress the warp / hierarchical raster paths and their queue overflow: 300 screen-filling
    triangles at distinct depths on a single-tile canvas and on a binned canvas."""
    for (W, H) in ((96, 80), (320, 200)):
        g = torch.Generator().manual_seed(3)
        n = 300
        base = torch.tensor([[-3.0, -3.0, 0.0], [3.0, -3.0, 0.0], [0.0, 3.5, 0.0]])
        pos = (base[None] + torch.rand(n, 3, 3, generator=g) * 0.5)
        pos[:, :, 2] = torch.linspace(-1.5, 0.5, n)[:, None] + torch.rand(n, 3, generator=g) * 0.01
        pos = pos.reshape(-1, 3)
        faces = torch.arange(3 * n, dtype=torch.int32).reshape(n, 3)
        cam = jr.Renderer.create_camera_from_parameters(jr.CameraParameters(
            viewWidth=W, viewHeight=H, position=torch.tensor((0.0, 0.0, 4.0)), up=(0.0, 1.0, 0.0)))
        z0 = torch.full((W, H), 1.0)
        out, tri = jr.render(_cam_d(cam), DepthShader, jr.Buffers(z0.to(DEV), ()), faces.to(DEV),
                             DepthExtraInput(position=pos.to(DEV)), return_tri_id=True)
        ref = O.render(cam, "depth", z0, (), faces, NS(position=pos))
        rep = compare(f"large{W}", out.zbuffer, None, tri, ref)
        print(rep)
        assert_parity(rep)
        assert int((ref.tri_id >= 0).sum()) > W * H // 4


def test_out_of_range_indices_are_clamped_not_read():
    """Bad face indices never read out of bounds: they are clamped (what the reference's gathers do)."""
    s = random_mesh_scene(9, n_tri=20)
    faces = s.faces.clone()
    faces[3, 1] = 10_000_000
    faces[7, 0] = -5
    z0 = torch.full((s.W, s.H), 1.0)
    out, tri = jr.render(_cam_d(s.cam), DepthShader, jr.Buffers(z0.to(DEV), ()), faces.to(DEV),
                         DepthExtraInput(position=s.pos.to(DEV)), return_tri_id=True)
    ref = O.render(s.cam, "depth", z0, (), faces, NS(position=s.pos))
    rep = compare("clamped", out.zbuffer, None, tri, ref)
    assert_parity(rep)


# ------------------------------------------------------------------ fused camera construction (8f-2)
def _torch_camera(cp):
    """The differentiable torch builders (forced by a requires_grad leaf)."""
    pos = torch.as_tensor(cp.position, dtype=torch.float32).clone().requires_grad_(True)
    return jr.Renderer.create_camera_from_parameters(cp._replace(position=pos))


@pytest.mark.gpu
@pytest.mark.parametrize("batched", [False, True])
def test_camera_kernel_matches_torch_builders(batched):
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    B = 37
    shape = (B,) if batched else ()
    cp = jr.CameraParameters(
        viewWidth=84, viewHeight=60, viewDepth=1.0, near=0.1, far=(torch.rand(shape, generator=g) * 50 + 20).to(dev),
        hfov=58.0, vfov=(torch.rand(shape, generator=g) * 20 + 30).to(dev),
        position=(torch.rand(*shape, 3, generator=g) * 4 + 1).to(dev),
        target=(torch.rand(*shape, 3, generator=g) - 0.5).to(dev), up=(0.0, 0.0, 1.0))
    fused = jr.Renderer.create_camera_from_parameters(cp)
    ref = _torch_camera(cp)
    for name, a, b in zip(fused._fields, fused, ref):
        assert a.shape == b.shape or a.shape == b.shape[-2:] or b.shape == a.shape[-2:], (name, a.shape, b.shape)
        torch.testing.assert_close(a.expand_as(b) if a.ndim < b.ndim else a, b.detach().expand_as(a),
                                   rtol=2e-5, atol=1e-6, msg=lambda m: f"{name}: {m}")


@pytest.mark.gpu
@pytest.mark.parametrize("batched", [False, True])
def test_light_camera_kernel_matches_torch_builders(batched):
    from jaxrenderer_b200.geometry import camera_build_native
    from jaxrenderer_b200.shadow import Shadow
    from jaxrenderer_b200 import _native

    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(6)
    shape = (19,) if batched else ()
    centre = (torch.rand(*shape, 3, generator=g) - 0.5).to(dev)
    ld = torch.tensor((0.3, -0.6, 0.9), device=dev)
    up = torch.tensor((0.0, 1.0, 0.0), device=dev)
    viewport = jr.Camera.viewport_matrix(torch.zeros(2), torch.tensor((84.0, 60.0)), torch.tensor(1.0)).to(dev)
    fused = camera_build_native(_native.JR_CAMERA_LIGHT, (
        (centre, 3), (ld, 3), (up, 3), (10.0, 1), (-1.0, 1), (1.0, 1), (-1.0, 1), (1.0, 1), (-1.0, 1), (1.0, 1)),
        dev, viewport=viewport)
    ref = Shadow._light_camera(centre, ld, up, 10.0, viewport, dev)
    for name, a, b in zip(fused._fields, fused, ref):
        b = b.expand_as(a) if b.ndim < a.ndim else b
        torch.testing.assert_close(a, b, rtol=2e-5, atol=1e-6, msg=lambda m: f"{name}: {m}")


# ------------------------------------------------------------------ CUDA graph capture of the facade
@pytest.mark.gpu
def test_facade_captures_into_a_cuda_graph():
    """After warm-up the whole ``get_camera_image`` call neither allocates outside the caching allocator nor
    synchronises, so it captures into one CUDA graph; the replay re-reads the (updated) input tensors."""
    from tests.helpers import load_brax_fixture

    dev = torch.device("cuda", 0)
    objs, cam = load_brax_fixture()
    objs = [jr.ModelObject(model=type(o.model)(*[t.to(dev) for t in o.model]), local_scaling=o.local_scaling.to(dev),
                           transform=o.transform.to(dev), double_sided=o.double_sided.to(dev)) for o in objs]
    cam = type(cam)(*[v.to(dev) if isinstance(v, torch.Tensor) else v for v in cam])._replace(viewWidth=84, viewHeight=84)
    light = jr.LightParameters()
    sp = jr.ShadowParameters(centre=cam.target)

    def full():
        return jr.Renderer.get_camera_image(objs, light, cam, 84, 84, shadow_param=sp)

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            full()
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        img = full()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(img, full())
    # move the robot: same graph, new transforms
    with torch.no_grad():
        for o in objs[1:]:
            o.transform[..., 2, 3] += 0.05
    graph.replay()
    torch.cuda.synchronize()
    moved = full()
    assert torch.equal(img, moved)
    assert img.shape == (4, 84, 84, 3) and bool(torch.isfinite(img).all())


@pytest.mark.gpu
def test_library_kernel_timing_lists_every_launch_of_a_call():
    """`jr_debug_kernel_timing` / `jr_debug_kernel_times` (include/jr_b200.h): one event per launch, names in launch
    order, grouped by entry-point call; nothing is recorded while the switch is off."""
    from jaxrenderer_b200 import _native

    s = random_mesh_scene(3, n_tri=30, W=48, H=40)
    cam = _cam_d(s.cam)
    faces, pos, nrm, col = (t.to(DEV) for t in (s.faces, s.pos, s.nrm, s.col))
    light = jr.LightSource(torch.tensor((0.0, 0.3, -1.0), device=DEV), torch.ones(3, device=DEV))

    def both():
        z = jr.render(cam, DepthShader, jr.Buffers(torch.ones(48, 40, device=DEV), ()), faces, DepthExtraInput(pos))
        g = jr.render(cam, GouraudShader, jr.Buffers(torch.ones(48, 40, device=DEV), (torch.zeros(48, 40, 3, device=DEV),)),
                      faces, GouraudExtraInput(pos, col, nrm, light))
        return z, g

    plain = both()
    _native.kernel_timing(True)
    try:
        timed = both()
        rows = _native.kernel_times()
        assert _native.kernel_times() == []          # a read forgets what it returned
    finally:
        _native.kernel_timing(False)
    assert [(n, c) for n, _, c in rows] == [("k_vis3", 0), ("k_vis3", 1), ("k_tri_attr", 1), ("k_shade_rec", 1)], rows
    assert all(0.0 < ms < 50.0 for _, ms, _ in rows), rows
    for a, b in zip(plain, timed):                   # the marks change nothing
        assert torch.equal(a.zbuffer, b.zbuffer)
    both()
    assert _native.kernel_times() == []              # switched off: nothing recorded


@pytest.mark.gpu
def test_graphed_replays_the_facade_with_new_inputs():
    """`jr.graphed` (the role `jax.jit` plays in the reference's examples): one capture, then replays with the inputs
    copied into the graph's buffers -- bit-equal to the eager call for every new set of transforms / eye positions, one
    graph per signature, rejected argument kinds reported."""
    from tests.helpers import load_brax_fixture

    dev = torch.device("cuda", 0)
    objs, cam = load_brax_fixture()
    objs = [jr.ModelObject(model=type(o.model)(*[t.to(dev) for t in o.model]), local_scaling=o.local_scaling.to(dev),
                           transform=o.transform.to(dev), double_sided=o.double_sided.to(dev)) for o in objs]
    cam = type(cam)(*[v.to(dev) if isinstance(v, torch.Tensor) else v for v in cam])._replace(viewWidth=84, viewHeight=84)
    light, sp = jr.LightParameters(), jr.ShadowParameters(centre=cam.target)

    def frame(transforms, eye, size=84):
        moved = [o._replace(transform=t) for o, t in zip(objs, transforms)]
        return jr.Renderer.get_camera_image(moved, light, cam._replace(position=eye, viewWidth=size, viewHeight=size),
                                            size, size, shadow_param=sp)

    render = jr.graphed(frame)
    tf = [o.transform.clone() for o in objs]
    eye = cam.position.clone()
    for step in range(3):
        got = render(tf, eye)
        want = frame(tf, eye)
        assert torch.equal(got, want), step
        tf = [t.clone() for t in tf]
        for t in tf[1:]:
            t[..., 2, 3] += 0.04                          # the robot rises, the camera drifts
        eye = eye + torch.tensor((0.05, -0.02, 0.03), device=dev)
    assert render.cache_size() == 1
    small = render(tf, eye, 64)                            # a static argument changed: a second graph
    assert render.cache_size() == 2 and small.shape[1:3] == (64, 64) and torch.equal(small, frame(tf, eye, 64))
    with pytest.raises(ValueError, match="integer tensors"):
        render(tf, eye, objs[0].model.faces)
    with pytest.raises(ValueError, match="CUDA tensors"):
        render([t.cpu() for t in tf], eye.cpu())
