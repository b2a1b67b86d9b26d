"""Parity against the REFERENCE'S OWN CODE.

`tests/golden/reference_run.npz` holds inputs and outputs of the unmodified reference (`/root/reference/renderer`)
executed on small scenes through a NumPy stand-in for jax (`tools/jax_numpy_shim`, generator
`tools/gen_reference_fixtures.py`): all seven built-in shaders through `pipeline.render`, the shadow pass,
`merge_objects`, `create_camera_from_parameters` and `Renderer.get_camera_image`; `reference_run_large.npz` the same
seven shaders on a 64x48 canvas covered to 86 % (`tools/gen_reference_fixtures_large.py`), `reference_run_tiled.npz` on a
136x96 canvas (the binned CUDA path: 3 x 2 tiles), `reference_run_brax84.npz`
the facade on the real Brax ant frame at 84x84 (BASELINE.json configs[1]'s canvas).  The CPU tests pin the oracle and
the host-side glue of the package against it; the GPU tests pin the CUDA path.

Tolerances: colours 1e-5 relative to the largest channel value (BASELINE.json), z 5e-6 absolute (~80 ulp of a window depth near 1:
the stand-in evaluates dot products through BLAS, not in the scalar order of the oracle); coverage (which pixels
were written) must agree except where the reference's competing edge values are within rounding, which is counted
and bounded."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import jaxrenderer_b200 as jr
from jaxrenderer_b200 import shaders as S
from oracle import jr_oracle as O

_GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
D = dict(np.load(os.path.join(_GOLDEN, "reference_run.npz")))
# soup5: 64x48, 48 triangles 2.6 x larger, 86 % of the canvas covered (tools/gen_reference_fixtures_large.py)
D.update(np.load(os.path.join(_GOLDEN, "reference_run_large.npz")))
# soup6: 136x96 -- beyond the single-tile kernel's 12 288 pixels: 3 x 2 tiles of the binned CUDA path (bitmasks,
# triangle queue, span raster, CTA-wide sweep); the GPU test checks that those kernels are the ones that ran
D.update(np.load(os.path.join(_GOLDEN, "reference_run_tiled.npz")))
SOUPS = sorted({k.split("/")[0] for k in D if k.startswith("soup")})
SHADERS = ("depth", "gouraud", "gouraud_texture", "phong", "phong_darboux", "phong_reflection",
           "phong_reflection_shadow")
Z_ATOL, C_RTOL, MAX_COVERAGE_FLIPS = 5e-6, 1e-5, 0   # no coverage flip is excused


def T(key, dev=None):
    t = torch.from_numpy(np.asarray(D[key]))
    return t.to(dev) if dev else t


def _scene(p, dev=None):
    cam = NS(world_to_clip=T(p + "/world_to_clip", dev), viewport=T(p + "/viewport", dev),
             world_to_eye_norm=T(p + "/world_to_eye_norm", dev))
    light = jr.LightSource(direction=T(p + "/light_direction", dev), colour=T(p + "/light_colour", dev))
    faces = T(p + "/faces", dev)
    n_tri = faces.shape[0]
    pos, nrm = T(p + "/position", dev), T(p + "/normal", dev)
    base = dict(position=pos, normal=nrm, uv=T(p + "/uv01", dev), light=light, light_dir_eye=T(p + "/light_dir_eye", dev),
                texture_shape=T(p + "/texture_shape", dev), texture_index=T(p + "/texture_index", dev),
                texture_offset=int(D[p + "/texture_offset"]), texture=T(p + "/atlas", dev),
                specular_map=T(p + "/specular_map", dev), ambient=T(p + "/ambient", dev),
                diffuse=T(p + "/diffuse", dev), specular=T(p + "/specular", dev))
    shadow_cam = NS(world_to_clip=T(p + "/shadow_world_to_clip", dev), viewport=T(p + "/shadow_viewport", dev))
    shadow = jr.Shadow(shadow_map=T(p + "/shadow_map", dev), strength=T(p + "/shadow_strength", dev), camera=shadow_cam)
    tex, uvt = T(p + "/texture", dev), T(p + "/uv_texel", dev)
    i2f = torch.arange(n_tri, dtype=torch.int32, device=dev).repeat_interleave(3)
    extras = {
        "depth": (S.DepthShader, S.DepthExtraInput(position=pos)),
        "gouraud": (S.GouraudShader, S.GouraudExtraInput(pos, T(p + "/colour", dev), nrm, light)),
        "gouraud_texture": (S.GouraudTextureShader, S.GouraudTextureExtraInput(pos, nrm, uvt, light, tex)),
        "phong": (S.PhongTextureShader, S.PhongTextureExtraInput(pos, nrm, uvt, light, tex)),
        "phong_darboux": (S.PhongTextureDarbouxShader, S.PhongTextureDarbouxExtraInput(
            pos, nrm, uvt, light, tex, T(p + "/normal_map", dev), i2f, faces)),
        "phong_reflection": (S.PhongReflectionTextureShader, S.PhongReflectionTextureExtraInput(**base)),
        "phong_reflection_shadow": (S.PhongReflectionShadowTextureShader,
                                    S.PhongReflectionShadowTextureExtraInput(**base, shadow=shadow, camera=cam)),
    }
    return cam, faces, extras, int(D[p + "/W"]), int(D[p + "/H"])


def _check(tag, z, c, p, name):
    zf = T(f"{p}/{name}/zbuffer")
    z = z.detach().cpu()
    wrote_ref, wrote = zf != 1.0, z != 1.0
    flips = int((wrote_ref != wrote).sum())
    same = wrote_ref == wrote
    dz = float((z - zf)[same].abs().max())
    msg = f"[{tag}] {p}/{name}: covered {int(wrote_ref.sum())}, coverage flips {flips}, max |dz| {dz:.3g}"
    assert flips <= MAX_COVERAGE_FLIPS, msg
    assert dz <= Z_ATOL, msg
    if c is not None:
        cf = T(f"{p}/{name}/canvas")
        dc = float((c.detach().cpu() - cf)[same].abs().max())
        msg += f", max |dcolour| {dc:.3g}"
        assert dc <= C_RTOL * max(1.0, float(cf.abs().max())), msg
    print(msg)


@pytest.mark.parametrize("p", SOUPS)
def test_oracle_matches_reference_run_all_shaders(p):
    cam, faces, extras, W, H = _scene(p)
    for name in SHADERS:
        _, extra = extras[name]
        z0, c0 = torch.ones(W, H), torch.full((W, H, 3), 0.25)
        ref = O.render(cam, name, z0, () if name == "depth" else (c0,), faces, extra)
        _check("oracle", ref.zbuffer, None if name == "depth" else ref.targets[0], p, name)
    # the shadow map the reference rendered (light camera given)
    sm = O.render_shadow_map(torch.full((W, H), torch.finfo(torch.float32).max), extras["depth"][1].position, faces,
                             extras["phong_reflection_shadow"][1].shadow.camera, 0.05)
    smf = T(p + "/shadow_map")
    assert bool(((sm < 1e30) == (smf < 1e30)).all())
    assert float((sm - smf)[smf < 1e30].abs().max()) < 2e-5


EDGE = ("tri0_backfacing", "tri0_frontfacing", "behind_and_straddling", "degenerate_and_duplicates", "old_z_closer")


def _edge_check(tag, z, name):
    zi = float(D[f"edge/{name}/z_init"])
    zf = T(f"edge/{name}/zbuffer")
    z = z.detach().cpu()
    flips = int(((zf != zi) != (z != zi)).sum())
    dz = float((z - zf).abs()[(zf != zi) == (z != zi)].max())
    print(f"[{tag}] edge/{name}: reference wrote {int((zf != zi).sum())} pixels, coverage flips {flips}, max |dz| {dz:.3g}")
    assert flips == 0 and dz <= Z_ATOL
    assert int((zf != zi).sum()) > 20


@pytest.mark.parametrize("name", EDGE)
def test_oracle_matches_reference_run_edge_cases(name):
    """Conventions restated from reading the reference, checked against its behaviour: a back-facing triangle 0
    still fills pixels nothing else covers, geometry behind / across the camera plane, degenerate and
    duplicate triangles (first index wins), no depth test against the incoming z-buffer."""
    cam = NS(world_to_clip=T("edge/world_to_clip"), viewport=T("edge/viewport"))
    W, H = int(D["edge/W"]), int(D["edge/H"])
    ref = O.render(cam, "depth", torch.full((W, H), float(D[f"edge/{name}/z_init"])), (), T(f"edge/{name}/faces"),
                   NS(position=T(f"edge/{name}/position")))
    _edge_check("oracle", ref.zbuffer, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", EDGE)
def test_cuda_path_matches_reference_run_edge_cases(name):
    dev = torch.device("cuda", 0)
    W, H = int(D["edge/W"]), int(D["edge/H"])
    camera = jr.Camera(*[None] * 8)._replace(world_to_clip=T("edge/world_to_clip", dev), viewport=T("edge/viewport", dev))
    z0 = torch.full((W, H), float(D[f"edge/{name}/z_init"]), device=dev)
    out = jr.render(camera, S.DepthShader, jr.Buffers(z0, ()), T(f"edge/{name}/faces", dev),
                    S.DepthExtraInput(position=T(f"edge/{name}/position", dev)))
    _edge_check("cuda", out.zbuffer, name)


def _facade_objects(dev=None):
    objs = []
    for i in range(3):
        g = lambda k: T(f"facade/obj{i}/{k}", dev)  # noqa: E731
        m = jr.Model(verts=g("verts"), norms=g("norms"), uvs=g("uvs"), faces=g("faces"), faces_norm=g("faces_norm"),
                     faces_uv=g("faces_uv"), diffuse_map=g("diffuse_map"), specular_map=g("specular_map"))
        objs.append(jr.ModelObject(model=m, local_scaling=g("local_scaling"), transform=g("transform")))
    cp = jr.CameraParameters(viewWidth=int(D["facade/camera_parameters/W"]), viewHeight=int(D["facade/camera_parameters/H"]),
                             position=T("facade/camera_parameters/position", dev),
                             target=T("facade/camera_parameters/target", dev), up=T("facade/camera_parameters/up", dev),
                             hfov=float(D["facade/camera_parameters/hfov"]), vfov=float(D["facade/camera_parameters/vfov"]))
    sp = jr.ShadowParameters(centre=T("facade/shadow_parameters/centre", dev), up=T("facade/shadow_parameters/up", dev),
                             strength=T("facade/shadow_parameters/strength", dev),
                             offset=float(D["facade/shadow_parameters/offset"]))
    return objs, cp, sp


def test_host_glue_matches_reference_run():
    """merge_objects and create_camera_from_parameters (torch builders, CPU) against the reference's outputs."""
    objs, cp, _ = _facade_objects()
    merged = jr.merge_objects(objs)
    for k in jr.MergedModel._fields:
        want = D[f"facade/merged/{k}"]
        got = getattr(merged, k)
        got = np.asarray(got.cpu() if isinstance(got, torch.Tensor) else got)
        assert got.shape == want.shape, (k, got.shape, want.shape)
        if want.dtype.kind == "f":
            np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6, err_msg=k)
        else:
            assert np.array_equal(got.astype(np.int64), want.astype(np.int64)), k
    cam = jr.Renderer.create_camera_from_parameters(cp)
    for k in jr.Camera._fields:
        np.testing.assert_allclose(getattr(cam, k).numpy(), D[f"facade/camera/{k}"], rtol=2e-5, atol=2e-5, err_msg=k)


# ----------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("p", SOUPS)
def test_cuda_path_matches_reference_run_all_shaders(p):
    dev = torch.device("cuda", 0)
    cam, faces, extras, W, H = _scene(p, dev)
    camera = jr.Camera(*[getattr(cam, k, None) if hasattr(cam, k) else None for k in jr.Camera._fields])
    camera = camera._replace(view=T(p + "/view", dev))
    from jaxrenderer_b200 import _native
    kernels = set()
    for name in SHADERS:
        shader, extra = extras[name]
        z0, c0 = torch.ones(W, H, device=dev), torch.full((W, H, 3), 0.25, device=dev)
        _native.kernel_timing(True)
        try:
            out = jr.render(camera, shader, jr.Buffers(z0, () if name == "depth" else (c0,)), faces, extra)
            kernels |= {k for k, _, _ in _native.kernel_times()}
        finally:
            _native.kernel_timing(False)
        _check("cuda", out.zbuffer, None if name == "depth" else out.targets[0], p, name)
    # which visibility path served the scene: the single-tile kernel up to 12 288 pixels, the two-level kernels beyond
    want = {"k_raster_tile", "memset+k_setup_bin"} if W * H * 8 > 96 * 1024 else {"k_vis3"}
    assert want <= kernels and not ({"k_raster_tile", "k_vis3"} - want) & kernels, (p, sorted(kernels))


@pytest.mark.gpu
@pytest.mark.parametrize("shadow", [True, False])
def test_cuda_facade_matches_reference_run(shadow):
    dev = torch.device("cuda", 0)
    objs, cp, sp = _facade_objects(dev)
    img = jr.Renderer.get_camera_image(objs, jr.LightParameters(), cp, cp.viewWidth, cp.viewHeight,
                                       shadow_param=sp if shadow else None)
    want = T("facade/with_shadow/canvas" if shadow else "facade/no_shadow/canvas")
    diff = (img.cpu() - want).abs().amax(-1)
    bad = int((diff > 2e-5).sum())
    print(f"facade shadow={shadow}: max |dcolour| {float(diff.max()):.3g}, pixels off by > 2e-5: {bad} of {diff.numel()}")
    assert bad == 0   # no allowance: every pixel of the reference's own facade render is reproduced


def test_host_helpers_match_reference_run():
    """Shapes, quaternion / rotation helpers, camera matrix builders and display utilities against the outputs of
    the reference's own functions for the same arguments."""
    from jaxrenderer_b200 import utils as U  # noqa: F401

    def close(name, got, want, tol=2e-6):
        got = np.asarray(got.detach().cpu() if isinstance(got, torch.Tensor) else got)
        want = np.asarray(want)
        assert got.shape == want.shape, (name, got.shape, want.shape)
        if want.dtype.kind == "f":
            assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max()), name
        else:
            assert np.array_equal(got.astype(np.int64), want.astype(np.int64)), name

    tex = T("helpers/cube/diffuse_map")
    cube = jr.create_cube(half_extents=torch.tensor((0.5, 1.5, 2.0)), texture_scaling=torch.tensor((2.0, 3.0)),
                          diffuse_map=tex, specular_map=torch.ones(3, 4) * 2.5)
    for k in cube._fields:
        close("cube." + k, getattr(cube, k), D["helpers/cube/" + k])
    for ax in (jr.UpAxis.X, jr.UpAxis.Y, jr.UpAxis.Z):
        cap = jr.create_capsule(radius=torch.tensor(0.25), half_height=torch.tensor(0.75), up_axis=ax, diffuse_map=tex,
                                specular_map=torch.ones(3, 4) * 2.5)
        for k in ("verts", "norms", "uvs", "faces"):
            close(f"capsule[{int(ax)}].{k}", getattr(cap, k), D[f"helpers/capsule_{int(ax)}/{k}"])
    axis, angle = T("helpers/rotation/axis"), T("helpers/rotation/angle")
    q1, q2 = jr.quaternion(axis, angle), T("helpers/rotation/quaternion2")
    close("quaternion", q1, D["helpers/rotation/quaternion"])
    close("quaternion_mul", jr.quaternion_mul(q1, q2), D["helpers/rotation/quaternion_mul"])
    close("rotation_matrix", jr.rotation_matrix(axis, angle), D["helpers/rotation/rotation_matrix"])
    close("normalise", jr.normalise(torch.tensor((3.0, -4.0, 12.0))), D["helpers/rotation/normalise"])
    mo = jr.ModelObject(model=cube).replace_with_orientation(q1).replace_with_position(torch.tensor((1.0, 2.0, 3.0)))
    close("ModelObject.transform", mo.transform, D["helpers/model_object/transform"])
    eye, centre, up = T("helpers/camera/eye"), T("helpers/camera/centre"), T("helpers/camera/up")
    C = jr.Camera
    close("view_matrix", C.view_matrix(eye, centre, up), D["helpers/camera/view"])
    close("view_matrix_inv", C.view_matrix_inv(eye, centre, up), D["helpers/camera/view_inv"])
    close("perspective", C.perspective_projection_matrix(40.0, 1.6, 0.1, 50.0), D["helpers/camera/perspective"])
    close("orthographic", C.orthographic_projection_matrix(-2.0, 3.0, -1.0, 1.5, 0.5, 20.0),
          D["helpers/camera/orthographic"])
    close("viewport", C.viewport_matrix(torch.tensor((1.0, 2.0)), torch.tensor((640.0, 480.0)), torch.tensor(2.0)),
          D["helpers/camera/viewport"])
    close("world_to_screen", C.world_to_screen_matrix(320, 200), D["helpers/camera/world_to_screen"])
    canvas = T("helpers/utils/canvas")
    close("transpose_for_display", jr.transpose_for_display(canvas), D["helpers/utils/transposed"])
    close("transpose_for_display(noflip)", jr.transpose_for_display(canvas, flip_vertical=False),
          D["helpers/utils/transposed_noflip"])
    close("build_texture_from_PyTinyrenderer", jr.build_texture_from_PyTinyrenderer(T("helpers/utils/pytiny_raw"), 4, 3),
          D["helpers/utils/pytiny_texture"])


BRAX_RUNS = ("reference_run_brax.npz", "reference_run_brax84.npz")   # 20x20, and configs[1]'s 84x84


def _brax_case(fixture, dev=None):
    """Frame 0 of the reference's own pre-generated Brax ant scene (18 objects, 3276 triangles, texture atlas, shadow
    pass) as `tools/gen_reference_fixtures_brax.py` gave it to the reference's `Renderer.get_camera_image`."""
    from tests.helpers import load_brax_fixture

    B = np.load(os.path.join(_GOLDEN, fixture))
    f, W, H = int(B["frame"]), int(B["W"]), int(B["H"])
    mv = (lambda t: t.to(dev)) if dev else (lambda t: t)
    objs, cam = load_brax_fixture()
    objs = [jr.ModelObject(model=type(o.model)(*[mv(t) for t in o.model]), local_scaling=mv(o.local_scaling[f]),
                           transform=mv(o.transform[f]), double_sided=mv(o.double_sided[f])) for o in objs]
    cp = jr.CameraParameters(viewWidth=W, viewHeight=H, viewDepth=float(cam.viewDepth[f]), near=float(cam.near[f]),
                             far=float(cam.far[f]), hfov=float(cam.hfov[f]), vfov=float(B["vfov"]),
                             position=mv(cam.position[f]), target=mv(cam.target[f]), up=mv(cam.up[f]))
    light = jr.LightParameters(direction=torch.from_numpy(B["light_direction"]), ambient=torch.from_numpy(B["ambient"]),
                               diffuse=torch.from_numpy(B["diffuse"]), specular=torch.from_numpy(B["specular"]))
    return objs, cp, light, jr.ShadowParameters(centre=mv(cam.target[f])), torch.from_numpy(B["canvas"]), W, H


def _assert_only_ties(tag, canvas, want, gap, texel_gap, shadow_gaps=(None, None)):
    """A pixel may differ from the reference's only where a DISCRETE choice of the reference hangs on the last bits of
    its arithmetic (the stand-in evaluates dot products through BLAS, the oracle and the kernels in scalar order:
    window depths differ by up to Z_ATOL): two triangles whose depths are closer than that (BASELINE.json's tie rule
    at the stand-in's noise level), or an atlas coordinate within TWO ulp of the interpolated uv of a texel boundary
    (`RenderOut.texel_gap`; the ground of a Brax scene carries uv ~ 10^4 and floor() flips its checker square), or a
    shadow test within Z_ATOL of flipping / a shadow-map lookup within 1e-3 pixel of the rounding boundary between two
    shadow-map pixels (`RenderOut.shadow_z_gap`, `shadow_xy_gap`).  Every differing pixel must be such a tie, and
    there must be few of them."""
    diff = (canvas.detach().cpu() - want).abs().amax(-1)
    bad = diff > 2e-5
    depth_tie, texel_tie = gap < Z_ATOL, texel_gap <= 2.0
    shadow_tie = torch.zeros_like(bad)
    if shadow_gaps[0] is not None:
        shadow_tie = (shadow_gaps[0] < Z_ATOL) | (shadow_gaps[1] < 1e-3)
    unexplained = bad & ~depth_tie & ~texel_tie & ~shadow_tie
    print(f"[{tag}] brax frame at {want.shape[0]}x{want.shape[1]}: max |dcolour| {float(diff.max()):.3g}, pixels off by > 2e-5: "
          f"{int(bad.sum())} of {diff.numel()} -- depth ties {int((bad & depth_tie).sum())}, texel-boundary ties "
          f"{int((bad & ~depth_tie & texel_tie).sum())}, shadow-test ties {int((bad & ~depth_tie & ~texel_tie & shadow_tie).sum())}, "
          f"unexplained {int(unexplained.sum())}")
    for x, y in unexplained.nonzero().tolist():
        print(f"   unexplained ({x}, {y}): got {canvas[x, y].tolist()} want {want[x, y].tolist()} depth gap {float(gap[x, y]):.3g} "
              f"texel gap {float(texel_gap[x, y]):.3g} shadow gaps "
              f"{[float(g[x, y]) for g in shadow_gaps if g is not None]}")
    assert int(unexplained.sum()) == 0
    assert int(bad.sum()) <= max(1, diff.numel() // 500)       # <= 0.2 % of the frame (measured: 8 oracle / 10 CUDA of 7056)
    return bad


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", BRAX_RUNS)
def test_cuda_facade_matches_reference_run_brax_frame(fixture):
    """`Renderer.get_camera_image` of the reference on its own Brax ant frame at 20x20 and at 84x84 (the canvas
    BASELINE.json's configs[1] names) vs the CUDA path; the tie diagnostics come from the oracle."""
    dev = torch.device("cuda", 0)
    objs, cp, light, sp, want, W, H = _brax_case(fixture, dev)
    img = jr.Renderer.get_camera_image(objs, light, cp, W, H, shadow_param=sp)
    o_objs, o_cp, o_light, o_sp, _, _, _ = _brax_case(fixture)
    o_canvas, gap = _oracle_facade(o_objs, o_cp, o_light, o_sp, W, H)
    _assert_only_ties("cuda", img, want, gap, _oracle_facade.texel_gap, _oracle_facade.shadow_gaps)
    _assert_only_ties("cuda vs oracle", img, o_canvas, gap, _oracle_facade.texel_gap, _oracle_facade.shadow_gaps)


def _oracle_facade(objs, cp, light, sp, W, H):
    """The reference facade restated on the CPU: host glue of the package (merge_objects, camera builders on CPU
    tensors) + the oracle's `Renderer.render`."""
    merged = jr.merge_objects(objs)
    cam = jr.Renderer.create_camera_from_parameters(cp)
    res = O.renderer_render(merged, light, cam, torch.ones(W, H), torch.ones(W, H, 3), shadow_param=sp)
    _oracle_facade.texel_gap = res["out"].texel_gap
    _oracle_facade.shadow_gaps = (res["out"].shadow_z_gap, res["out"].shadow_xy_gap)
    return res["out"].targets[0], res["out"].gap


@pytest.mark.parametrize("shadow", [True, False])
def test_oracle_facade_matches_reference_run(shadow):
    objs, cp, sp = _facade_objects()
    canvas, gap = _oracle_facade(objs, cp, jr.LightParameters(), sp if shadow else None, cp.viewWidth, cp.viewHeight)
    want = T("facade/with_shadow/canvas" if shadow else "facade/no_shadow/canvas")
    diff = (canvas - want).abs().amax(-1)
    assert int((diff > 2e-5).sum()) == 0, float(diff.max())


@pytest.mark.parametrize("fixture", BRAX_RUNS)
def test_oracle_facade_matches_reference_run_brax_frame(fixture):
    objs, cp, light, sp, want, W, H = _brax_case(fixture)
    canvas, gap = _oracle_facade(objs, cp, light, sp, W, H)
    _assert_only_ties("oracle", canvas, want, gap, _oracle_facade.texel_gap, _oracle_facade.shadow_gaps)


# ------------------------------------------------------------------ the reference's batching idiom, run by the reference
def _vmap_case(dev=None):
    """`tools/gen_reference_fixtures_vmap.py`: three poses of a small scene through
    `jax.vmap(lambda model, buffer: Renderer.render(...))(batch_models(merged_models), buffers)`
    (`examples/batch_rendering.py:83-95`), executed by the unmodified reference."""
    V = np.load(os.path.join(_GOLDEN, "reference_run_vmap.npz"))
    mv = (lambda a: torch.from_numpy(np.asarray(a)).to(dev)) if dev else (lambda a: torch.from_numpy(np.asarray(a)))
    W, H, n = int(V["W"]), int(V["H"]), int(V["poses"])
    merged = []
    for k in range(n):
        objs = []
        for i in range(3):
            g = lambda f: mv(V[f"pose{k}/obj{i}/{f}"])  # noqa: E731
            m = jr.Model(verts=g("verts"), norms=g("norms"), uvs=g("uvs"), faces=g("faces"), faces_norm=g("faces_norm"),
                         faces_uv=g("faces_uv"), diffuse_map=g("diffuse_map"), specular_map=g("specular_map"))
            objs.append(jr.ModelObject(model=m, local_scaling=g("local_scaling"), transform=g("transform")))
        merged.append(jr.merge_objects(objs))
    cp = jr.CameraParameters(viewWidth=W, viewHeight=H, position=mv(V["cam_position"]), target=mv(V["cam_target"]),
                             up=mv(V["cam_up"]), hfov=float(V["hfov"]), vfov=float(V["vfov"]))
    camera = jr.Renderer.create_camera_from_parameters(cp)
    sp = jr.ShadowParameters(centre=mv(V["shadow_centre"]))
    return jr.batch_models(merged), jr.LightParameters(), camera, sp, V, W, H, n


def _check_vmap(tag, z, c, V):
    zf, cf = torch.from_numpy(V["zbuffer"]), torch.from_numpy(V["canvas"])
    z, c = z.detach().cpu(), c.detach().cpu()
    flips = int(((zf != 1.0) != (z != 1.0)).sum())
    dz, dc = float((z - zf).abs().max()), float((c - cf).abs().max())
    print(f"[{tag}] vmap idiom, {tuple(cf.shape)}: covered {int((zf != 1.0).sum())}, coverage flips {flips}, "
          f"max |dz| {dz:.3g}, max |dcolour| {dc:.3g}")
    assert tuple(z.shape) == tuple(zf.shape) and tuple(c.shape) == tuple(cf.shape)
    assert flips == 0 and dz <= Z_ATOL and dc <= 2e-5


def test_oracle_matches_reference_run_vmap_idiom():
    batch, light, camera, sp, V, W, H, n = _vmap_case()
    zs, cs = [], []
    for k in range(n):
        model = type(batch)(*[(f[k] if isinstance(f, torch.Tensor) and name != "offset" else f)
                              for name, f in zip(batch._fields, batch)])
        res = O.renderer_render(model, light, camera, torch.ones(W, H), torch.ones(W, H, 3), shadow_param=sp)
        zs.append(res["out"].zbuffer); cs.append(res["out"].targets[0])
    _check_vmap("oracle", torch.stack(zs), torch.stack(cs), V)


@pytest.mark.gpu
def test_cuda_native_batch_and_torch_vmap_match_reference_run_vmap_idiom():
    """The package's leading batch axis IS the reference's `vmap`: `Renderer.render` on `batch_models(...)` with batched
    buffers, and `torch.func.vmap` of the reference's own lambda, against what the reference's `jax.vmap` produced."""
    from torch.func import vmap

    dev = torch.device("cuda", 0)
    batch, light, camera, sp, V, W, H, n = _vmap_case(dev)
    buffers = jr.Renderer.create_buffers(W, H, batch=n, device=dev)
    out = jr.Renderer.render(model=batch, light=light, camera=camera, buffers=buffers, shadow_param=sp)
    _check_vmap("cuda, native batch", out.zbuffer, out.targets[0], V)
    in_model = type(batch)(*[(None if name == "offset" or not isinstance(f, torch.Tensor) else 0)
                             for name, f in zip(batch._fields, batch)])
    zv, cv = vmap(lambda model, z, c: (lambda o: (o.zbuffer, o.targets[0]))(
        jr.Renderer.render(model=model, light=light, camera=camera, buffers=jr.Buffers(z, (c,)), shadow_param=sp)),
        in_dims=(in_model, 0, 0))(batch, buffers.zbuffer, buffers.targets[0])
    _check_vmap("cuda, torch.func.vmap", zv, cv, V)
    assert torch.equal(zv, out.zbuffer) and torch.equal(cv, out.targets[0])
