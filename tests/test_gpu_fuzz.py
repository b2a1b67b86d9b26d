"""Fuzz test of the conservative culling (bbox margin, back-face / behind-camera rejects): scenes of
thousands of tiny, needle-shaped and near-degenerate triangles at random sub-pixel positions, compared
bit-for-bit with the brute-force C oracle.  A triangle wrongly dropped before rasterisation shows up as
a triangle-id mismatch."""
import math

import numpy as np
import pytest
import torch

import jaxrenderer_b200 as jr
from jaxrenderer_b200.shaders import DepthExtraInput, DepthShader
from oracle import c_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _scene(seed: int, W: int, H: int, n_tri: int, B: int, len_px=(0.05, 6.0), aligned: float = 0.0,
           asp_min: float = 1.0):
    """`len_px`: range of the long edge in pixels (log-uniform); `aligned`: share of the triangles whose long edge is
    exactly parallel to a screen axis."""
    rng = np.random.default_rng(seed)
    d = 3.0
    fovy = 50.0
    px = 2 * d * math.tan(math.radians(fovy) / 2) / H            # world size of one pixel at z = 0
    cx = rng.uniform(-W / 2 - 2, W / 2 + 2, size=(B, n_tri)) * px
    cy = rng.uniform(-H / 2 - 2, H / 2 + 2, size=(B, n_tri)) * px
    L = np.exp(rng.uniform(math.log(len_px[0]), math.log(len_px[1]), size=(B, n_tri))) * px
    asp = np.exp(rng.uniform(math.log(asp_min), math.log(3000.0), size=(B, n_tri)))
    h = L / asp
    ang = rng.uniform(0, 2 * math.pi, size=(B, n_tri))
    ang = np.where(rng.uniform(size=(B, n_tri)) < aligned, np.round(ang / (math.pi / 2)) * (math.pi / 2), ang)
    t = rng.uniform(0.05, 0.95, size=(B, n_tri))                    # foot of the height on the long edge
    flip = rng.uniform(size=(B, n_tri)) < 0.3
    ux, uy = np.cos(ang), np.sin(ang)
    p0 = np.stack((cx - 0.5 * L * ux, cy - 0.5 * L * uy), -1)
    p1 = np.stack((cx + 0.5 * L * ux, cy + 0.5 * L * uy), -1)
    foot = p0 + (p1 - p0) * t[..., None]
    p2 = foot + np.stack((-uy, ux), -1) * h[..., None]
    z = rng.uniform(-0.3, 0.3, size=(B, n_tri, 3))
    tri = np.stack((p0, p1, p2), axis=2)                            # (B,T,3,2)
    tri = np.where(flip[..., None, None], tri[:, :, ::-1], tri)
    pos = np.concatenate((tri, z[..., None]), -1).reshape(B, n_tri * 3, 3).astype(np.float32)
    faces = np.arange(n_tri * 3, dtype=np.int32).reshape(n_tri, 3)
    cam = jr.Renderer.create_camera_from_parameters(jr.CameraParameters(
        viewWidth=W, viewHeight=H, hfov=2 * math.degrees(math.atan(math.tan(math.radians(fovy) / 2) * W / H)),
        vfov=fovy, position=torch.tensor((0.0, 0.0, d)), target=torch.zeros(3), up=(0.0, 1.0, 0.0)))
    return torch.from_numpy(pos), torch.from_numpy(faces), cam


@pytest.mark.parametrize("case", [(84, 84, 4000, 12, 1), (32, 32, 1500, 16, 2), (200, 150, 20000, 3, 3),
                                  (640, 360, 60000, 1, 4)])
def test_fuzz_small_and_needle_triangles(case):
    W, H, n_tri, B, seed = case
    pos, faces, cam = _scene(seed, W, H, n_tri, B)
    camd = type(cam)(*[t.to(DEV) for t in cam])
    out, tri = jr.render(camd, DepthShader, jr.Buffers(torch.full((B, W, H), 1.0, device=DEV), ()), faces.to(DEV),
                         DepthExtraInput(position=pos.to(DEV)), return_tri_id=True)
    zo, to = c_oracle.render_depth(cam.world_to_clip.numpy(), cam.viewport.numpy(), pos.numpy(), faces.numpy(),
                                   np.ones((B, W, H), np.float32))
    got = tri.cpu().numpy()
    mism = int((got != to).sum())
    covered = int((to >= 0).sum())
    print(f"{W}x{H} T={n_tri} B={B}: covered {covered} px by {len(np.unique(to)) - 1} triangles, mismatches {mism}")
    assert covered > 0.02 * B * W * H
    assert mism == 0
    assert np.array_equal(out.zbuffer.cpu().numpy(), zo)
    # the guard of the conservative culls: every triangle brute-forced over the whole canvas
    if W * H * n_tri * B <= 3_000_000_000:
        from jaxrenderer_b200 import pipeline
        audit = pipeline.audit_cull(camd, faces.to(DEV), pos.to(DEV))
        print("  cull audit:", audit)
        assert audit["filter_lost_triangles"] == 0 and audit["pixels_outside_bbox"] == 0
        assert audit["pixels_of_rejected_triangles"] == 0 and audit["triangles_kept"] > 0


@pytest.mark.parametrize("tri_id", [True, False])
@pytest.mark.parametrize("case", [(320, 200, 2500, 2, 11, 0.0, 1.0), (640, 360, 3000, 1, 12, 0.25, 1.0),
                                  (200, 150, 1500, 2, 13, 1.0, 1.0), (640, 360, 20000, 1, 14, 0.1, 30.0)])
def test_fuzz_long_needles_on_tiled_canvases(case, tri_id):
    """The span raster of the tiled path (`raster_span2_warp`: line intervals from analytic roots + a rounding-error
    margin, exact per-pixel test inside): thousands of LONG needles -- 16 to 300 pixels, aspect ratios up to 3000, any
    orientation, a share of them exactly parallel to a screen axis (the edge coefficient along the line vanishes) --
    crossing several 64x64 tiles, bit-for-bit against the brute-force C oracle, with packed keys (triangle ids) and with
    the z-only keys of a depth pass.  A pixel missed by a too-narrow interval shows up as a mismatch."""
    W, H, n_tri, B, seed, aligned, asp_min = case
    pos, faces, cam = _scene(seed, W, H, n_tri, B, len_px=(16.0, 300.0), aligned=aligned, asp_min=asp_min)
    camd = type(cam)(*[t.to(DEV) for t in cam])
    res = jr.render(camd, DepthShader, jr.Buffers(torch.full((B, W, H), 1.0, device=DEV), ()), faces.to(DEV),
                    DepthExtraInput(position=pos.to(DEV)), return_tri_id=tri_id)
    out, tri = res if tri_id else (res, None)
    zo, to = c_oracle.render_depth(cam.world_to_clip.numpy(), cam.viewport.numpy(), pos.numpy(), faces.numpy(),
                                   np.ones((B, W, H), np.float32))
    covered = int((to >= 0).sum())
    zg = out.zbuffer.cpu().numpy()
    print(f"{W}x{H} T={n_tri} B={B} aligned={aligned} tri_id={tri_id}: covered {covered} px by {len(np.unique(to)) - 1} "
          f"triangles, z mismatches {int((zg != zo).sum())}")
    assert covered > 0.05 * B * W * H
    if tri_id:
        assert int((tri.cpu().numpy() != to).sum()) == 0
    assert np.array_equal(zg, zo)


def test_cull_audit_reads_zero_on_brax_scenes():
    """The same guard on the scene families of the benchmark: synthetic ant-like scenes at 84x84 and 32x32, the real
    Brax ant fixture (instanced geometry), and a 480x270 binned-path scene."""
    from jaxrenderer_b200 import pipeline, synthetic
    from tests.helpers import load_brax_fixture

    for (W, H, B, n_caps) in ((84, 84, 16, 10), (32, 32, 16, 10), (480, 270, 1, 17)):
        objs, eye, tgt = synthetic.brax_like_objects(B, n_capsules=n_caps, env0=123, device=DEV)
        cam = synthetic.brax_cameras(eye, tgt, W, H)
        camd = type(cam)(*[t.to(DEV) for t in cam])
        m = jr.merge_objects(objs)
        audit = pipeline.audit_cull(camd, m.faces, m.verts)          # factored geometry, instanced in the kernel
        print(f"  {W}x{H} B={B}: {audit}")
        assert audit["filter_lost_triangles"] == 0 and audit["pixels_outside_bbox"] == 0
        assert audit["pixels_of_rejected_triangles"] == 0 and audit["inside_pixels"] > W * H * B
    objs, camp = load_brax_fixture()
    objs_d = [jr.ModelObject(model=type(o.model)(*[t.to(DEV) for t in o.model]), local_scaling=o.local_scaling.to(DEV),
                             transform=o.transform.to(DEV), double_sided=o.double_sided) for o in objs]
    cp = jr.CameraParameters(**{k: v.to(DEV) for k, v in camp._asdict().items()})._replace(viewWidth=84, viewHeight=84, vfov=58.0)
    cam = jr.Renderer.create_camera_from_parameters(cp)
    m = jr.merge_objects(objs_d)
    audit = pipeline.audit_cull(cam, m.faces, m.verts)
    print("  brax ant fixture 84x84:", audit)
    assert audit["filter_lost_triangles"] == 0 and audit["pixels_outside_bbox"] == 0 and audit["pixels_of_rejected_triangles"] == 0
