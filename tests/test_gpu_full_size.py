"""Parity at BASELINE.json's full sizes.

The brute-force oracle cannot render 4096 images in a test, so the full-size runs are checked
through size-independent properties (determinism, batch independence, untouched-pixel accounting)
plus exact comparison of a random sample of images against the C oracle (bit-equal to the torch
oracle, tests/test_oracle_c.py)."""
import numpy as np
import pytest
import torch

import jaxrenderer_b200 as jr
from jaxrenderer_b200 import synthetic
from jaxrenderer_b200.shaders import DepthExtraInput, DepthShader
from oracle import c_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cam_d(cam):
    return type(cam)(*[t.to(DEV) for t in cam])


def test_config2_full_batch_4096_depth_84():
    """configs[1] (the bench workload): 4096 x 84x84, 1932 triangles each."""
    B, W, H = 4096, 84, 84
    sc = synthetic.brax_like_batch(B, n_capsules=10)
    cam = synthetic.brax_cameras(sc["eye"], sc["target"], W, H)
    pos, faces, camd = sc["position"].to(DEV), sc["faces"].to(DEV), _cam_d(cam)

    def run():
        z = torch.full((B, W, H), 1.0, device=DEV)
        out, tri = jr.render(camd, DepthShader, jr.Buffers(z, ()), faces, DepthExtraInput(position=pos),
                             inplace=True, return_tri_id=True)
        return out.zbuffer, tri

    z1, t1 = run()
    z2, t2 = run()
    assert torch.equal(z1, z2) and torch.equal(t1, t2), "two runs must be bit-identical"
    # pixels without a triangle keep the incoming value, all others are overwritten
    assert bool((z1[t1 < 0] == 1.0).all())
    assert int((t1 >= 0).sum()) > 0.9 * t1.numel()       # the ground plane fills the view
    # batch independence: an image rendered alone equals its batch element
    for b in (0, 1234, B - 1):
        camb = camd._replace(world_to_clip=camd.world_to_clip[b])
        one = jr.render(camb, DepthShader, jr.Buffers(torch.full((W, H), 1.0, device=DEV), ()), faces[b],
                        DepthExtraInput(position=pos[b]))
        assert torch.equal(one.zbuffer, z1[b])
    # random sample against the brute-force C oracle: bit-exact z and triangle ids
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(B, size=24, replace=False))
    zo, to = c_oracle.render_depth(cam.world_to_clip[idx].numpy(), cam.viewport.numpy(), sc["position"][idx].numpy(),
                                   sc["faces"][idx].numpy(), np.ones((len(idx), W, H), np.float32))
    assert np.array_equal(t1[idx].cpu().numpy(), to)
    assert np.array_equal(z1[idx].cpu().numpy(), zo)
    # the BENCH variant (no triangle-id output -> z-only 32-bit keys) directly against the C oracle, and
    # against the 64-bit-key variant on all 4096 images
    z3 = torch.full((B, W, H), 1.0, device=DEV)
    jr.render(camd, DepthShader, jr.Buffers(z3, ()), faces, DepthExtraInput(position=pos), inplace=True)
    assert np.array_equal(z3[idx].cpu().numpy(), zo), "z-only-key kernel differs from the C oracle"
    assert torch.equal(z3, z1), "z-only-key kernel differs from the packed-key kernel"


def test_config2_counters_and_filter_audit():
    """The counting variant of the visibility kernel on the bench workload: N_test far below the reference's
    W*H*T, every fragment accounted for, and results unchanged."""
    from jaxrenderer_b200 import pipeline

    B, W, H = 256, 84, 84
    sc = synthetic.brax_like_batch(B, n_capsules=10, env0=5000)
    cam = synthetic.brax_cameras(sc["eye"], sc["target"], W, H)
    pos, faces, camd = sc["position"].to(DEV), sc["faces"].to(DEV), _cam_d(cam)
    z_plain = torch.full((B, W, H), 1.0, device=DEV)
    jr.render(camd, DepthShader, jr.Buffers(z_plain, ()), faces, DepthExtraInput(position=pos), inplace=True)
    z_cnt = torch.full((B, W, H), 1.0, device=DEV)
    with pipeline.visibility_stats(DEV) as st:
        jr.render(camd, DepthShader, jr.Buffers(z_cnt, ()), faces, DepthExtraInput(position=pos), inplace=True)
    c = st.read()
    print("visibility counters:", c)
    assert torch.equal(z_cnt, z_plain)
    T = faces.shape[1]
    assert c["triangles"] == B * T
    assert 0 < c["exact_kept"] <= c["filter_passed"] < c["triangles"]
    assert c["fragments"] <= c["n_test"] < B * W * H * T // 20     # reference: W*H*T tests per image
    assert c["fragments"] > 0


def test_config5_canvas_480x270_ant_3276_triangles():
    """configs[4] canvas and triangle count (binned path): 2 images against the C oracle."""
    B, W, H = 2, 480, 270
    sc = synthetic.brax_like_batch(B, n_capsules=17)
    cam = synthetic.brax_cameras(sc["eye"], sc["target"], W, H)
    out, tri = jr.render(_cam_d(cam), DepthShader, jr.Buffers(torch.full((B, W, H), 1.0, device=DEV), ()),
                         sc["faces"].to(DEV), DepthExtraInput(position=sc["position"].to(DEV)), return_tri_id=True)
    zo, to = c_oracle.render_depth(cam.world_to_clip.numpy(), cam.viewport.numpy(), sc["position"].numpy(),
                                   sc["faces"].numpy(), np.ones((B, W, H), np.float32))
    mism = int((tri.cpu().numpy() != to).sum())
    print("480x270 tri-id mismatches:", mism)
    assert mism == 0
    assert np.array_equal(out.zbuffer.cpu().numpy(), zo)


def test_config4_canvas_960x540_humanoid_19980_triangles():
    """configs[3] canvas and triangle count: one image, depth pass, against the C oracle
    (10.4 G pixel-triangle tests on the host cores)."""
    W, H = 960, 540
    sc = synthetic.brax_like_batch(1, n_capsules=104)
    assert sc["faces"].shape[1] == 19980
    cam = synthetic.brax_cameras(sc["eye"], sc["target"], W, H)
    out, tri = jr.render(_cam_d(cam), DepthShader, jr.Buffers(torch.full((1, W, H), 1.0, device=DEV), ()),
                         sc["faces"].to(DEV), DepthExtraInput(position=sc["position"].to(DEV)), return_tri_id=True)
    zo, to = c_oracle.render_depth(cam.world_to_clip.numpy(), cam.viewport.numpy(), sc["position"].numpy(),
                                   sc["faces"].numpy(), np.ones((1, W, H), np.float32))
    mism = int((tri.cpu().numpy() != to).sum())
    print("960x540 tri-id mismatches:", mism)
    assert mism == 0
    assert np.array_equal(out.zbuffer.cpu().numpy(), zo)


@pytest.mark.parametrize("n_caps", [3, 7, 17, 20])
def test_config3_gouraud_texture_32x32_mixed_envs(n_caps):
    """configs[2]: GouraudTextureShader at 32x32 over Brax-like environments of different sizes (T = 588 ... 3852,
    `32x32 A100 Various Envs.ipynb`): tri-ids, z and colours against the torch oracle for a few images per size,
    batch independence for the whole group."""
    from jaxrenderer_b200.shaders import GouraudTextureExtraInput, GouraudTextureShader
    from oracle import jr_oracle as O
    from tests.helpers import assert_parity, cam_at, compare

    B, W, H = 64, 32, 32
    sc = synthetic.brax_like_batch(B, n_capsules=n_caps, env0=7000 * n_caps, with_attributes=True)
    cam = synthetic.brax_cameras(sc["eye"], sc["target"], W, H)
    tex = synthetic.checker_texture()
    light = jr.LightSource(torch.tensor((0.57735, -0.57735, 0.57735)), torch.ones(3))
    uv = sc["uv"] * 100
    extra = GouraudTextureExtraInput(sc["position"].to(DEV), sc["normal"].to(DEV), uv.to(DEV),
                                     jr.LightSource(light.direction.to(DEV), light.colour.to(DEV)), tex.to(DEV))
    bufs = jr.Buffers(torch.full((B, W, H), 1.0, device=DEV), (torch.zeros(B, W, H, 3, device=DEV),))
    out, tri = jr.render(_cam_d(cam), GouraudTextureShader, bufs, sc["faces"].to(DEV), extra, return_tri_id=True)
    assert sc["faces"].shape[1] == 12 + 192 * n_caps
    for b in (0, 31, B - 1):
        ex_b = GouraudTextureExtraInput(sc["position"][b], sc["normal"][b], uv, light, tex)
        ref = O.render(cam_at(cam, b), "gouraud_texture", torch.full((W, H), 1.0), (torch.zeros(W, H, 3),),
                       sc["faces"][b], ex_b)
        rep = compare(f"cfg3 T={sc['faces'].shape[1]} b={b}", out.zbuffer[b], out.targets[0][b], tri[b], ref)
        assert_parity(rep)
    # batch independence
    b = 17
    one = jr.render(_cam_d(cam)._replace(world_to_clip=cam.world_to_clip[b].to(DEV)), GouraudTextureShader,
                    jr.Buffers(torch.full((W, H), 1.0, device=DEV), (torch.zeros(W, H, 3, device=DEV),)),
                    sc["faces"][b].to(DEV),
                    GouraudTextureExtraInput(extra.position[b], extra.normal[b], extra.uv, extra.light, extra.texture))
    assert torch.equal(one.zbuffer, out.zbuffer[b]) and torch.equal(one.targets[0], out.targets[0][b])
