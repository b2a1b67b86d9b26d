"""The C oracle (bench CPU baseline) must equal the torch oracle bit for bit."""
from types import SimpleNamespace as NS

import numpy as np
import torch

from jaxrenderer_b200 import synthetic
from oracle import c_oracle
from oracle import jr_oracle as O
from tests.helpers import cam_at, smoke_scene


def test_c_oracle_equals_torch_oracle_brax_like():
    W, H, B = 40, 32, 2
    sc = synthetic.brax_like_batch(B, n_capsules=2)
    cam = synthetic.brax_cameras(sc["eye"], sc["target"], W, H)
    z0 = torch.full((B, W, H), 1.0)
    z, tri = c_oracle.render_depth(cam.world_to_clip.numpy(), cam.viewport.numpy(),
                                   sc["position"].numpy(), sc["faces"].numpy(), z0.numpy())
    for b in range(B):
        ref = O.render(cam_at(cam, b), "depth", z0[b], (), sc["faces"][b], NS(position=sc["position"][b]))
        assert np.array_equal(tri[b], ref.tri_id.numpy())
        assert np.array_equal(z[b], ref.zbuffer.numpy())


def test_c_oracle_triangle0_leak():
    W, H = 40, 36
    cam, _, extra = smoke_scene(W, H, depth=1.0)
    pos = torch.cat((extra.position, torch.tensor(((-2.0, -1.0, 0.0), (-1.0, -1.0, 0.0), (-1.5, -0.2, 0.3)))))
    faces = torch.tensor(((0, 2, 1), (6, 7, 8)), dtype=torch.int32)
    z0 = torch.full((W, H), 7.0)
    ref = O.render(cam, "depth", z0, (), faces, NS(position=pos))
    z, tri = c_oracle.render_depth(cam.world_to_clip.numpy(), cam.viewport.numpy(), pos.numpy(),
                                   faces.numpy(), z0.numpy())
    assert int(((ref.tri_id == 0) & ~ref.has).sum()) > 0
    assert np.array_equal(tri[0], ref.tri_id.numpy()) and np.array_equal(z[0], ref.zbuffer.numpy())


def test_c_visibility_equals_torch_visibility_any_shader():
    """``c_oracle.visibility`` (the visibility stage for ANY shader) returns the torch oracle's tuple bit for bit,
    and ``O.render(..., vis_fn=...)`` therefore the same image: Brax-like scene + a random soup with ties / empties."""
    from tests.helpers import random_mesh_scene

    W, H = 40, 32
    sc = synthetic.brax_like_batch(1, n_capsules=2, with_attributes=True)
    cam = cam_at(synthetic.brax_cameras(sc["eye"], sc["target"], W, H), 0)
    scenes = [(cam, sc["position"][0], sc["faces"][0], W, H)]
    s = random_mesh_scene(3, n_tri=40, W=36, H=28)
    scenes.append((s.cam, s.pos, s.faces, 36, 28))
    # duplicate triangles (exact depth ties -> gap 0) and a degenerate one
    dup_faces = torch.cat((s.faces[:5], s.faces[:5], torch.tensor([[0, 0, 1]], dtype=torch.int32), s.faces[5:]))
    scenes.append((s.cam, s.pos, dup_faces, 36, 28))
    for cam_i, pos, faces, w, h in scenes:
        clip_v = O.mat4_apply(pos, O._t(cam_i.world_to_clip), w_one=True)
        setup = O.primitive_setup(clip_v, faces.long())
        idx, has, kc, gap = O.visibility(setup, O._t(cam_i.viewport), w, h)
        cidx, chas, ckc, cgap = c_oracle.visibility(cam_i.world_to_clip, cam_i.viewport, pos, faces, w, h)
        assert torch.equal(idx, cidx) and torch.equal(has, chas) and torch.equal(kc, ckc)
        assert torch.equal(gap, cgap)          # inf == inf, bit-equal finite gaps
    # end to end through a shaded render
    light = NS(direction=torch.tensor((0.3, 0.5, 0.8)), colour=torch.tensor((1.0, 0.9, 0.8)))
    extra = NS(position=s.pos, normal=s.nrm, colour=s.col, light=light)
    a = O.render(s.cam, "gouraud", torch.full((36, 28), 1.0), (torch.zeros(36, 28, 3),), s.faces, extra)
    b = O.render(s.cam, "gouraud", torch.full((36, 28), 1.0), (torch.zeros(36, 28, 3),), s.faces, extra,
                 vis_fn=c_oracle.visibility)
    assert torch.equal(a.tri_id, b.tri_id) and torch.equal(a.zbuffer, b.zbuffer)
    assert torch.equal(a.targets[0], b.targets[0])
