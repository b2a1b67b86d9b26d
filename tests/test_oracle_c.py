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
