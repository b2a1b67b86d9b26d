"""Brax glue: host logic on CPU, one end-to-end render on the GPU."""
import pytest
import torch

import jaxrenderer_b200 as jr
from jaxrenderer_b200 import brax_adapter as BA
from jaxrenderer_b200.geometry import transform_matrix_from_rotation


def _unit_quats(n, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n, 4, generator=g)
    return q / q.norm(dim=-1, keepdim=True)


def _ant_like():
    """Ground plane + torso sphere + 4 legs of 2 capsules: 9 links, world plane."""
    geoms = [BA.Geom("plane", None), BA.Geom("sphere", 0, radius=0.25, rgba=(0.8, 0.6, 0.4, 1.0))]
    for leg in range(4):
        geoms.append(BA.Geom("capsule", 1 + 2 * leg, pos=(0.1, 0.0, 0.0), rot=(0.7071068, 0.0, 0.7071068, 0.0),
                             radius=0.08, length=0.28, rgba=(0.8, 0.6, 0.4, 1.0)))
        geoms.append(BA.Geom("capsule", 2 + 2 * leg, pos=(0.2, 0.0, 0.0), rot=(0.7071068, 0.0, 0.7071068, 0.0),
                             radius=0.08, length=0.56, rgba=(0.8, 0.6, 0.4, 1.0)))
    geoms.append(BA.Geom("convex", 0))          # not visual
    geoms.append(BA.Geom("box", 0, halfsize=(0.05, 0.05, 0.05)))
    return geoms


def test_quaternion_helpers_match_rotation_matrices():
    q, p = _unit_quats(16, 0), _unit_quats(16, 1)
    v = torch.randn(16, 3, generator=torch.Generator().manual_seed(2))
    R = transform_matrix_from_rotation(q)
    torch.testing.assert_close(BA.rotate(v, q), (R @ v[..., None])[..., 0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(transform_matrix_from_rotation(BA.quat_mul(q, p)),
                               R @ transform_matrix_from_rotation(p), rtol=1e-5, atol=1e-6)


def test_build_objects_and_with_state():
    objs = BA.build_objects(_ant_like())
    # grouped by link in order of first appearance; convex skipped; box of link 0 follows the sphere
    assert [o.link_idx for o in objs] == [-1, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8]
    assert objs[0].instance.model.verts.shape == (24, 3) and objs[0].instance.model.diffuse_map.shape == (100, 100, 3)
    assert objs[1].instance.model.verts.shape == (576, 3) and objs[2].instance.model.verts.shape == (24, 3)
    # sphere = capsule with half_height 0: all vertices at distance radius
    torch.testing.assert_close(objs[1].instance.model.verts.norm(dim=-1), torch.full((576,), 0.25), rtol=2e-2, atol=0.0)  # table is ~1.5 % off a sphere
    B, L = 5, 9
    x_pos = torch.randn(B, L, 3, generator=torch.Generator().manual_seed(3))
    x_rot = _unit_quats(B * L, 4).reshape(B, L, 4)
    inst = BA.with_state(objs, x_pos, x_rot)
    assert len(inst) == len(objs) and all(i.transform.shape == (B, 4, 4) for i in inst)
    # world object: identity; linked object: link pose composed with the collider's local transform
    torch.testing.assert_close(inst[0].transform, torch.eye(4).expand(B, 4, 4))
    o = objs[4]
    Rl = transform_matrix_from_rotation(x_rot[:, o.link_idx])
    want_R = Rl @ transform_matrix_from_rotation(o.rot)
    want_t = x_pos[:, o.link_idx] + (Rl @ o.off[:, None])[..., 0]
    torch.testing.assert_close(inst[4].transform[:, :3, :3], want_R, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(inst[4].transform[:, :3, 3], want_t, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(inst[4].transform[:, 3], torch.tensor((0.0, 0.0, 0.0, 1.0)).expand(B, 4))
    # un-batched state works too
    assert BA.with_state(objs, x_pos[0], x_rot[0])[3].transform.shape == (4, 4)


def test_camera_rule():
    L = 3
    x_pos = torch.tensor([[0.5, -0.25, 0.6], [1.5, -0.25, 0.6], [0.5, 1.75, 0.6]])
    x_rot = torch.tensor([[1.0, 0.0, 0.0, 0.0]]).expand(L, 4)
    cam = BA.get_camera(x_pos, x_rot, torch.zeros(L, 3), width=84, height=84)
    d = 5.0 ** 0.5                      # farthest pair of joints: (1.5, -0.25) <-> (0.5, 1.75)
    torch.testing.assert_close(cam.position, x_pos[0] + torch.tensor((2 * d, -2 * d, d)))
    torch.testing.assert_close(cam.target, torch.tensor((0.5, -0.25, 0.0)))
    assert cam.hfov == 58.0 and cam.vfov == 58.0 and cam.viewWidth == 84
    camb = BA.get_camera(x_pos.expand(7, L, 3), x_rot.expand(7, L, 4), torch.zeros(L, 3), 84, 84)
    assert camb.position.shape == (7, 3) and camb.target.shape == (7, 3)


@pytest.mark.gpu
def test_brax_state_to_pixels_on_gpu():
    dev = torch.device("cuda", 0)
    objs = BA.build_objects(_ant_like(), device=dev)
    B, L = 6, 9
    g = torch.Generator().manual_seed(5)
    x_pos = torch.randn(B, L, 3, generator=g) * 0.3 + torch.tensor((0.0, 0.0, 0.7))
    x_rot = _unit_quats(B * L, 6).reshape(B, L, 4)
    inst = BA.with_state(objs, x_pos.to(dev), x_rot.to(dev))
    cam = BA.get_camera(x_pos.to(dev), x_rot.to(dev), torch.zeros(L, 3), 84, 84)
    img = jr.Renderer.get_camera_image(inst, jr.LightParameters(), cam, 84, 84,
                                       shadow_param=jr.ShadowParameters(centre=cam.target))
    assert img.shape == (B, 84, 84, 3) and bool(torch.isfinite(img).all())
    background = (img == 1.0).all(-1)
    assert 0.02 < float((~background).float().mean()) <= 1.0     # the ground and the robot are visible
    # per-environment result equals the un-batched call
    one = jr.Renderer.get_camera_image(BA.with_state(objs, x_pos[2].to(dev), x_rot[2].to(dev)), jr.LightParameters(),
                                       BA.get_camera(x_pos[2].to(dev), x_rot[2].to(dev), torch.zeros(L, 3), 84, 84),
                                       84, 84, shadow_param=jr.ShadowParameters(centre=cam.target[2]))
    torch.testing.assert_close(one, img[2], rtol=1e-5, atol=1e-6)
