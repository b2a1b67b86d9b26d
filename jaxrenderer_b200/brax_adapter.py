"""Brax glue: system geometry + link states -> ``ModelObject`` list + ``CameraParameters``.

Restates cell 3 of the reference's ``notebooks/Generate Data.ipynb`` (``_build_objects``, ``_with_state``,
``_eye``, ``get_target``, ``get_camera``) without importing brax: colliders are described by the duck-typed
``Geom`` record below (the fields the notebook reads from ``brax.base.{Capsule,Box,Sphere,Plane,Mesh}``).
``with_state`` and ``get_camera`` are batch-native: link positions ``(..., L, 3)`` and rotations ``(..., L, 4)``
(quaternions ``w, x, y, z``) may carry leading environment axes, which flow straight into the batched renderer
(the notebook gets that from ``jax.vmap``).
"""
from __future__ import annotations

from typing import Any, List, NamedTuple, Optional, Sequence

import numpy as np
import torch

from .geometry import Camera, transform_matrix_from_rotation
from .model import Model, ModelObject
from .renderer import CameraParameters
from .shapes.capsule import UpAxis, create_capsule
from .shapes.cube import create_cube
from .types import Tensor, _f32


class Geom(NamedTuple):
    """One collider of a Brax system (``sys.geoms[*]`` element)."""

    kind: str                      # "capsule" | "box" | "sphere" | "plane" | "mesh" | "convex"
    link_idx: Optional[int]        # None = world
    pos: Any = (0.0, 0.0, 0.0)     # col.transform.pos (offset in the link frame)
    rot: Any = (1.0, 0.0, 0.0, 0.0)   # col.transform.rot
    rgba: Any = (0.5, 0.5, 0.5, 1.0)
    radius: Any = 0.0              # capsule / sphere
    length: Any = 0.0              # capsule (full length of the cylinder part)
    halfsize: Any = (1.0, 1.0, 1.0)   # box
    vert: Any = None               # mesh
    face: Any = None               # mesh
    vertex_normals: Any = None     # mesh (the notebook takes trimesh's)


class Obj(NamedTuple):
    """``Obj`` of the notebook: a renderable instance bound to a link."""

    instance: ModelObject
    link_idx: int
    off: Tensor
    rot: Tensor


def ground_texture(grid_size: int = 100, colour: Sequence[int] = (200, 200, 200)) -> Tensor:
    """``grid()`` of the notebook: flat colour with a black first row and last column."""
    g = np.zeros((grid_size, grid_size, 3), dtype=np.float32)
    g[:, :] = np.asarray(colour, dtype=np.float32) / 255.0
    g[0] = 0.0
    g[:, -1] = 0.0
    return torch.from_numpy(g)


def build_objects(geoms: Sequence[Geom], device: Any = None) -> List[Obj]:
    """``_build_objects``: one ``Model`` per visual collider (convex colliders are not visual)."""
    objs: List[Obj] = []
    # the notebook groups colliders by link (insertion order of first appearance), keeping order inside a link
    order: List[Optional[int]] = []
    for g in geoms:
        if g.link_idx not in order:
            order.append(g.link_idx)
    for link in order:
        for g in geoms:
            if g.link_idx != link:
                continue
            tex = _f32(g.rgba, device)[:3].reshape(1, 1, 3)
            spec = torch.full((1, 1), 2.0, dtype=torch.float32, device=device)
            if g.kind == "capsule":
                model = create_capsule(radius=_f32(g.radius, device), half_height=_f32(g.length, device) / 2,
                                       up_axis=UpAxis.Z, diffuse_map=tex, specular_map=spec)
            elif g.kind == "sphere":
                model = create_capsule(radius=_f32(g.radius, device), half_height=torch.zeros((), device=device),
                                       up_axis=UpAxis.Z, diffuse_map=tex, specular_map=spec)
            elif g.kind == "box":
                model = create_cube(half_extents=_f32(g.halfsize, device), texture_scaling=16.0,
                                    diffuse_map=tex, specular_map=spec)
            elif g.kind == "plane":
                tex = ground_texture().to(device) if device is not None else ground_texture()
                model = create_cube(half_extents=_f32((1000.0, 1000.0, 0.0001), device), texture_scaling=8192.0,
                                    diffuse_map=tex, specular_map=torch.full(tex.shape[:2], 2.0, device=device))
            elif g.kind == "convex":
                continue
            elif g.kind == "mesh":
                verts = _f32(g.vert, device)
                norms = _f32(g.vertex_normals, device) if g.vertex_normals is not None else _vertex_normals(
                    verts, torch.as_tensor(g.face, device=device).long())
                model = Model.create(verts=verts, norms=norms, uvs=torch.zeros(verts.shape[0], 2, device=device),
                                     faces=torch.as_tensor(g.face, device=device), diffuse_map=tex)
            else:
                raise RuntimeError(f"unrecognized collider: {g.kind}")
            objs.append(Obj(instance=ModelObject(model=model), link_idx=-1 if g.link_idx is None else int(g.link_idx),
                            off=_f32(g.pos, device), rot=_f32(g.rot, device)))
    return objs


def _vertex_normals(verts: Tensor, faces: Tensor) -> Tensor:
    """Area-weighted vertex normals (what ``trimesh.Trimesh.vertex_normals`` gives for a clean mesh)."""
    fn = torch.linalg.cross(verts[faces[:, 1]] - verts[faces[:, 0]], verts[faces[:, 2]] - verts[faces[:, 0]], dim=-1)
    out = torch.zeros_like(verts)
    for k in range(3):
        out.index_add_(0, faces[:, k], fn)
    return out / torch.linalg.norm(out, dim=-1, keepdim=True).clamp_min(1e-20)


def quat_mul(a: Tensor, b: Tensor) -> Tensor:
    """``brax.math.quat_mul`` on ``(..., 4)`` (w, x, y, z)."""
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz,
                        aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw), dim=-1)


def rotate(vec: Tensor, quat: Tensor) -> Tensor:
    """``brax.math.rotate``: rotate ``vec (..., 3)`` by the unit quaternion ``quat (..., 4)``."""
    s, u = quat[..., :1], quat[..., 1:]
    r = 2 * (u * vec).sum(-1, keepdim=True) * u + (s * s - (u * u).sum(-1, keepdim=True)) * vec
    return r + 2 * s * torch.linalg.cross(u, vec.expand_as(u), dim=-1)


def with_state(objs: Sequence[Obj], x_pos: Tensor, x_rot: Tensor) -> List[ModelObject]:
    """``_with_state``: place every object at its link's pose.  ``x_pos (..., L, 3)`` / ``x_rot (..., L, 4)`` are
    ``state.x`` WITHOUT the extra world entry -- it is appended here (``x.concatenate(Transform.zero((1,)))``),
    so ``link_idx == -1`` addresses the world frame.  Leading axes are environments."""
    x_pos, x_rot = _f32(x_pos), _f32(x_rot)
    batch = x_pos.shape[:-2]
    zero_p = torch.zeros(*batch, 1, 3, device=x_pos.device)
    zero_r = torch.tensor((1.0, 0.0, 0.0, 0.0), device=x_pos.device).expand(*batch, 1, 4)   # Transform.zero: identity
    pos_all = torch.cat((x_pos, zero_p), dim=-2)
    rot_all = torch.cat((x_rot, zero_r), dim=-2)
    instances: List[ModelObject] = []
    for o in objs:
        lp, lr = pos_all[..., o.link_idx, :], rot_all[..., o.link_idx, :]
        off, orot = o.off.to(lp.device), o.rot.to(lp.device)
        pos = lp + rotate(off, lr)
        rot = quat_mul(lr, orot.expand_as(lr))
        # replace_with_position + replace_with_orientation (model.py:365-399), batch-aware
        transform = Camera._affine(transform_matrix_from_rotation(rot), pos)
        instances.append(o.instance._replace(transform=transform))
    return instances


def get_target(x_pos: Tensor) -> Tensor:
    """``get_target``: the root link's x, y on the ground."""
    root = _f32(x_pos)[..., 0, :]
    return torch.stack((root[..., 0], root[..., 1], torch.zeros_like(root[..., 0])), dim=-1)


def eye(x_pos: Tensor, x_rot: Tensor, joint_pos: Tensor) -> Tensor:
    """``_eye``: root position + ``(2d, -2d, d)`` with ``d`` the largest distance between two joints
    (``joint_pos (L, 3)`` = ``sys.link.joint.pos``, expressed in world space through the link poses)."""
    x_pos, x_rot = _f32(x_pos), _f32(x_rot)
    xj = x_pos + rotate(_f32(joint_pos).to(x_pos.device).expand_as(x_pos), x_rot)
    d = torch.linalg.norm(xj[..., None, :, :] - xj[..., :, None, :], dim=-1).flatten(-2).max(dim=-1).values
    off = torch.stack((2 * d, -2 * d, d), dim=-1)
    return x_pos[..., 0, :] + off


def get_camera(x_pos: Tensor, x_rot: Tensor, joint_pos: Tensor, width: int = 960, height: int = 540) -> CameraParameters:
    """``get_camera``: hfov 58, vfov scaled by the aspect ratio, up = +z, full-view viewport."""
    hfov = 58.0
    return CameraParameters(viewWidth=width, viewHeight=height, position=eye(x_pos, x_rot, joint_pos),
                            target=get_target(x_pos), up=torch.tensor((0.0, 0.0, 1.0), device=_f32(x_pos).device),
                            hfov=hfov, vfov=hfov * height / width)


__all__ = ["Geom", "Obj", "build_objects", "with_state", "get_camera", "get_target", "eye", "ground_texture",
           "quat_mul", "rotate"]
