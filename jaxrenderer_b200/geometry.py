"""Host-side geometry helpers (torch), mirroring ``renderer/geometry.py``.

These are the cheap O(1)/O(V) data-preparation functions that stay in the host
framework ("Above the hot path").  Only ``Camera.to_clip``,
``apply_vec``, ``normalise`` and ``interpolate`` have device-side twins inside
the CUDA kernels (``csrc/jr_device.cuh``).

Every function accepts arbitrary leading batch axes where the reference takes
an un-batched value (the reference gets that from ``jax.vmap``).
"""
from __future__ import annotations

import enum
import math
from typing import Any, NamedTuple, Optional, Sequence, Tuple

import numpy as np

import torch

from .types import Tensor, _device_constant, _f32, _is_vmapped

View = Tensor
Projection = Tensor
Viewport = Tensor
World2Screen = Tensor


def normalise(vector: Tensor) -> Tensor:
    """``vector / ||vector||`` (``geometry.py:39-47``).

    NB the reference divides by the norm of the WHOLE array, also for batched
    ``(N, 3)`` input (Frobenius norm); kept as is.
    """
    vector = _f32(vector)
    return vector / torch.linalg.norm(vector)


def _normalise_last(vector: Tensor) -> Tensor:
    """Per-vector normalisation over the last axis (what ``vmap(normalise)``
    does in the reference)."""
    return vector / torch.linalg.norm(vector, dim=-1, keepdim=True)


class Interpolation(enum.Enum):
    """``geometry.py:50-110``."""

    FLAT = 0
    NOPERSPECTIVE = 1
    SMOOTH = 2

    def __call__(self, values: Tensor, barycentric_screen: Tensor, barycentric_clip: Tensor) -> Tensor:
        if self == Interpolation.FLAT:
            coef = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float32, device=values.device)
        elif self == Interpolation.NOPERSPECTIVE:
            coef = barycentric_screen
        else:
            coef = barycentric_clip
        return torch.tensordot(coef.to(torch.float32), values.to(torch.float32), dims=([0], [0]))


def interpolate(values: Tensor, barycentric_screen: Tensor, barycentric_clip: Tensor,
                mode: Interpolation = Interpolation.SMOOTH) -> Tensor:
    """``geometry.py:113-141``."""
    return mode(values, barycentric_screen, barycentric_clip)


def to_homogeneous(coordinates: Tensor, value: Any = 1.0) -> Tensor:
    """Append ``value`` on the last axis (``geometry.py:144-163``)."""
    coordinates = _f32(coordinates)
    pad = _f32(value, coordinates.device)
    pad = pad.reshape(pad.shape + (1,) * (coordinates.ndim - pad.ndim)) if pad.ndim else pad
    pad = pad.expand(*coordinates.shape[:-1], 1)
    return torch.cat((coordinates, pad), dim=-1)


def normalise_homogeneous(coordinates: Tensor) -> Tensor:
    """``geometry.py:166-180``."""
    return coordinates / coordinates[..., -1:]


def to_cartesian(coordinates: Tensor) -> Tensor:
    """``geometry.py:183-202``: drop ``w`` (no division when ``w == 0``)."""
    return torch.where(
        coordinates[..., -1:] == 0.0,
        coordinates[..., :-1],
        normalise_homogeneous(coordinates)[..., :-1],
    )


def _eye4(like: Tensor, batch: tuple) -> Tensor:
    return torch.eye(4, dtype=torch.float32, device=like.device).expand(*batch, 4, 4).clone()


class Camera(NamedTuple):
    """The 8 matrices of ``geometry.py:205-225``.  The pipeline reads only
    ``world_to_clip``, ``viewport`` and ``world_to_eye_norm``."""

    view: View
    projection: Projection
    viewport: Viewport
    world_to_clip: Projection
    world_to_eye_norm: Projection
    world_to_screen: World2Screen
    view_inv: View
    screen_to_world: World2Screen

    @classmethod
    def create(cls, view: View, projection: Projection, viewport: Viewport,
               view_inv: Optional[View] = None) -> "Camera":
        """``geometry.py:235-278``."""
        view, projection, viewport = _f32(view), _f32(projection), _f32(viewport)
        if view_inv is None:
            view_inv = torch.linalg.inv(view)
        is_persp = torch.isclose(projection[..., 3, 3], torch.zeros((), device=projection.device))
        projection_inv = torch.where(
            is_persp[..., None, None],
            cls.perspective_projection_matrix_inv(projection),
            cls.orthographic_projection_matrix_inv(projection),
        )
        viewport_inv = cls.viewport_matrix_inv(viewport)
        return cls(
            view=view,
            projection=projection,
            viewport=viewport,
            world_to_clip=projection @ view,
            world_to_eye_norm=view_inv.transpose(-1, -2),
            world_to_screen=viewport @ projection @ view,
            view_inv=view_inv,
            screen_to_world=view_inv @ projection_inv @ viewport_inv,
        )

    # ---- transforms -------------------------------------------------------
    @staticmethod
    def apply(points: Tensor, matrix: Tensor) -> Tensor:
        """``points @ matrix.T`` (``geometry.py:284-315``)."""
        return points.to(torch.float32) @ matrix.to(torch.float32).transpose(-1, -2)

    @classmethod
    def apply_pos(cls, points: Tensor, matrix: Tensor) -> Tensor:
        """``geometry.py:317-350``."""
        return to_cartesian(cls.apply(to_homogeneous(points), matrix))

    @classmethod
    def apply_vec(cls, vectors: Tensor, matrix: Tensor) -> Tensor:
        """``geometry.py:352-389`` (whole-array normalise, as the reference)."""
        n = normalise(vectors)
        t = cls.apply(to_homogeneous(n, 0.0), matrix)[..., :3]
        return normalise(t)

    def to_screen(self, points: Tensor) -> Tensor:
        return normalise_homogeneous(self.apply(points, self.world_to_screen))

    def to_clip(self, points: Tensor) -> Tensor:
        """``geometry.py:420-438``."""
        return self.apply(points, self.world_to_clip)

    # ---- matrix builders --------------------------------------------------
    @staticmethod
    def inv_scale_translation_matrix(m: Tensor) -> Tensor:
        """``geometry.py:472-511``."""
        d = torch.diagonal(m, dim1=-2, dim2=-1)
        scale_inv = torch.diag_embed(1.0 / d)
        translation = scale_inv @ m
        t_inv = _eye4(m, m.shape[:-2])
        t_inv[..., :3, 3] = -translation[..., :3, 3]
        return t_inv @ scale_inv

    @staticmethod
    def _look_at_basis(eye: Tensor, centre: Tensor, up: Tensor):
        forward = _normalise_last(centre - eye)
        up = _normalise_last(up)
        side = _normalise_last(torch.linalg.cross(forward, up, dim=-1))
        up = torch.linalg.cross(side, forward, dim=-1)
        return side, up, forward

    @classmethod
    def _look_at(cls, eye: Any, centre: Any, up: Any):
        eye, centre, up = _f32(eye), _f32(centre), _f32(up)
        eye, centre, up = torch.broadcast_tensors(eye, centre.to(eye.device), up.to(eye.device))
        side, up2, forward = cls._look_at_basis(eye, centre, up)
        rot = torch.stack((side, up2, -forward), dim=-2)          # (..., 3, 3), rows = camera axes
        return eye, rot

    @staticmethod
    def _affine(rot: Tensor, trans: Tensor) -> Tensor:
        """[[rot, trans], [0, 0, 0, 1]] built with two concatenations (few kernels)."""
        top = torch.cat((rot, trans.unsqueeze(-1)), dim=-1)
        bottom = torch.tensor((0.0, 0.0, 0.0, 1.0), device=rot.device).expand(*rot.shape[:-2], 1, 4)
        return torch.cat((top, bottom), dim=-2)

    @classmethod
    def view_matrix(cls, eye: Any, centre: Any, up: Any) -> View:
        """``lookAt`` (``geometry.py:536-575``): ``rotation @ translation(-eye)``."""
        eye, rot = cls._look_at(eye, centre, up)
        return cls._affine(rot, -(rot @ eye.unsqueeze(-1)).squeeze(-1))

    @classmethod
    def view_matrix_inv(cls, eye: Any, centre: Any, up: Any) -> View:
        """``geometry.py:577-636``: ``translation(eye) @ rotation^T``."""
        eye, rot = cls._look_at(eye, centre, up)
        return cls._affine(rot.transpose(-1, -2), eye)

    @staticmethod
    def perspective_projection_matrix(fovy: Any, aspect: Any, z_near: Any, z_far: Any) -> Projection:
        """``gluPerspective`` (``geometry.py:638-684``)."""
        fovy, aspect, z_near, z_far = torch.broadcast_tensors(
            _f32(fovy), _f32(aspect), _f32(z_near), _f32(z_far))
        f = 1.0 / torch.tan(torch.deg2rad(fovy) / 2.0)
        p = torch.zeros(*fovy.shape, 4, 4, dtype=torch.float32, device=fovy.device)
        p[..., 0, 0] = f / aspect
        p[..., 1, 1] = f
        p[..., 2, 2] = (z_far + z_near) / (z_near - z_far)
        p[..., 2, 3] = (2.0 * z_far * z_near) / (z_near - z_far)
        p[..., 3, 2] = -1.0
        return p

    @classmethod
    def perspective_projection_matrix_inv(cls, mat: Projection) -> Projection:
        """``geometry.py:686-718``."""
        shuffle = [0, 1, 3, 2]
        return cls.inv_scale_translation_matrix(mat[..., :, shuffle])[..., shuffle, :]

    @staticmethod
    def orthographic_projection_matrix(left: Any, right: Any, bottom: Any, top: Any,
                                       z_near: Any, z_far: Any) -> Projection:
        """``glOrtho`` (``geometry.py:720-763``)."""
        left, right, bottom, top, z_near, z_far = torch.broadcast_tensors(
            *[_f32(v) for v in (left, right, bottom, top, z_near, z_far)])
        l_op = torch.stack((right, top, z_far), dim=-1)
        r_op = torch.stack((left, bottom, z_near), dim=-1)
        p = torch.zeros(*left.shape, 4, 4, dtype=torch.float32, device=left.device)
        p[..., 0, 0] = 2 / (right - left)
        p[..., 1, 1] = 2 / (top - bottom)
        p[..., 2, 2] = -2 / (z_far - z_near)
        p[..., 3, 3] = 1
        p[..., :3, 3] = -(l_op + r_op) / (l_op - r_op)
        return p

    @classmethod
    def orthographic_projection_matrix_inv(cls, mat: Projection) -> Projection:
        return cls.inv_scale_translation_matrix(mat)

    @staticmethod
    def perspective_projection_matrix_tinyrenderer(eye: Any, centre: Any) -> Projection:
        """``geometry.py:783-811``."""
        eye, centre = _f32(eye), _f32(centre)
        p = _eye4(eye, eye.shape[:-1])
        p[..., 3, 2] = -1 / torch.linalg.norm(eye - centre, dim=-1)
        return p

    @staticmethod
    def viewport_matrix(lowerbound: Any, dimension: Any, depth: Any) -> Viewport:
        """``geometry.py:813-845``: NDC cube -> ``[x,x+w] x [y,y+h] x [0,d]``."""
        lowerbound, dimension, depth = _f32(lowerbound), _f32(dimension), _f32(depth)
        dimension = dimension.to(lowerbound.device) if lowerbound.is_cuda else dimension
        lowerbound = lowerbound.to(dimension.device)
        depth = depth.to(dimension.device)
        batch = torch.broadcast_shapes(lowerbound.shape[:-1], dimension.shape[:-1], depth.shape)
        v = _eye4(dimension, batch)
        v[..., :2, 3] = lowerbound + dimension / 2
        v[..., 0, 0] = dimension[..., 0] / 2
        v[..., 1, 1] = dimension[..., 1] / 2
        v[..., 2, 2] = depth / 2
        v[..., 2, 3] = depth / 2
        return v

    @classmethod
    def viewport_matrix_inv(cls, viewport: Viewport) -> Viewport:
        return cls.inv_scale_translation_matrix(viewport)

    @staticmethod
    def world_to_screen_matrix(width: int, height: int) -> World2Screen:
        """``geometry.py:863-882``."""
        a = torch.eye(4); a[0, 0] = 0.5; a[1, 1] = 0.5
        b = torch.eye(4); b[0, 0] = width; b[1, 1] = height
        c = torch.eye(4); c[:2, -1] = 1
        return a @ b @ c


def camera_build_native(mode: int, fields: Sequence[Tuple[Any, int]], dev: torch.device,
                        viewport: Optional[Tensor] = None) -> Optional[Camera]:
    """All 8 ``Camera`` matrices in ONE launch of ``jr_camera_build`` (``csrc/jr_camera.cu``) instead of
    ~120 framework ops.  ``fields`` are the (value, width) pairs of the 16-float parameter row described
    in ``include/jr_b200.h``; every value is un-batched or carries ONE leading batch axis.  Returns
    ``None`` when the fast path does not apply (deeper batch nesting, inside ``torch.func.vmap``): the caller
    then uses the torch builders above.  Differentiable: gradients flow through ``jr_camera_vjp``."""
    import ctypes as C

    from . import _native

    host = np.zeros(16, dtype=np.float32)
    dev_parts = []
    batch: Optional[int] = None
    tensors = [v for v, _ in fields] + [viewport]
    if _is_vmapped(*tensors):
        return None
    needs_grad = torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)
    off = 0
    for v, w in fields:
        base = 1 if w > 1 else 0
        nd = v.ndim if isinstance(v, torch.Tensor) else np.ndim(v)
        if nd > base + 1:
            return None
        if nd == base + 1 or (isinstance(v, torch.Tensor) and (v.is_cuda or v.requires_grad)):
            t = _f32(v, dev).reshape(-1, w)
            if t.shape[0] > 1:
                if batch not in (None, t.shape[0]):
                    raise ValueError(f"inconsistent batch sizes {batch} and {t.shape[0]}")
                batch = t.shape[0]
            dev_parts.append((off, w, t))
        else:
            host[off:off + w] = np.asarray(v, dtype=np.float32).reshape(-1)
        off += w
    assert off == 16
    vp = None
    if viewport is not None:
        vp = _f32(viewport, dev).contiguous()
        if vp.ndim > 3:
            return None
        if tuple(vp.shape[-2:]) != (4, 4):
            raise ValueError(f"viewport matrix must be (..., 4, 4), got {tuple(vp.shape)}")
        if vp.ndim == 3 and vp.shape[0] > 1:
            if batch not in (None, vp.shape[0]):
                raise ValueError(f"inconsistent batch sizes {batch} and {vp.shape[0]}")
            batch = vp.shape[0]
    B = batch or 1
    rows = _device_constant(host, dev if dev.index is not None else torch.device('cuda', torch.cuda.current_device()))
    if dev_parts:
        rows = rows.expand(B, 16).clone()
        for o, w, t in dev_parts:
            rows[:, o:o + w] = t
    if needs_grad:
        # differentiable: the same launch, with jr_camera_vjp (the kernel's own formulas on dual numbers) as its
        # reverse mode -- gradients w.r.t. CameraParameters / the light camera's inputs without the ~120 torch ops
        out = _CameraFn.apply(rows if rows.ndim == 2 else rows.expand(B, 16), vp, mode, B)
    else:
        out = _camera_launch(rows, vp, mode, B, dev)
    mats = out.unbind(0)
    if batch is None:
        mats = tuple(m[0] for m in mats)
    return Camera(*mats)


def _camera_args(rows: Tensor, vp: Optional[Tensor], mode: int, B: int):
    from . import _native

    a = _native.JrCameraArgs()
    a.B, a.mode = B, mode
    a.params = _native.JrF32(rows.data_ptr(), 16 if rows.ndim == 2 else 0)
    if vp is not None:
        a.viewport = _native.JrF32(vp.data_ptr(), 16 if (vp.ndim == 3 and vp.shape[0] > 1) else 0)
    return a


def _camera_launch(rows: Tensor, vp: Optional[Tensor], mode: int, B: int, dev: torch.device) -> Tensor:
    import ctypes as C

    from . import _native

    out = torch.empty((8, B, 4, 4), dtype=torch.float32, device=dev)
    a = _camera_args(rows, vp, mode, B)
    a.out = out.data_ptr()
    lib = _native.load()
    with torch.cuda.device(dev):
        _native.check(lib.jr_camera_build(C.byref(a), _native.stream_ptr(dev)))
    return out


class _CameraFn(torch.autograd.Function):
    """``jr_camera_build`` with ``jr_camera_vjp`` as its reverse mode."""

    @staticmethod
    def forward(rows: Tensor, vp: Optional[Tensor], mode: int, B: int):  # type: ignore[override]
        return _camera_launch(rows.contiguous(), vp, mode, B, rows.device)

    @staticmethod
    def setup_context(ctx: Any, inputs: Any, output: Any) -> None:
        rows, vp, mode, B = inputs
        ctx.mode, ctx.B = mode, B
        ctx.save_for_backward(rows, vp)

    @staticmethod
    def backward(ctx: Any, d_out: Tensor):  # type: ignore[override]
        import ctypes as C

        from . import _native

        rows, vp = ctx.saved_tensors
        rows = rows.contiguous()
        dev = rows.device
        d_out = d_out.contiguous()
        d_rows = torch.empty((ctx.B, 16), dtype=torch.float32, device=dev)
        want_vp = vp is not None and ctx.needs_input_grad[1] and ctx.mode == _native.JR_CAMERA_LIGHT
        d_vp = torch.empty((ctx.B, 4, 4), dtype=torch.float32, device=dev) if want_vp else None
        a = _camera_args(rows, vp, ctx.mode, ctx.B)
        lib = _native.load()
        with torch.cuda.device(dev):
            _native.check(lib.jr_camera_vjp(C.byref(a), d_out.data_ptr(), d_rows.data_ptr(),
                                            d_vp.data_ptr() if d_vp is not None else None, _native.stream_ptr(dev)))
        if d_vp is not None and not (vp.ndim == 3 and vp.shape[0] > 1):
            d_vp = d_vp.sum(0).reshape(vp.shape)          # shared viewport: the batch sum
        return d_rows, d_vp, None, None

    @staticmethod
    def vmap(info: Any, in_dims: Any, *args: Any):
        raise NotImplementedError("camera construction inside torch.func.vmap uses the torch builders")


def compute_normal(triangle_verts: Tensor) -> Tensor:
    """``geometry.py:885-900``."""
    n = torch.linalg.cross(triangle_verts[2] - triangle_verts[0],
                           triangle_verts[1] - triangle_verts[0])
    return n / torch.linalg.norm(n)


def quaternion(rotation_axis: Any, rotation_angle: Any) -> Tensor:
    """(w, x, y, z), angle in degrees (``geometry.py:903-932``)."""
    axis = normalise(_f32(rotation_axis))
    angle = torch.deg2rad(_f32(rotation_angle))
    return torch.cat((torch.cos(angle / 2).reshape(1), axis * torch.sin(angle / 2)))


def quaternion_mul(a: Tensor, b: Tensor) -> Tensor:
    """``geometry.py:935-966``."""
    a, b = _f32(a), _f32(b)
    return torch.stack((
        a[0] * b[0] - a[1:] @ b[1:],
        a[:3] @ b[[1, 0, 3]] - a[3] * b[2],
        a[[2, 3, 0]] @ b[:3] - a[1] * b[3],
        a[[0, 1, 3]] @ b[[3, 2, 0]] - a[2] * b[1],
    ))


def rotation_matrix(rotation_axis: Any, rotation_angle: Any) -> Tensor:
    """``geometry.py:969-1001``."""
    axis = normalise(_f32(rotation_axis))
    angle = torch.deg2rad(_f32(rotation_angle))
    c = torch.cos(angle)
    eye = torch.eye(3)
    cross = torch.linalg.cross(axis.expand(3, 3), eye, dim=-1)
    return eye * c - torch.sin(angle) * cross + (1 - c) * torch.outer(axis, axis)


def transform_matrix_from_rotation(rotation: Tensor) -> Tensor:
    """Quaternion (w, x, y, z) -> 3x3 (``geometry.py:1004-1037``; batch-aware)."""
    rotation = _f32(rotation)
    d = (rotation * rotation).sum(-1)
    s = 2.0 / d
    w, x, y, z = rotation.unbind(-1)
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    rows = (
        torch.stack((1.0 - (yy + zz), xy - wz, xz + wy), -1),
        torch.stack((xy + wz, 1.0 - (xx + zz), yz - wx), -1),
        torch.stack((xz - wy, yz + wx, 1.0 - (xx + yy)), -1),
    )
    return torch.stack(rows, -2)


def _deg_tan_half(deg: Any) -> Tensor:
    return torch.tan(torch.deg2rad(_f32(deg)) / 2.0)


__all__ = [
    "Camera", "Interpolation", "interpolate", "normalise", "normalise_homogeneous",
    "to_cartesian", "to_homogeneous", "quaternion", "quaternion_mul", "rotation_matrix",
    "transform_matrix_from_rotation", "compute_normal",
]
_ = math  # keep import (used by callers for constants)
