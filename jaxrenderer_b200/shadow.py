"""Shadow-map pass (``renderer/shadow.py:27-153``).

``render_shadow_map`` is pass 1 of ``phong_reflection_shadow``: an orthographic
light camera + the depth kernel + ``+ offset``.  The lookup ``Shadow.get``
(:129-153) is fused into the S7 shading kernel (``k_shade<6>``); the method
here is the host twin kept for API parity.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, NamedTuple

import torch

from . import _native
from .geometry import Camera, camera_build_native
from .types import Tensor, _f32


class ConstantFill(NamedTuple):
    """A shadow map that is still a constant (what ``Renderer.render`` passes: the largest float everywhere), not
    materialised: lets the depth kernel write the fill itself."""

    shape: Any
    value: float
    device: Any


class Shadow(NamedTuple):
    """``shadow.py:27-38``."""

    shadow_map: Tensor
    strength: Any
    camera: Camera

    @staticmethod
    def render_shadow_map(shadow_map: Tensor, verts: Tensor, faces: Tensor, light_direction: Any,
                          viewport_matrix: Tensor, centre: Any, up: Any, strength: Any,
                          offset: float = 0.001, distance: float = 10.0,
                          loop_unroll: int = 1) -> "Shadow":
        """``shadow.py:49-125``.  NB the light direction is used as given
        (un-normalised), exactly like the reference (``renderer.py:357``)."""
        dev = shadow_map.device if isinstance(shadow_map, (torch.Tensor, ConstantFill)) else None
        if dev is not None and not isinstance(dev, torch.device):
            dev = torch.device(dev)
        centre = _f32(centre, dev)
        ld = _f32(light_direction, dev)
        up = _f32(up, dev)
        cam = None
        if dev is not None and dev.type == "cuda":   # one launch instead of ~100 framework ops
            cam = camera_build_native(_native.JR_CAMERA_LIGHT, (
                (centre, 3), (ld, 3), (up, 3), (float(distance), 1), (-1.0, 1), (1.0, 1), (-1.0, 1), (1.0, 1),
                (-1.0, 1), (1.0, 1)), dev, viewport=viewport_matrix)
        if cam is None:
            cam = Shadow._light_camera(centre, ld, up, distance, viewport_matrix, dev)
        arrays = {"world_to_clip": cam.world_to_clip, "viewport": cam.viewport,
                  "position": verts, "faces": faces}
        return Shadow._finish(cam, arrays, shadow_map, strength, offset)

    @staticmethod
    def _light_camera(centre: Tensor, ld: Tensor, up: Tensor, distance: float, viewport_matrix: Tensor,
                      dev: Any) -> Camera:
        """Torch form of the light camera (``shadow.py:73-98``); differentiable / host inputs."""
        eye = centre + ld * distance
        view = Camera.view_matrix(eye=eye, centre=centre, up=up)
        proj = Camera.orthographic_projection_matrix(-1.0, 1.0, -1.0, 1.0, -1.0, 1.0).to(view.device)
        # The reference inverts `view` numerically here (Camera.create without view_inv); the analytic
        # inverse only enters fields of the light camera that the path never reads.
        return Camera.create(view=view, projection=proj, viewport=_f32(viewport_matrix, dev),
                             view_inv=Camera.view_matrix_inv(eye=eye, centre=centre, up=up))

    @staticmethod
    def _finish(cam: Camera, arrays: dict, shadow_map: Any, strength: Any, offset: float) -> "Shadow":
        from .pipeline import _render_arrays  # local: pipeline imports this module's users

        if isinstance(shadow_map, ConstantFill):
            # `Renderer.render`'s map (renderer.py:349-354: filled with the largest float): ONE launch -- the depth
            # kernel writes `z + offset` where a triangle covers the pixel and `fill + offset` elsewhere
            # (JrRenderArgs.depth_offset / depth_fill), no fill pass before, no `+ offset` pass after.
            z = torch.empty(shadow_map.shape, dtype=torch.float32, device=shadow_map.device)
            z, _, _ = _render_arrays(_native.JR_DEPTH, arrays, z, None, inplace=True,
                                     depth_epilogue=(float(offset), float(shadow_map.value)))
            return Shadow(shadow_map=z, strength=strength, camera=cam)
        z, _, _ = _render_arrays(_native.JR_DEPTH, arrays, shadow_map, None, inplace=False)
        if z.is_cuda:
            lib = _native.load()
            with torch.cuda.device(z.device):
                _native.check(lib.jr_add_scalar(z.data_ptr(), z.numel(), C.c_float(float(offset)),
                                                _native.stream_ptr(z.device)))
        else:  # host buffers came back from the device already; tiny epilogue
            z = z + float(offset)
        return Shadow(shadow_map=z, strength=strength, camera=cam)

    def get(self, position: Tensor) -> Tensor:
        """Host twin of the in-kernel lookup (``shadow.py:129-153``): round half
        away from zero, one negative wrap, out of bounds -> +inf."""
        r = torch.trunc(position)
        pos = (r + torch.where((position - r).abs() >= 0.5, torch.sign(position),
                               torch.zeros_like(position))).to(torch.int64)
        n0, n1 = self.shadow_map.shape[-2:]
        u = torch.where(pos[..., 0] < 0, pos[..., 0] + n0, pos[..., 0])
        v = torch.where(pos[..., 1] < 0, pos[..., 1] + n1, pos[..., 1])
        ok = (u >= 0) & (u < n0) & (v >= 0) & (v < n1)
        val = self.shadow_map[u.clamp(0, n0 - 1), v.clamp(0, n1 - 1)]
        return torch.where(ok, val, torch.full_like(val, float("inf")))
