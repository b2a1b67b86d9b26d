"""Facade (``renderer/renderer.py``): ``CameraParameters`` (:46-73),
``LightParameters`` (:76-109), ``ShadowParameters`` (:112-133) and
``Renderer.{create_camera_from_parameters, create_buffers, render,
get_camera_image}`` (:141-476).

Host glue only: it assembles the shader inputs and calls the CUDA path
(``pipeline._render_arrays``).  Unlike the reference it does NOT materialise
the corner-expanded ``position/normal/uv`` copies (``renderer.py:277-296``):
the C ABI takes the model's three index buffers directly, which is
arithmetically identical (gather-then-compute == compute-on-gathered).
All leaves may carry one leading batch axis.
"""
from __future__ import annotations

from typing import Any, NamedTuple, Optional, Sequence, Tuple, Union

import torch

from . import _native
from .geometry import Camera, _deg_tan_half, _normalise_last, camera_build_native
from .model import MergedModel, ModelObject, merge_objects
from .pipeline import _render_arrays
from .shadow import ConstantFill, Shadow
from .types import Buffers, DtypeInfo, Tensor, _f32


class CameraParameters(NamedTuple):
    """``renderer.py:46-73`` (defaults of tinyrenderer's camera)."""

    viewWidth: Any = 640
    viewHeight: Any = 480
    viewDepth: Any = 1.0
    near: Any = 0.01
    far: Any = 1000.0
    hfov: Any = 58.0
    vfov: Any = 45.0
    position: Any = (1.0, 1.0, 1.0)
    target: Any = (0.0, 0.0, 0.0)
    up: Any = (0.0, 0.0, 1.0)


_DIR = 0.57735 / (3 * 0.57735 ** 2) ** 0.5


class LightParameters(NamedTuple):
    """``renderer.py:76-109``."""

    direction: Any = (_DIR, _DIR, _DIR)
    colour: Any = (1.0, 1.0, 1.0)
    ambient: Any = (0.6, 0.6, 0.6)
    diffuse: Any = (0.35, 0.35, 0.35)
    specular: Any = (0.05, 0.05, 0.05)


class ShadowParameters(NamedTuple):
    """``renderer.py:112-133``."""

    centre: Any = (0.0, 0.0, 0.0)
    up: Any = (0.0, 0.0, 1.0)
    strength: Any = (0.6, 0.6, 0.6)
    offset: float = 0.05


class Renderer:
    @staticmethod
    def create_camera_from_parameters(camera: CameraParameters, device: Any = None) -> Camera:
        """``renderer.py:141-196``.  On a CUDA device (``device`` argument or where ``position`` lives)
        all 8 matrices come from one ``jr_camera_build`` launch; the torch builders below are the
        differentiable / host form of the same formulas."""
        dev = torch.device(device) if device is not None else (
            camera.position.device if isinstance(camera.position, torch.Tensor) else None)
        if dev is not None and dev.type == "cuda":
            cam = camera_build_native(_native.JR_CAMERA_PERSPECTIVE, (
                (camera.position, 3), (camera.target, 3), (camera.up, 3), (camera.vfov, 1), (camera.hfov, 1),
                (camera.near, 1), (camera.far, 1), (camera.viewWidth, 1), (camera.viewHeight, 1),
                (camera.viewDepth, 1)), dev)
            if cam is not None:
                return cam
        eye, centre, up = _f32(camera.position, device), _f32(camera.target, device), _f32(camera.up, device)
        view = Camera.view_matrix(eye=eye, centre=centre, up=up)
        view_inv = Camera.view_matrix_inv(eye=eye, centre=centre, up=up)
        dev = view.device
        proj = Camera.perspective_projection_matrix(
            fovy=_f32(camera.vfov, dev),
            aspect=_deg_tan_half(_f32(camera.hfov, dev)) / _deg_tan_half(_f32(camera.vfov, dev)),
            z_near=_f32(camera.near, dev), z_far=_f32(camera.far, dev))
        dim = torch.stack(torch.broadcast_tensors(_f32(camera.viewWidth, dev), _f32(camera.viewHeight, dev)), -1)
        viewport = Camera.viewport_matrix(lowerbound=torch.zeros(2, device=dev), dimension=dim,
                                          depth=_f32(camera.viewDepth, dev))
        return Camera.create(view=view, projection=proj, viewport=viewport, view_inv=view_inv)

    @staticmethod
    def create_buffers(width: int, height: int, batch: Optional[int] = None,
                       colour_default: Any = (1.0, 1.0, 1.0), zbuffer_default: Any = 1.0,
                       device: Any = None) -> Buffers:
        """``renderer.py:201-243``."""
        b = (batch,) if batch is not None else ()
        colour_default = _f32(colour_default, device)
        z = torch.full((*b, width, height), float(zbuffer_default), dtype=torch.float32, device=device)
        c = torch.empty((*b, width, height, colour_default.numel()), dtype=torch.float32,
                        device=colour_default.device).copy_(colour_default)   # never aliases the constant
        return Buffers(zbuffer=z, targets=(c,))

    @classmethod
    def render(cls, model: MergedModel, light: "LightParameters", camera: Camera, buffers: Buffers,
               shadow_param: Optional[ShadowParameters] = None, loop_unroll: int = 1,
               *, inplace: bool = False, display_uint8: Optional[Sequence[float]] = None) -> Buffers:
        """``renderer.py:254-385``: phong_reflection, or shadow-map pass +
        phong_reflection_shadow when ``shadow_param`` is given.

        ``display_uint8=(r, g, b)`` (extension): the target returned is the display image
        ``(B?, H, W, 3) uint8`` -- ``transpose_for_display((clip(canvas, 0, 1) * 255).astype(uint8))``, what every
        published benchmark of the reference computes next (``notebooks/32x32/A100.ipynb:326-337``) -- written by
        the shading kernel itself; the fp32 canvas is never materialised and ``buffers.targets`` may be ``()``
        (background = the given colour) or hold the incoming canvas (read only)."""
        del loop_unroll
        dev = buffers.zbuffer.device
        ldir_raw = _f32(light.direction, dev)
        light_dir = _normalise_last(ldir_raw)
        view = camera.view.to(dev)
        # Camera.apply_vec(light_dir, camera.view) (renderer.py:301-304), batch-aware
        lde = _normalise_last(light_dir)
        lde = (lde.unsqueeze(-2) @ view[..., :3, :3].transpose(-1, -2)).squeeze(-2)
        light_dir_eye = _normalise_last(lde)
        arrays = {
            "world_to_clip": camera.world_to_clip, "viewport": camera.viewport,
            "world_to_eye_norm": camera.world_to_eye_norm,
            "position": model.verts, "faces": model.faces,
            "normal": model.norms, "faces_norm": model.faces_norm,
            "uv": model.uvs, "faces_uv": model.faces_uv,
            "texture_index": model.texture_index, "faces_tex": model.faces_uv,
            "light_colour": _f32(light.colour, dev), "light_dir_eye": light_dir_eye,
            "ambient": _f32(light.ambient, dev), "diffuse": _f32(light.diffuse, dev),
            "specular": _f32(light.specular, dev),
            "texture": model.diffuse_map, "specular_map": model.specular_map,
            "texture_shape": model.texture_shape, "texture_offset": int(model.offset),
        }
        canvas = buffers.targets[0] if len(buffers.targets) else None
        if canvas is None and display_uint8 is None:
            raise ValueError("Renderer.render needs a canvas in buffers.targets (or display_uint8=background)")
        if shadow_param is None:
            z, c, _ = _render_arrays(_native.JR_PHONG_REFLECTION, arrays, buffers.zbuffer, canvas,
                                     inplace=inplace, display_u8=display_uint8)
            return Buffers(zbuffer=z, targets=(c,))
        zb = buffers.zbuffer
        fill = DtypeInfo.create(zb.dtype).max
        shadow = Shadow.render_shadow_map(
            shadow_map=(ConstantFill(tuple(zb.shape), fill, zb.device) if zb.is_cuda else torch.full_like(zb, fill)),
            # no gradient reaches the shadow map (it is only compared against, shadow.py:129-153)
            verts=model.verts.detach(), faces=model.faces, light_direction=ldir_raw.detach(),
            viewport_matrix=camera.viewport.detach(), centre=shadow_param.centre, up=shadow_param.up,
            strength=shadow_param.strength, offset=float(shadow_param.offset))
        arrays.update({
            "shadow_map": shadow.shadow_map, "shadow_strength": _f32(shadow.strength, dev),
            "shadow_world_to_clip": shadow.camera.world_to_clip,
            "shadow_viewport": shadow.camera.viewport,
        })
        z, c, _ = _render_arrays(_native.JR_PHONG_REFLECTION_SHADOW, arrays, buffers.zbuffer, canvas,
                                 inplace=inplace, display_u8=display_uint8)
        return Buffers(zbuffer=z, targets=(c,))

    @classmethod
    def get_camera_image(cls, objects: Sequence[ModelObject], light: "LightParameters",
                         camera: Union[Camera, CameraParameters], width: int, height: int,
                         colour_default: Any = (1.0, 1.0, 1.0), zbuffer_default: Any = 1.0,
                         shadow_param: Optional[ShadowParameters] = None, loop_unroll: int = 1,
                         *, display_uint8: bool = False) -> Tensor:
        """``renderer.py:395-476``.  Batched inputs (leading axis on object
        transforms / camera parameters) give a batched canvas ``(B, W, H, 3)``.

        ``display_uint8=True`` (extension, CUDA): returns the display image ``(B?, H, W, 3) uint8`` instead
        (see ``render``); no fp32 canvas is allocated, filled or written."""
        model = merge_objects(objects)
        dev = model.verts.device
        cam = cls.create_camera_from_parameters(camera, dev) if isinstance(camera, CameraParameters) else camera
        batch = None
        for t in (model.verts, cam.world_to_clip, model.diffuse_map):
            base = 2 if t is not model.diffuse_map else 3
            if t.ndim == base + 1:
                batch = t.shape[0]
        if display_uint8:
            b = (batch,) if batch is not None else ()
            z = torch.full((*b, width, height), float(zbuffer_default), dtype=torch.float32, device=dev)
            bg = [float(v) for v in (colour_default.reshape(-1).tolist() if isinstance(colour_default, torch.Tensor)
                                     else colour_default)]
            out = cls.render(model=model, light=light, camera=cam, buffers=Buffers(z, ()), shadow_param=shadow_param,
                             inplace=True, display_uint8=bg)
            return out.targets[0]
        buffers = cls.create_buffers(width, height, batch, colour_default, zbuffer_default, device=dev)
        out = cls.render(model=model, light=light, camera=cam, buffers=buffers,
                         shadow_param=shadow_param, loop_unroll=loop_unroll, inplace=dev.type == "cuda")
        return out.targets[0]
