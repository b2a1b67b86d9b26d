"""Phong + tangent-space normal map (``renderer/shaders/phong_darboux.py:44-316``)."""
from typing import Any, NamedTuple

from .. import _native
from ..shader import Shader, _stage
from ..types import LightSource


class PhongTextureDarbouxExtraInput(NamedTuple):
    position: Any       # (V, 3)
    normal: Any         # (V, 3)
    uv: Any             # (V, 2) texel units
    light: LightSource
    texture: Any        # (Wt, Ht, 3)
    normal_map: Any     # (Wt, Ht, 3) Darboux frame
    id_to_face: Any     # (V,) face each vertex belongs to
    faces_indices: Any  # (F, 3)


class PhongTextureDarbouxExtraFragmentData(NamedTuple):
    normal: Any = (0.0, 0.0, 0.0)
    uv: Any = (0.0, 0.0)
    triangle: Any = None
    triangle_uv: Any = None
    colour: Any = (0.0, 0.0, 0.0)


class PhongTextureDarbouxExtraMixerOutput(NamedTuple):
    canvas: Any


class PhongTextureDarbouxShader(Shader):
    _jr_shader = _native.JR_PHONG_DARBOUX
    vertex = _stage("phong_darboux_vertex")
    interpolate = _stage("phong_darboux_interpolate")
    fragment = _stage("phong_darboux_fragment")
    mix = _stage("phong_darboux_mix")
