"""Phong reflection model + texture atlas (``renderer/shaders/phong_reflection.py:35-258``)."""
from typing import Any, NamedTuple

from .. import _native
from ..shader import Shader, _stage
from ..types import LightSource


class PhongReflectionTextureExtraInput(NamedTuple):
    position: Any        # (V, 3)
    normal: Any          # (V, 3)
    uv: Any              # (V, 2) in [0, 1], repeat
    light: LightSource
    light_dir_eye: Any   # (3,)
    texture_shape: Any   # (objects, 2) int
    texture_index: Any   # (V,) int
    texture_offset: Any  # int
    texture: Any         # atlas (sum W, Hmax, 3)
    specular_map: Any    # atlas (sum Ws, Hs)
    ambient: Any
    diffuse: Any
    specular: Any


class PhongReflectionTextureExtraFragmentData(NamedTuple):
    normal: Any = (0.0, 0.0, 0.0)
    uv: Any = (0.0, 0.0)
    texture_index: Any = 0
    colour: Any = (0.0, 0.0, 0.0)


class PhongReflectionTextureExtraMixerOutput(NamedTuple):
    canvas: Any


class PhongReflectionTextureShader(Shader):
    _jr_shader = _native.JR_PHONG_REFLECTION
    vertex = _stage("phong_reflection_vertex")
    interpolate = _stage("phong_reflection_interpolate")
    fragment = _stage("phong_reflection_fragment")
    mix = _stage("phong_reflection_mix")
