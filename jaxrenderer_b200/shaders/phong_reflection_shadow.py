"""Phong reflection + shadow map (``renderer/shaders/phong_reflection_shadow.py:36-295``)."""
from typing import Any, NamedTuple

from .. import _native
from ..shader import Shader, _stage
from ..types import LightSource


class PhongReflectionShadowTextureExtraInput(NamedTuple):
    position: Any
    normal: Any
    uv: Any
    light: LightSource
    light_dir_eye: Any
    texture_shape: Any
    texture_index: Any
    texture_offset: Any
    texture: Any
    specular_map: Any
    shadow: Any   # shadow.Shadow
    camera: Any   # geometry.Camera
    ambient: Any
    diffuse: Any
    specular: Any


class PhongReflectionShadowTextureExtraFragmentData(NamedTuple):
    normal: Any = (0.0, 0.0, 0.0)
    uv: Any = (0.0, 0.0)
    texture_index: Any = 0
    shadow_coord: Any = (0.0, 0.0, 0.0, 1.0)
    colour: Any = (0.0, 0.0, 0.0)


class PhongReflectionShadowTextureExtraMixerOutput(NamedTuple):
    canvas: Any


class PhongReflectionShadowTextureShader(Shader):
    _jr_shader = _native.JR_PHONG_REFLECTION_SHADOW
    vertex = _stage("phong_reflection_shadow_vertex")
    interpolate = _stage("phong_reflection_interpolate")
    fragment = _stage("phong_reflection_shadow_fragment")
    mix = _stage("phong_reflection_shadow_mix")
