"""Phong + texture shader (``renderer/shaders/phong.py:22-187``)."""
from typing import Any, NamedTuple

from .. import _native
from ..shader import Shader, _stage
from ..types import LightSource


class PhongTextureExtraInput(NamedTuple):
    position: Any  # (V, 3)
    normal: Any    # (V, 3)
    uv: Any        # (V, 2) in texel units
    light: LightSource  # direction is used as-is in eye space (head-light)
    texture: Any   # (Wt, Ht, 3)


class PhongTextureExtraFragmentData(NamedTuple):
    normal: Any = (0.0, 0.0, 0.0)
    uv: Any = (0.0, 0.0)
    colour: Any = (0.0, 0.0, 0.0)


class PhongTextureExtraMixerOutput(NamedTuple):
    canvas: Any


class PhongTextureShader(Shader):
    _jr_shader = _native.JR_PHONG
    vertex = _stage("phong_vertex")
    fragment = _stage("phong_fragment")
    mix = _stage("phong_mix")
