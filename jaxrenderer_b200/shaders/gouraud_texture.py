"""Gouraud + texture shader (``renderer/shaders/gouraud_texture.py:22-167``)."""
from typing import Any, NamedTuple

from .. import _native
from ..shader import Shader, _stage
from ..types import LightSource


class GouraudTextureExtraInput(NamedTuple):
    position: Any  # (V, 3)
    normal: Any    # (V, 3)
    uv: Any        # (V, 2) in texel units
    light: LightSource
    texture: Any   # (Wt, Ht, 3)


class GouraudTextureExtraFragmentData(NamedTuple):
    colour: Any = (0.0, 0.0, 0.0)
    uv: Any = (0.0, 0.0)


class GouraudTextureExtraMixerOutput(NamedTuple):
    canvas: Any


class GouraudTextureShader(Shader):
    _jr_shader = _native.JR_GOURAUD_TEXTURE
    vertex = _stage("gouraud_texture_vertex")
    fragment = _stage("gouraud_texture_fragment")
    mix = _stage("gouraud_texture_mix")
