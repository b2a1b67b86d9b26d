"""Gouraud shader (``renderer/shaders/gouraud.py:21-124``)."""
from typing import Any, NamedTuple

from .. import _native
from ..shader import Shader, _stage
from ..types import LightSource


class GouraudExtraInput(NamedTuple):
    position: Any  # (V, 3) world space
    colour: Any    # (V, 3)
    normal: Any    # (V, 3) world space
    light: LightSource


class GouraudExtraFragmentData(NamedTuple):
    colour: Any = (0.0, 0.0, 0.0)


class GouraudExtraMixerOutput(NamedTuple):
    canvas: Any


class GouraudShader(Shader):
    _jr_shader = _native.JR_GOURAUD
    vertex = _stage("gouraud_vertex")
    fragment = _stage("gouraud_fragment")
    mix = _stage("gouraud_mix")
