"""Depth shader (``renderer/shaders/depth.py:20-61``)."""
from typing import Any, NamedTuple

from .. import _native
from ..shader import Shader, _stage


class DepthExtraInput(NamedTuple):
    position: Any  # (V, 3) world space


class DepthExtraFragmentData(NamedTuple):
    pass


class DepthExtraMixerOutput(NamedTuple):
    pass


class DepthShader(Shader):
    """Writes the z-buffer only; fused into ``k_visibility<true>``."""

    _jr_shader = _native.JR_DEPTH
    vertex = _stage("depth_vertex")
