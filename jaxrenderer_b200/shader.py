"""The ``Shader`` plugin protocol of the reference (``renderer/shader.py:42-396``).

The five stage names (``vertex``, ``primitive_chooser``, ``interpolate``,
``fragment``, ``mix``) and the carrier tuples ``PerVertex`` / ``PerFragment`` /
``MixerOutput`` are kept so code written against the reference imports
unchanged.  In this implementation the stages of the seven built-in shaders
are FUSED into hand-written CUDA kernels (``csrc/jr_forward.cu``); they are
never executed as Python.  ``pipeline.render`` therefore accepts only the
built-in shader classes themselves; a user subclass (e.g. the custom shader of
the reference's ``tests/smoke_test.py:155-232``) is rejected with
``UnsupportedShaderError`` -- there is no Python/CPU fallback.
"""
from __future__ import annotations

from typing import Any, NamedTuple, Tuple

import torch

ID = torch.Tensor


class UnsupportedShaderError(NotImplementedError):
    """Raised when ``render`` is given anything but a built-in shader class."""


class PerVertex(NamedTuple):
    """``shader.py:42-51``: clip-space position."""

    gl_Position: Any


class PerFragment(NamedTuple):
    """``shader.py:54-64``."""

    gl_FragDepth: Any = float("inf")
    keeps: Any = True
    use_default_depth: Any = False


class MixerOutput(NamedTuple):
    """``shader.py:80-88``."""

    keep: Any
    zbuffer: Any


_FUSED = (
    "stage `{}` of `{}` is fused into the CUDA kernels of jaxrenderer_b200 and is "
    "not callable from Python; call `jaxrenderer_b200.pipeline.render` with the "
    "built-in shader class instead."
)


class Shader:
    """Base class (``shader.py:91-396``).  Subclass-and-override is the
    reference's extension mechanism; here only the built-ins run (see module
    docstring).  ``_jr_shader`` is the C-ABI shader id of a built-in."""

    _jr_shader: int = -1

    @classmethod
    def _fused(cls, stage: str) -> "UnsupportedShaderError":
        return UnsupportedShaderError(_FUSED.format(stage, cls.__name__))

    @classmethod
    def vertex(cls, gl_VertexID: ID, gl_InstanceID: ID, camera: Any, extra: Any) -> Tuple[PerVertex, Any]:
        raise cls._fused("vertex")

    @classmethod
    def primitive_chooser(cls, gl_FragCoord: Any, gl_FrontFacing: Any, gl_PointCoord: Any, keeps: Any,
                          values: Any, barycentric_screen: Any, barycentric_clip: Any) -> Tuple[Any, ...]:
        raise cls._fused("primitive_chooser")

    @classmethod
    def interpolate(cls, values: Any, barycentric_screen: Any, barycentric_clip: Any) -> Any:
        raise cls._fused("interpolate")

    @classmethod
    def fragment(cls, gl_FragCoord: Any, gl_FrontFacing: Any, gl_PointCoord: Any, varying: Any,
                 extra: Any) -> Tuple[PerFragment, Any]:
        raise cls._fused("fragment")

    @classmethod
    def mix(cls, gl_FragDepth: Any, keeps: Any, extra: Any) -> Tuple[MixerOutput, Any]:
        raise cls._fused("mix")
