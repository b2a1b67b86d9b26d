"""The ``Shader`` plugin protocol of the reference (``renderer/shader.py:42-396``).

The five stage names (``vertex``, ``primitive_chooser``, ``interpolate``,
``fragment``, ``mix``) and the carrier tuples ``PerVertex`` / ``PerFragment`` /
``MixerOutput`` are kept so code written against the reference imports
unchanged.  In this implementation the stages of the seven built-in shaders
are FUSED into hand-written CUDA kernels (``csrc/jr_forward.cu``); rendering
never executes them as Python.  ``pipeline.render`` therefore accepts only the
built-in shader classes themselves; a user subclass (e.g. the custom shader of
the reference's ``tests/smoke_test.py:155-232``) is rejected with
``UnsupportedShaderError`` -- there is no Python/CPU fallback.  The stage
methods themselves stay callable with the reference's signatures (host-side
tensor code in ``stages.py``, checked against the oracle stage by stage).
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


class Shader:
    """Base class (``shader.py:91-396``).  Subclass-and-override is the reference's extension mechanism; here only the
    built-ins RENDER (see module docstring): ``_jr_shader`` is the C-ABI shader id of a built-in.  The five stage methods
    keep the reference's signatures and semantics as host-side tensor code (``stages.py``) so that code which calls,
    composes or introspects them keeps working; ``pipeline.render`` never calls them."""

    _jr_shader: int = -1

    @staticmethod
    def vertex(gl_VertexID: ID, gl_InstanceID: ID, camera: Any, extra: Any) -> Tuple[PerVertex, Any]:
        """Abstract in the reference as well (``shader.py:103-157``)."""
        raise NotImplementedError("vertex shader not implemented")

    @staticmethod
    def primitive_chooser(gl_FragCoord: Any, gl_FrontFacing: Any, gl_PointCoord: Any, keeps: Any,
                          values: Any, barycentric_screen: Any, barycentric_clip: Any) -> Tuple[Any, ...]:
        from . import stages

        return stages.base_primitive_chooser(gl_FragCoord, gl_FrontFacing, gl_PointCoord, keeps, values,
                                             barycentric_screen, barycentric_clip)

    @staticmethod
    def interpolate(values: Any, barycentric_screen: Any, barycentric_clip: Any) -> Any:
        from . import stages

        return stages.base_interpolate(values, barycentric_screen, barycentric_clip)

    @staticmethod
    def fragment(gl_FragCoord: Any, gl_FrontFacing: Any, gl_PointCoord: Any, varying: Any,
                 extra: Any) -> Tuple[PerFragment, Any]:
        from . import stages

        return stages.base_fragment(gl_FragCoord, gl_FrontFacing, gl_PointCoord, varying, extra)

    @staticmethod
    def mix(gl_FragDepth: Any, keeps: Any, extra: Any) -> Tuple[MixerOutput, Any]:
        from . import stages

        return stages.base_mix(gl_FragDepth, keeps, extra)


def _stage(name: str) -> staticmethod:
    """A static method that forwards to ``stages.<name>`` (imported lazily: ``stages`` imports this module)."""

    def call(*args: Any, **kwargs: Any) -> Any:
        from . import stages

        return getattr(stages, name)(*args, **kwargs)

    call.__name__ = name.split("_")[-1]
    call.__doc__ = f"Host-side evaluator of this stage: ``jaxrenderer_b200.stages.{name}`` (never called by ``render``)."
    return staticmethod(call)
