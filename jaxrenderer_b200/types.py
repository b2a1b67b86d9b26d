"""Data carriers of the rasterisation path (host side, torch tensors).

Mirrors ``renderer/types.py:134-155`` of the reference (``LightSource``,
``Buffers``, ``DtypeInfo``).  Conventions kept from the reference: all floats
are fp32, all integers int32, buffers are **x-major** (``zbuffer[x, y]``,
``canvas[x, y, c]``) with the origin at the bottom-left.

Every leaf may carry ONE extra leading batch axis; this replaces the
reference's ``jax.vmap`` idiom (``examples/batch_rendering.py:87-95``).
"""
from __future__ import annotations

from typing import Any, Generic, NamedTuple, Tuple, TypeVar

import torch

Tensor = torch.Tensor

_TargetsT = TypeVar("_TargetsT", bound=Tuple[Any, ...])


def _f32(x: Any, device: Any = None) -> Tensor:
    """``jnp.asarray(x, dtype=float32)`` equivalent."""
    if isinstance(x, torch.Tensor):
        t = x if x.dtype == torch.float32 else x.to(torch.float32)
        return t if device is None else t.to(device)
    return torch.as_tensor(x, dtype=torch.float32, device=device)


def _i32(x: Any, device: Any = None) -> Tensor:
    if isinstance(x, torch.Tensor):
        t = x if x.dtype == torch.int32 else x.to(torch.int32)
        return t if device is None else t.to(device)
    return torch.as_tensor(x, dtype=torch.int32, device=device)


class LightSource(NamedTuple):
    """Parallel light (``renderer/types.py:134-141``)."""

    direction: Any = (0.0, 0.0, -1.0)
    colour: Any = (1.0, 1.0, 1.0)


class Buffers(NamedTuple, Generic[_TargetsT]):
    """``zbuffer (W, H)`` + tuple of targets ``(W, H, ...)``
    (``renderer/types.py:148-155``)."""

    zbuffer: Tensor
    targets: _TargetsT


class DtypeInfo(NamedTuple):
    """``renderer/types.py:95-131``."""

    min: float
    max: float
    bits: int
    dtype: torch.dtype

    @classmethod
    def create(cls, dtype: torch.dtype) -> "DtypeInfo":
        if dtype.is_floating_point:
            fi = torch.finfo(dtype)
            return cls(min=fi.min, max=fi.max, bits=fi.bits, dtype=dtype)
        ii = torch.iinfo(dtype)
        return cls(min=ii.min, max=ii.max, bits=ii.bits, dtype=dtype)
