"""Data carriers of the rasterisation path (host side, torch tensors).

Mirrors ``renderer/types.py:134-155`` of the reference (``LightSource``,
``Buffers``, ``DtypeInfo``).  Conventions kept from the reference: all floats
are fp32, all integers int32, buffers are **x-major** (``zbuffer[x, y]``,
``canvas[x, y, c]``) with the origin at the bottom-left.

Every leaf may carry ONE extra leading batch axis; this replaces the
reference's ``jax.vmap`` idiom (``examples/batch_rendering.py:87-95``).
"""
from __future__ import annotations

import collections
from typing import Any, Generic, NamedTuple, Tuple, TypeVar

import numpy as np
import torch

Tensor = torch.Tensor
# array aliases of the reference (``renderer/types.py``): all plain fp32 tensors here
Colour = Vec3f = Texture = SpecularMap = Tensor

_TargetsT = TypeVar("_TargetsT", bound=Tuple[Any, ...])


_CONST_CACHE: "collections.OrderedDict[tuple, Tensor]" = collections.OrderedDict()
_CONST_CACHE_SIZE = 512
_CONST_MAX_ELEMS = 64


def _device_constant(arr: np.ndarray, device: torch.device) -> Tensor:
    """Device copy of a small host constant (light parameters, camera intrinsics given as Python
    numbers...), memoised by VALUE.  A copy from pageable host memory makes the host wait for the
    stream, which would serialise the host with all queued kernels once per parameter and per call;
    this way the wait is paid once per distinct value.  The returned tensor is shared: read-only."""
    key = (arr.dtype.str, arr.shape, arr.tobytes(), device.type, device.index)
    t = _CONST_CACHE.get(key)
    if t is None:
        t = torch.from_numpy(arr.copy()).to(device)
        _CONST_CACHE[key] = t
        while len(_CONST_CACHE) > _CONST_CACHE_SIZE:
            _CONST_CACHE.popitem(last=False)
    else:
        _CONST_CACHE.move_to_end(key)
    return t


def _as_dtype(x: Any, dtype: torch.dtype, np_dtype: Any, device: Any) -> Tensor:
    dev = torch.device(device) if device is not None else None
    to_cuda = dev is not None and dev.type == "cuda"
    if isinstance(x, torch.Tensor):
        if to_cuda and not x.is_cuda and x.numel() <= _CONST_MAX_ELEMS and not x.requires_grad:
            if dev.index is None:
                dev = torch.device("cuda", torch.cuda.current_device())
            return _device_constant(x.detach().to(dtype).numpy(), dev)
        t = x if x.dtype == dtype else x.to(dtype)
        return t if dev is None else t.to(dev)
    if to_cuda:
        arr = np.asarray(x, dtype=np_dtype)
        if arr.size <= _CONST_MAX_ELEMS:
            if dev.index is None:
                dev = torch.device("cuda", torch.cuda.current_device())
            return _device_constant(arr, dev)
        return torch.from_numpy(arr.copy()).to(dev)
    return torch.as_tensor(x, dtype=dtype, device=dev)


def _is_vmapped(*values: Any) -> bool:
    """True when any value is a ``torch.func.vmap`` batched tensor (the native fast paths take raw device pointers:
    inside a vmap they step aside for the torch builders, and ``pipeline._render_arrays`` re-batches natively)."""
    f = torch._C._functorch
    return any(isinstance(v, torch.Tensor) and f.is_batchedtensor(v) for v in values)


def _f32(x: Any, device: Any = None) -> Tensor:
    """``jnp.asarray(x, dtype=float32)`` equivalent."""
    return _as_dtype(x, torch.float32, np.float32, device)


def _i32(x: Any, device: Any = None) -> Tensor:
    return _as_dtype(x, torch.int32, np.int32, device)


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
