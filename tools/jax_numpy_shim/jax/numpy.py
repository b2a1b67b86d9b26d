"""`jax.numpy` on NumPy: float32 / int32 canonical dtypes, arrays with the functional `.at[...]` API."""
from __future__ import annotations

import builtins as _bi
import os as _os

import numpy as _np

single = float32 = _np.float32
int32 = _np.int32
uint8 = _np.uint8
bool_ = _np.bool_
integer = _np.integer
floating = _np.floating
inf = float("inf")
pi = float(_np.pi)
newaxis = None
ndarray = _np.ndarray


# JAX_SHIM_X64=1: keep float64 (the analogue of jax_enable_x64) -- used for finite-difference gradient references
X64 = _os.environ.get("JAX_SHIM_X64", "0") == "1"
_FLOAT = _np.float64 if X64 else _np.float32


def _canon(dtype):
    dtype = _np.dtype(dtype)
    if dtype == _np.float64:
        return _np.dtype(_FLOAT)
    if X64 and dtype == _np.float32:
        return _np.dtype(_np.float64)
    if dtype == _np.int64:
        return _np.dtype(_np.int32)
    if dtype == _np.uint64:
        return _np.dtype(_np.uint32)
    return dtype


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _norm_idx(self):
        idx = self.idx
        if isinstance(idx, tuple):
            return tuple(_np.asarray(i) if isinstance(i, _np.ndarray) else i for i in idx)
        return _np.asarray(idx) if isinstance(idx, _np.ndarray) else idx

    def set(self, v):
        out = _np.array(self.arr, copy=True)
        out[self._norm_idx()] = _np.asarray(v)
        return _wrap(out)

    def add(self, v):
        out = _np.array(self.arr, copy=True)
        _np.add.at(out, self._norm_idx(), _np.asarray(v).astype(out.dtype))
        return _wrap(out)

    def get(self, mode=None, fill_value=None, **kw):
        """Gather with JAX out-of-bounds semantics: negative indices wrap ONCE (NumPy style), then
        mode="fill" returns fill_value when still out of range, default ("clip"/promise) clamps."""
        arr = _np.asarray(self.arr)
        idx = self.idx if isinstance(self.idx, tuple) else (self.idx,)
        assert _bi.all(_np.ndim(i) == 0 for i in idx), "shim .at[].get supports scalar indices only"
        fixed, oob = [], False
        for i, n in zip(idx, arr.shape):
            i = int(i)
            if i < 0:
                i += n
            if i < 0 or i >= n:
                oob = True
                i = min(max(i, 0), n - 1)
            fixed.append(i)
        if oob and mode == "fill":
            return _wrap(_np.asarray(fill_value, dtype=arr.dtype))
        return _wrap(arr[tuple(fixed)])


class Array(_np.ndarray):
    """ndarray with `.at`; arithmetic keeps the subclass, dtypes stay float32 / int32."""

    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self

    def __array_finalize__(self, obj):
        pass

    def __getitem__(self, idx):
        # JAX clamps out-of-range gather indices instead of raising (default mode for `x[idx]`)
        if isinstance(idx, _np.ndarray) and idx.dtype.kind in "iu" and self.ndim >= 1:
            n = self.shape[0]
            idx = _np.where(idx < 0, idx + n, idx)
            idx = _np.clip(idx, 0, n - 1)
            return _wrap(_np.asarray(self)[idx])
        if (isinstance(idx, tuple) and _bi.all(isinstance(i, (int, _np.integer, _np.ndarray)) for i in idx)
                and _bi.any(isinstance(i, _np.ndarray) for i in idx)):
            fixed = []
            for i, n in zip(idx, self.shape):
                i = _np.asarray(i)
                i = _np.where(i < 0, i + n, i)
                fixed.append(_np.clip(i, 0, n - 1))
            return _wrap(_np.asarray(self)[tuple(fixed)])
        r = super().__getitem__(idx)
        return _wrap(r) if isinstance(r, _np.ndarray) or _np.isscalar(r) else r


def _wrap(x):
    a = _np.asarray(x)
    c = _canon(a.dtype)
    if a.dtype != c:
        a = a.astype(c)
    return a.view(Array)


def array(x, dtype=None, copy=True):
    a = _np.array(x, dtype=dtype)
    return _wrap(a)


def asarray(x, dtype=None):
    return _wrap(_np.asarray(x, dtype=dtype))


def zeros(shape, dtype=None):
    return _wrap(_np.zeros(shape, dtype=dtype or _FLOAT))


def ones(shape, dtype=None):
    return _wrap(_np.ones(shape, dtype=dtype or _FLOAT))


def full(shape, fill_value, dtype=None):
    return _wrap(_np.full(shape, fill_value, dtype=dtype))


def ones_like(x, dtype=None):
    return _wrap(_np.ones_like(_np.asarray(x), dtype=dtype))


def zeros_like(x, dtype=None):
    return _wrap(_np.zeros_like(_np.asarray(x), dtype=dtype))


def identity(n, dtype=None):
    return _wrap(_np.identity(n, dtype=dtype or _FLOAT))


def eye(n, dtype=None):
    return _wrap(_np.eye(n, dtype=dtype or _FLOAT))


def arange(*a, dtype=None):
    return _wrap(_np.arange(*a, dtype=dtype))


def _lift(fn):
    def g(*a, **k):
        r = fn(*[_np.asarray(x) if isinstance(x, (_np.ndarray, list, tuple)) else x for x in a], **k)
        if isinstance(r, tuple):
            return tuple(_wrap(v) for v in r)
        return _wrap(r)
    g.__name__ = getattr(fn, "__name__", "lifted")
    return g


cross = _lift(_np.cross)
where = _lift(_np.where)
radians = _lift(_np.radians)
logical_and = _lift(_np.logical_and)
dot = _lift(_np.dot)
all = _lift(_np.all)  # noqa: A001
any = _lift(_np.any)  # noqa: A001
vstack = _lift(_np.vstack)
stack = _lift(_np.stack)
concatenate = _lift(_np.concatenate)
sin = _lift(_np.sin)
cos = _lift(_np.cos)
tan = _lift(_np.tan)
outer = _lift(_np.outer)
maximum = _lift(_np.maximum)
minimum = _lift(_np.minimum)
diag = _lift(_np.diag)
swapaxes = _lift(_np.swapaxes)
reshape = _lift(_np.reshape)
modf = _lift(_np.modf)
isclose = _lift(_np.isclose)
broadcast_to = _lift(_np.broadcast_to)
abs = _lift(_np.abs)  # noqa: A001
floor = _lift(_np.floor)
sqrt = _lift(_np.sqrt)
clip = _lift(_np.clip)
sum = _lift(_np.sum)  # noqa: A001
transpose = _lift(_np.transpose)
expand_dims = _lift(_np.expand_dims)
squeeze = _lift(_np.squeeze)


def argmin(x, axis=None):
    """NaN-propagating first-index argmin (numpy semantics, which jnp follows)."""
    return _wrap(_np.argmin(_np.asarray(x), axis=axis))


def argmax(x, axis=None):
    return _wrap(_np.argmax(_np.asarray(x), axis=axis))


def ndim(x):
    return _np.ndim(x)


def shape(x):
    return _np.shape(x)


def issubdtype(a, b):
    return _np.issubdtype(a, b)


def iinfo(d):
    return _np.iinfo(d)


def finfo(d):
    return _np.finfo(d)


class _Linalg:
    @staticmethod
    def inv(a):
        a = _np.asarray(a, dtype=_FLOAT)
        try:
            return _wrap(_np.linalg.inv(a))
        except _np.linalg.LinAlgError:
            # jax returns inf / nan for a singular input instead of raising (the reference masks such
            # triangles with `keep = |det| > 1e-6`)
            return _wrap(_np.full(a.shape, _np.nan, dtype=_FLOAT))

    @staticmethod
    def det(a):
        """3x3: the closed form jax lowers to (`_det_3x3`), same term order."""
        a = _np.asarray(a, dtype=_FLOAT)
        if a.shape[-2:] == (3, 3):
            return _wrap(a[..., 0, 0] * a[..., 1, 1] * a[..., 2, 2] + a[..., 0, 1] * a[..., 1, 2] * a[..., 2, 0]
                         + a[..., 0, 2] * a[..., 1, 0] * a[..., 2, 1] - a[..., 0, 2] * a[..., 1, 1] * a[..., 2, 0]
                         - a[..., 0, 0] * a[..., 1, 2] * a[..., 2, 1] - a[..., 0, 1] * a[..., 1, 0] * a[..., 2, 2])
        return _wrap(_np.linalg.det(a))

    @staticmethod
    def norm(a, *args, **kw):
        return _wrap(_np.linalg.norm(_np.asarray(a), *args, **kw))


linalg = _Linalg()
