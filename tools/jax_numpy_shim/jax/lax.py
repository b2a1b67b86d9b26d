"""`jax.lax` subset on NumPy (eager)."""
from __future__ import annotations

import builtins as _bi

import numpy as _np

from .numpy import _wrap, _canon, _FLOAT
from .tree_util import tree_map


def full(shape, fill_value, dtype=None):
    if dtype is None:
        dtype = _canon(_np.asarray(fill_value).dtype)
    return _wrap(_np.full(shape, fill_value, dtype=dtype))


def full_like(x, fill_value, dtype=None):
    return _wrap(_np.full_like(_np.asarray(x), fill_value, dtype=dtype))


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(_np.asarray(pred)) else false_fun(*operands)


def select(pred, a, b):
    return _wrap(_np.where(_np.asarray(pred), _np.asarray(a), _np.asarray(b)))


def concatenate(operands, dimension):
    return _wrap(_np.concatenate([_np.asarray(o) for o in operands], axis=dimension))


def dot(a, b, **kw):
    return _wrap(_np.dot(_np.asarray(a), _np.asarray(b)))


def dot_general(lhs, rhs, dimension_numbers, **kw):
    (lc, rc), (lb, rb) = dimension_numbers
    lhs, rhs = _np.asarray(lhs), _np.asarray(rhs)
    assert tuple(lb) == tuple(rb) == (), "shim dot_general: no batch dimensions"
    return _wrap(_np.tensordot(lhs, rhs, axes=(list(lc), list(rc))))


def floor(x):
    return _wrap(_np.floor(_np.asarray(x)))


def tan(x):
    return _wrap(_np.tan(_np.asarray(x)))


def abs(x):  # noqa: A001
    return _wrap(_np.abs(_np.asarray(x)))


def pow(x, y):  # noqa: A001
    return _wrap(_np.power(_np.asarray(x), _np.asarray(y)))


def max(x, y):  # noqa: A001
    return _wrap(_np.maximum(_np.asarray(x), _np.asarray(y)))


def min(x, y):  # noqa: A001
    return _wrap(_np.minimum(_np.asarray(x), _np.asarray(y)))


def round(x, rounding_method=None):  # noqa: A001
    """lax.round default = AWAY_FROM_ZERO."""
    x = _np.asarray(x)
    return _wrap(_np.sign(x) * _np.floor(_np.abs(x) + _np.asarray(0.5, dtype=x.dtype)))


def transpose(x, permutation):
    return _wrap(_np.transpose(_np.asarray(x), permutation))


def iota(dtype, size):
    return _wrap(_np.arange(size, dtype=dtype))


def broadcasted_iota(dtype, shape, dimension):
    idx = _np.arange(shape[dimension], dtype=dtype)
    view = [1] * len(shape)
    view[dimension] = shape[dimension]
    return _wrap(_np.broadcast_to(idx.reshape(view), shape).copy())


def pad(operand, padding_value, padding_config):
    operand = _np.asarray(operand)
    assert _bi.all(interior == 0 for _, _, interior in padding_config)
    widths = [(lo, hi) for lo, hi, _ in padding_config]
    assert _bi.all(lo >= 0 and hi >= 0 for lo, hi in widths)
    return _wrap(_np.pad(operand, widths, mode="constant", constant_values=_np.asarray(padding_value).item()))


def dynamic_slice_in_dim(operand, start_index, slice_size, axis=0):
    """Start index is clamped so that the slice fits (lax semantics)."""
    operand = _np.asarray(operand)
    n = operand.shape[axis]
    start = int(_np.asarray(start_index))
    if start < 0:
        start += n
    start = int(_np.clip(start, 0, n - slice_size))
    sl = [slice(None)] * operand.ndim
    sl[axis] = slice(start, start + slice_size)
    return _wrap(operand[tuple(sl)])


def scan(f, init, xs, length=None, unroll=1):
    from .tree_util import tree_flatten

    leaves = tree_flatten(xs)[0]
    n = length if length is not None else _np.shape(leaves[0])[0]
    carry, ys = init, []
    for i in range(n):
        x = tree_map(lambda a: _wrap(_np.asarray(a)[i]), xs) if xs is not None else None
        carry, y = f(carry, x)
        ys.append(y)
    if ys and ys[0] is not None:
        ys = tree_map(lambda *v: _wrap(_np.stack([_np.asarray(t) for t in v])), *ys)
    else:
        ys = None
    return carry, ys


def stop_gradient(x):
    return x


class _Linalg:
    @staticmethod
    def triangular_solve(a, b, left_side=False, lower=False, transpose_a=False, conjugate_a=False,
                         unit_diagonal=False):
        a, b = _np.asarray(a, dtype=_FLOAT), _np.asarray(b, dtype=_FLOAT)
        tri = _np.tril(a) if lower else _np.triu(a)
        if unit_diagonal:
            tri = tri - _np.diag(_np.diag(tri)) + _np.eye(a.shape[-1], dtype=a.dtype)
        if transpose_a:
            tri = tri.T
        if left_side:      # solve tri @ x = b
            return _wrap(_np.linalg.solve(tri, b).astype(_FLOAT))
        # solve x @ tri = b
        return _wrap(_np.linalg.solve(tri.T, b.T).T.astype(_FLOAT))


linalg = _Linalg()
