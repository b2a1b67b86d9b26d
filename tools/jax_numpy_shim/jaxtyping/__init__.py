"""`jaxtyping` stand-in: `isinstance(x, Float[Array, "..."])` checks "is an array of that dtype family"
(shapes are not checked); Python scalars are not arrays, as with the real package."""
from __future__ import annotations

import numpy as _np


class _Meta(type):
    kinds = None  # dtype kinds accepted, None = any

    def __instancecheck__(cls, obj):
        if not isinstance(obj, _np.ndarray):
            return False
        return cls.kinds is None or obj.dtype.kind in cls.kinds

    def __getitem__(cls, item):
        return cls


def _family(name, kinds):
    return _Meta(name, (), {"kinds": kinds})


class Array(metaclass=_Meta):
    pass


Float = _family("Float", "f")
Integer = Int = _family("Integer", "iu")
UInt8 = _family("UInt8", "u")
Num = Inexact = _family("Num", "fiu")
Bool = _family("Bool", "b")
Shaped = PyTree = _family("Shaped", None)
_Meta.kinds = None


def jaxtyped(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn
