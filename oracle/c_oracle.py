"""ctypes wrapper of the C oracle (``oracle/jr_oracle_c.c``) -- test
infrastructure only (tests, smoke, bench CPU legs)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libjr_oracle.so")
_lib: Optional[C.CDLL] = None


def load(build_if_missing: bool = True) -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) and build_if_missing:
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        lib = C.CDLL(LIB)
        lib.jr_oracle_depth.restype = C.c_int
        lib.jr_oracle_depth.argtypes = [
            C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong,
            C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int]
        lib.jr_oracle_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(load().jr_oracle_max_threads())


def _arr(a, dtype, rank: int) -> Tuple[np.ndarray, int]:
    a = np.ascontiguousarray(np.asarray(a, dtype=dtype))
    if a.ndim == rank:
        return a, 0
    assert a.ndim == rank + 1, (a.shape, rank)
    return a, int(np.prod(a.shape[1:]))


def render_depth(world_to_clip, viewport, position, faces, zbuffer, num_threads: int = 0):
    """Brute-force depth render of a batch; returns (zbuffer (B,W,H), tri_id (B,W,H))."""
    lib = load()
    z = np.array(zbuffer, dtype=np.float32, copy=True, order="C")
    if z.ndim == 2:
        z = z[None]
    B, W, H = z.shape
    w2c, s0 = _arr(world_to_clip, np.float32, 2)
    vp, s1 = _arr(viewport, np.float32, 2)
    pos, s2 = _arr(position, np.float32, 2)
    f, s3 = _arr(faces, np.int32, 2)
    tri = np.empty((B, W, H), dtype=np.int32)
    rc = lib.jr_oracle_depth(B, W, H, f.shape[-2], w2c.ctypes.data, s0, vp.ctypes.data, s1,
                             pos.ctypes.data, s2, f.ctypes.data, s3, z.ctypes.data, tri.ctypes.data,
                             int(num_threads))
    if rc != 0:
        raise MemoryError("jr_oracle_depth failed")
    return z, tri
