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
        lib.jr_oracle_visibility.restype = C.c_int
        lib.jr_oracle_visibility.argtypes = [
            C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong,
            C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
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


def visibility(world_to_clip, viewport, position, faces, W: int, H: int, num_threads: int = 0):
    """The visibility stage for ONE image, any shader: ``(idx int64, has bool, keeps_chosen bool, gap f32)``
    as torch tensors of shape (W, H) -- the tuple ``jr_oracle.visibility`` returns (bit-equal,
    ``tests/test_oracle_c.py``), computed by the C brute force in seconds at 960x540 x 19 980 triangles."""
    import torch

    lib = load()
    w2c, _ = _arr(world_to_clip, np.float32, 2)
    vp, _ = _arr(viewport, np.float32, 2)
    pos, _ = _arr(position, np.float32, 2)
    f, _ = _arr(faces, np.int32, 2)
    assert w2c.ndim == 2 and vp.ndim == 2 and pos.ndim == 2 and f.ndim == 2, "one image at a time"
    f = np.clip(f, 0, max(pos.shape[0] - 1, 0)).astype(np.int32)   # out-of-range ids clamp (jnp gather)
    idx = np.empty((W, H), np.int32); has = np.empty((W, H), np.uint8)
    kc = np.empty((W, H), np.uint8); gap = np.empty((W, H), np.float32)
    rc = lib.jr_oracle_visibility(1, W, H, f.shape[0], w2c.ctypes.data, 0, vp.ctypes.data, 0, pos.ctypes.data, 0,
                                  f.ctypes.data, 0, idx.ctypes.data, has.ctypes.data, kc.ctypes.data,
                                  gap.ctypes.data, int(num_threads))
    if rc != 0:
        raise MemoryError("jr_oracle_visibility failed")
    return (torch.from_numpy(idx.astype(np.int64)), torch.from_numpy(has.astype(bool)),
            torch.from_numpy(kc.astype(bool)), torch.from_numpy(gap))
