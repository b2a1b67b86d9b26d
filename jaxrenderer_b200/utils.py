"""``renderer/utils.py`` shims that touch the path's outputs."""
from __future__ import annotations

from typing import Any, Sequence

import torch

from . import _native
from .types import Tensor


def transpose_for_display(matrix: Tensor, flip_vertical: bool = True) -> Tensor:
    """``(fst, snd, *c) -> (snd, fst, *c)``, optionally flipped so the origin
    is top-left (``utils.py:79-98``)."""
    mat = matrix.transpose(0, 1)
    if flip_vertical:
        mat = mat.flip(0)
    return mat


def canvas_to_uint8_display(canvas: Tensor) -> Tensor:
    """Fused output epilogue (the ``clip -> *255 -> uint8 ->
    transpose_for_display`` tail of every published benchmark,
    ``notebooks/32x32/A100.ipynb:326-337``): ``(B, W, H, 3) fp32 -> (B, H, W, 3) uint8``
    in one CUDA kernel."""
    import ctypes as C

    squeeze = canvas.ndim == 3
    c = canvas.unsqueeze(0) if squeeze else canvas
    if not c.is_cuda:
        raise RuntimeError("canvas_to_uint8_display needs a CUDA tensor (no CPU fallback)")
    c = c.contiguous().to(torch.float32)
    B, W, H, _ = c.shape
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=c.device)
    lib = _native.load()
    with torch.cuda.device(c.device):
        _native.check(lib.jr_canvas_to_uint8_display(c.data_ptr(), out.data_ptr(), B, W, H,
                                                     _native.stream_ptr(c.device)))
    return out[0] if squeeze else out


def build_texture_from_PyTinyrenderer(texture: Any, width: int, height: int) -> Tensor:
    """``utils.py:101-125``."""
    t = torch.as_tensor(texture, dtype=torch.float32).reshape(width, height, -1)
    return t.transpose(0, 1).flip(1)
