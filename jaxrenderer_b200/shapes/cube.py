"""Unit cube, 24 vertices / 12 faces (``renderer/shapes/cube.py:16-159``).

The numeric tables are data extracted by ``tools/extract_reference_data.py``.
"""
from __future__ import annotations

import os
from functools import lru_cache
from typing import Any

import numpy as np
import torch

from ..model import Model
from ..types import _f32

_DATA = os.path.join(os.path.dirname(__file__), "_data", "cube.npz")


@lru_cache(maxsize=None)
def _tables():
    d = np.load(_DATA)
    return {k: torch.from_numpy(d[k].copy()) for k in ("verts", "normals", "uvs", "faces")}


def create_cube(half_extents: Any, texture_scaling: Any, diffuse_map: Any, specular_map: Any) -> Model:
    """``cube.py:143-159``."""
    t = _tables()
    half_extents = _f32(half_extents)
    dev = half_extents.device
    return Model(
        verts=t["verts"].to(dev) * half_extents,
        norms=t["normals"].to(dev),
        uvs=t["uvs"].to(dev) * _f32(texture_scaling, dev),
        faces=t["faces"].to(dev),
        faces_norm=t["faces"].to(dev),
        faces_uv=t["faces"].to(dev),
        diffuse_map=_f32(diffuse_map, dev),
        specular_map=_f32(specular_map, dev),
    )
