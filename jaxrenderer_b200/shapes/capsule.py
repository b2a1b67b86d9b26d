"""Capsule, 576 vertices / 192 faces (``renderer/shapes/capsule.py:19-2015``).

The numeric tables are data extracted by ``tools/extract_reference_data.py``.
"""
from __future__ import annotations

import enum
import os
from functools import lru_cache
from typing import Any

import numpy as np
import torch

from ..model import Model
from ..types import _f32

_DATA = os.path.join(os.path.dirname(__file__), "_data", "capsule.npz")


class UpAxis(enum.IntEnum):
    X = 0
    Y = 1
    Z = 2


@lru_cache(maxsize=None)
def _tables():
    d = np.load(_DATA)
    return {k: torch.from_numpy(d[k].copy()) for k in ("verts", "normals", "uvs", "faces")}


_SHUFFLE = ((1, 2, 0), (0, 1, 2), (2, 0, 1))


def create_capsule(radius: Any, half_height: Any, up_axis: UpAxis, diffuse_map: Any,
                   specular_map: Any) -> Model:
    """``capsule.py:1964-2015``: scale the unit capsule by ``radius`` and push
    both hemispheres ``half_height`` apart along ``up_axis``."""
    t = _tables()
    radius = _f32(radius)
    dev = radius.device
    shuffled = list(_SHUFFLE[int(up_axis)])
    verts = t["verts"].to(dev)[:, shuffled] * radius
    up = verts[:, int(up_axis)]
    hh = _f32(half_height, dev)
    verts = verts.clone()
    verts[:, int(up_axis)] = up + torch.where(up > 0, hh, -hh)
    return Model(
        verts=verts,
        norms=t["normals"].to(dev)[:, shuffled],
        uvs=t["uvs"].to(dev),
        faces=t["faces"].to(dev),
        faces_norm=t["faces"].to(dev),
        faces_uv=t["faces"].to(dev),
        diffuse_map=_f32(diffuse_map, dev),
        specular_map=_f32(specular_map, dev),
    )
