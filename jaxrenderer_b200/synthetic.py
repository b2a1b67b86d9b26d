"""Synthetic Brax-like scenes for tests and ``bench.py``.

Scenes are built from the reference's own cube / capsule tables so triangle
shapes match Brax environments: a huge, thin ground cube plus ``n_capsules``
randomly posed capsules near the origin ("ant-like": 10 capsules -> 1932
triangles, 5784 vertices; 17 -> 3276 / 9816 like the real ant; 104 -> 19980).
Everything is generated on the host with a fixed seed per environment
(``numpy.random.default_rng(20230701 + env)``) and is deterministic.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .shapes.capsule import _tables as _capsule_tables
from .shapes.cube import _tables as _cube_tables

SEED0 = 20230701
GROUND_TEXTURE_SCALING = 8192.0


def _quat_to_mat(q: np.ndarray) -> np.ndarray:
    """(…,4) (w,x,y,z) unit quaternions -> (…,3,3)."""
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    m = np.empty(q.shape[:-1] + (3, 3), dtype=np.float64)
    m[..., 0, 0] = 1 - 2 * (y * y + z * z); m[..., 0, 1] = 2 * (x * y - w * z); m[..., 0, 2] = 2 * (x * z + w * y)
    m[..., 1, 0] = 2 * (x * y + w * z); m[..., 1, 1] = 1 - 2 * (x * x + z * z); m[..., 1, 2] = 2 * (y * z - w * x)
    m[..., 2, 0] = 2 * (x * z - w * y); m[..., 2, 1] = 2 * (y * z + w * x); m[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return m


def scene_sizes(n_capsules: int) -> Tuple[int, int]:
    """(vertices, triangles) of a ground cube + n capsules."""
    return 24 + 576 * n_capsules, 12 + 192 * n_capsules


def brax_like_batch(B: int, n_capsules: int = 10, env0: int = 0,
                    with_attributes: bool = False) -> Dict[str, torch.Tensor]:
    """World-space merged meshes of ``B`` environments (host tensors).

    Returns ``position (B,Nv,3)``, ``faces (B,T,3)`` and per-environment camera
    ``eye (B,3)`` / ``target (B,3)``; with ``with_attributes`` also world-space
    ``normal (B,Nv,3)``, ``uv (Nv,2)`` and ``texture_index (Nv,)`` (object id).
    """
    cap, cube = _capsule_tables(), _cube_tables()
    cv = cap["verts"].numpy().astype(np.float64)        # (576,3), up axis = Y in table order
    cn = cap["normals"].numpy().astype(np.float64)
    shuf = [2, 0, 1]                                    # UpAxis.Z (capsule.py:1985-1991)
    cv, cn = cv[:, shuf], cn[:, shuf]
    gv = cube["verts"].numpy().astype(np.float64) * np.array([1000.0, 1000.0, 1e-4])
    gn = cube["normals"].numpy().astype(np.float64)
    Nv, T = scene_sizes(n_capsules)
    pos = np.empty((B, Nv, 3), dtype=np.float32)
    nrm = np.empty((B, Nv, 3), dtype=np.float32) if with_attributes else None
    eye = np.empty((B, 3), dtype=np.float32)
    tgt = np.empty((B, 3), dtype=np.float32)
    for b in range(B):
        rng = np.random.default_rng(SEED0 + env0 + b)
        radius = rng.uniform(0.04, 0.1, size=n_capsules)
        hh = rng.uniform(0.05, 0.3, size=n_capsules)
        q = rng.normal(size=(n_capsules, 4)); q /= np.linalg.norm(q, axis=-1, keepdims=True)
        R = _quat_to_mat(q)                              # (N,3,3)
        d = rng.normal(size=(n_capsules, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True)
        centre = d * rng.uniform(0.0, 1.0, size=(n_capsules, 1)) ** (1 / 3) * 0.6
        centre[:, 2] = np.abs(centre[:, 2]) * 0.5 + 0.45
        v = cv[None] * radius[:, None, None]             # (N,576,3)
        v[..., 2] += np.where(v[..., 2] > 0, hh[:, None], -hh[:, None])
        world = np.einsum("nij,nvj->nvi", R, v) + centre[:, None, :]
        pos[b, :24] = gv
        pos[b, 24:] = world.reshape(-1, 3)
        if nrm is not None:
            nrm[b, :24] = gn
            nrm[b, 24:] = np.einsum("nij,vj->nvi", R, cn).reshape(-1, 3)
        root = np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), 0.0])
        dist = rng.uniform(1.15, 1.45)
        eye[b] = root + np.array([2.0 * dist, -2.0 * dist, 1.5 * dist])
        tgt[b] = root
    gf = cube["faces"].numpy()
    cf = cap["faces"].numpy()
    faces = np.concatenate([gf] + [cf + 24 + 576 * i for i in range(n_capsules)], axis=0).astype(np.int32)
    out = {
        "position": torch.from_numpy(pos),
        "faces": torch.from_numpy(np.broadcast_to(faces, (B, T, 3)).copy()),
        "eye": torch.from_numpy(eye), "target": torch.from_numpy(tgt),
    }
    if with_attributes:
        # ground uvs carry the Brax glue's texture scaling (8192 repeats of the 100x100 grid over the plane,
        # notebooks/Generate Data.ipynb), as in the real fixture; shaders taking texel-space uvs rescale them
        uv = np.concatenate([cube["uvs"].numpy() * GROUND_TEXTURE_SCALING] + [cap["uvs"].numpy()] * n_capsules, axis=0)
        tix = np.concatenate([np.zeros(24, np.int32)] + [np.full(576, i + 1, np.int32) for i in range(n_capsules)])
        out.update({"normal": torch.from_numpy(nrm), "uv": torch.from_numpy(uv.astype(np.float32)),
                    "texture_index": torch.from_numpy(tix)})
    return out


def brax_like_objects(B: int, n_capsules: int = 10, env0: int = 0, body_seed: int = 0, device=None):
    """The same kind of scene in FACTORED form, the way Brax hands it to the renderer (``notebooks/Generate
    Data.ipynb``: one ``ModelObject`` per body, the mesh fixed, the transform per environment): a ground cube
    (``local_scaling`` = half extents) + ``n_capsules`` capsules whose sizes (the robot's BODY, seeded by
    ``body_seed``) are shared by all environments and whose POSES (rotation + position, seeded per environment like
    ``brax_like_batch``) carry the batch axis.  Returns ``(objects, eye (B,3), target (B,3))``; host tensors unless
    ``device`` is given.  ``merge_objects(objects)`` has 24 + 576 n vertices and 12 + 192 n triangles."""
    from .model import ModelObject
    from .shapes.capsule import UpAxis, create_capsule
    from .shapes.cube import create_cube

    body = np.random.default_rng(SEED0 - 1 - body_seed)
    radius = body.uniform(0.04, 0.1, size=n_capsules).astype(np.float32)
    hh = body.uniform(0.05, 0.3, size=n_capsules).astype(np.float32)
    T = np.zeros((n_capsules, B, 4, 4), dtype=np.float32)
    T[..., 3, 3] = 1.0
    eye = np.empty((B, 3), dtype=np.float32)
    tgt = np.empty((B, 3), dtype=np.float32)
    for b in range(B):
        rng = np.random.default_rng(SEED0 + env0 + b)
        q = rng.normal(size=(n_capsules, 4)); q /= np.linalg.norm(q, axis=-1, keepdims=True)
        d = rng.normal(size=(n_capsules, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True)
        centre = d * rng.uniform(0.0, 1.0, size=(n_capsules, 1)) ** (1 / 3) * 0.6
        centre[:, 2] = np.abs(centre[:, 2]) * 0.5 + 0.45
        T[:, b, :3, :3] = _quat_to_mat(q)
        T[:, b, :3, 3] = centre
        root = np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), 0.0])
        dist = rng.uniform(1.15, 1.45)
        eye[b] = root + np.array([2.0 * dist, -2.0 * dist, 1.5 * dist])
        tgt[b] = root
    dev = torch.device(device) if device is not None else torch.device("cpu")
    g = torch.Generator().manual_seed(1)
    ground = create_cube(torch.ones(3), torch.tensor((GROUND_TEXTURE_SCALING,) * 2), checker_texture(), torch.full((100, 100), 2.0))
    objects = [ModelObject(model=type(ground)(*[t.to(dev) for t in ground]),
                           local_scaling=torch.tensor((1000.0, 1000.0, 1e-4), device=dev))]
    for i in range(n_capsules):
        m = create_capsule(float(radius[i]), float(hh[i]), UpAxis.Z, torch.rand(1, 1, 3, generator=g), torch.full((1, 1), 2.0))
        objects.append(ModelObject(model=type(m)(*[t.to(dev) for t in m]), transform=torch.from_numpy(T[i]).to(dev)))
    return objects, torch.from_numpy(eye), torch.from_numpy(tgt)


def brax_cameras(eye: torch.Tensor, target: torch.Tensor, width: int, height: int,
                 hfov: float = 58.0, vfov: Optional[float] = None):
    """Full-view cameras (``viewWidth=width, viewHeight=height``, SURVEY 6 note 3)."""
    from .renderer import CameraParameters, Renderer

    vfov = hfov * height / width if vfov is None else vfov
    return Renderer.create_camera_from_parameters(CameraParameters(
        viewWidth=width, viewHeight=height, hfov=hfov, vfov=vfov, position=eye, target=target))


def checker_texture(w: int = 100, h: int = 100, cell: int = 10) -> torch.Tensor:
    ii, jj = np.meshgrid(np.arange(w), np.arange(h), indexing="ij")
    on = ((ii // cell + jj // cell) % 2).astype(np.float32)
    tex = np.stack([0.2 + 0.6 * on, 0.3 + 0.4 * (1 - on), 0.5 + 0.3 * on], axis=-1)
    return torch.from_numpy(tex.astype(np.float32))


def merged_model_from_batch(sc: Dict[str, torch.Tensor], n_capsules: int, device, atlas_tex: int = 100):
    """``MergedModel`` of a ``brax_like_batch(..., with_attributes=True)``: shared topology, batched
    world-space attributes, an atlas with a checker texture for the ground and 1x1 textures for the
    capsules (the layout of the real Brax ant fixture)."""
    from .model import MergedModel

    nv, _ = scene_sizes(n_capsules)
    n_obj = n_capsules + 1
    g = torch.Generator().manual_seed(1)
    atlas = torch.zeros(n_obj * atlas_tex, atlas_tex, 3)
    atlas[:atlas_tex] = checker_texture(atlas_tex, atlas_tex)
    shapes = [[atlas_tex, atlas_tex]] + [[1, 1]] * n_capsules
    for i in range(1, n_obj):
        atlas[i * atlas_tex, 0] = torch.rand(3, generator=g)
    faces = sc["faces"][0].to(device)
    return MergedModel(
        verts=sc["position"].to(device), norms=sc["normal"].to(device), uvs=sc["uv"].to(device),
        faces=faces, faces_norm=faces, faces_uv=faces,
        texture_index=sc["texture_index"].to(device),
        double_sided=torch.zeros(nv, dtype=torch.bool, device=device),
        texture_shape=torch.tensor(shapes, dtype=torch.int32, device=device), offset=atlas_tex,
        diffuse_map=atlas.to(device), specular_map=torch.full((n_obj, 1), 2.0, device=device))
