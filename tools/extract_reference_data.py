#!/usr/bin/env python
"""Extract the *numbers* the build needs from the read-only reference checkout.

Run in the build container only (``/root/reference`` does not exist on the GPU
box); its outputs are committed:

* ``jaxrenderer_b200/shapes/_data/{cube,capsule}.npz`` -- the constant mesh
  tables of ``renderer/shapes/cube.py:16-140`` and
  ``renderer/shapes/capsule.py:19-1956`` (vertices, normals, uvs, faces).
  Only the literal tuples are read (with ``ast``); no reference code runs.
* ``tests/golden/brax_ant_frames.npz`` -- a few frames of the pickled Brax
  "ant" scene ``test_resources/pre-gen-brax/inputs-30.zip`` (18 objects,
  3276 triangles), un-pickled without JAX.
"""
from __future__ import annotations

import ast
import io
import os
import pickle
import sys
import zipfile
from collections import namedtuple

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tables(path: str) -> dict[str, np.ndarray]:
    """Find ``_name = jnp.array(<tuple literal>)`` assignments, return them."""
    tree = ast.parse(open(path).read())
    out: dict[str, np.ndarray] = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.AnnAssign) and isinstance(node.target, ast.Name):
            name, value = node.target.id, node.value
        elif isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
            name, value = node.targets[0].id, node.value
        else:
            continue
        if name not in ("_verts", "_normals", "_uvs", "_faces"):
            continue
        if not (isinstance(value, ast.Call) and value.args):
            continue
        try:
            lit = ast.literal_eval(value.args[0])
        except ValueError:
            continue
        out[name] = np.asarray(lit)
    return out


def extract_shapes() -> None:
    dst = os.path.join(ROOT, "jaxrenderer_b200", "shapes", "_data")
    os.makedirs(dst, exist_ok=True)
    for shape in ("cube", "capsule"):
        t = _tables(os.path.join(REF, "renderer", "shapes", f"{shape}.py"))
        assert set(t) == {"_verts", "_normals", "_uvs", "_faces"}, t.keys()
        np.savez_compressed(
            os.path.join(dst, f"{shape}.npz"),
            verts=t["_verts"].astype(np.float32),
            normals=t["_normals"].astype(np.float32),
            uvs=t["_uvs"].astype(np.float32),
            faces=t["_faces"].astype(np.int32),
        )
        print(shape, {k: v.shape for k, v in t.items()})


Model = namedtuple(
    "Model", "verts norms uvs faces faces_norm faces_uv diffuse_map specular_map"
)
ModelObject = namedtuple("ModelObject", "model local_scaling transform double_sided")
CameraParameters = namedtuple(
    "CameraParameters",
    "viewWidth viewHeight viewDepth near far hfov vfov position target up",
)


def _reconstruct_array(fun, args, arr_state, aval_state):
    a = fun(*args)
    a.__setstate__(arr_state)
    return a


class _Unpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        if module == "renderer.model":
            return {"Model": Model, "ModelObject": ModelObject}[name]
        if module == "renderer.renderer":
            return {"CameraParameters": CameraParameters}[name]
        if module == "jax._src.array" and name == "_reconstruct_array":
            return _reconstruct_array
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core", 1)
        return super().find_class(module, name)


def extract_brax(frames=(0, 7, 15, 29)) -> None:
    zpath = os.path.join(REF, "test_resources", "pre-gen-brax", "inputs-30.zip")
    with zipfile.ZipFile(zpath) as z:
        names = z.namelist()
        payload = z.read(names[0])
    instances, camera, targets = _Unpickler(io.BytesIO(payload)).load()
    fr = list(frames)
    out: dict[str, np.ndarray] = {"n_objects": np.int32(len(instances))}
    for i, obj in enumerate(instances):
        m = obj.model
        # geometry / maps are constant across frames: keep frame 0 only.
        out[f"o{i}_verts"] = np.asarray(m.verts)[0].astype(np.float32)
        out[f"o{i}_norms"] = np.asarray(m.norms)[0].astype(np.float32)
        out[f"o{i}_uvs"] = np.asarray(m.uvs)[0].astype(np.float32)
        out[f"o{i}_faces"] = np.asarray(m.faces)[0].astype(np.int32)
        out[f"o{i}_faces_norm"] = np.asarray(m.faces_norm)[0].astype(np.int32)
        out[f"o{i}_faces_uv"] = np.asarray(m.faces_uv)[0].astype(np.int32)
        for k in ("verts", "norms", "uvs", "faces", "faces_norm", "faces_uv",
                  "diffuse_map", "specular_map"):
            a = np.asarray(getattr(m, k))
            assert (a == a[:1]).all(), (i, k)
        out[f"o{i}_diffuse_map"] = np.asarray(m.diffuse_map)[0].astype(np.float32)
        out[f"o{i}_specular_map"] = np.asarray(m.specular_map)[0].astype(np.float32)
        out[f"o{i}_local_scaling"] = np.asarray(obj.local_scaling)[fr].astype(np.float32)
        out[f"o{i}_transform"] = np.asarray(obj.transform)[fr].astype(np.float32)
        out[f"o{i}_double_sided"] = np.asarray(obj.double_sided)[fr]
    for k in CameraParameters._fields:
        out[f"cam_{k}"] = np.asarray(getattr(camera, k))[fr]
    out["targets"] = np.asarray(targets)[fr].astype(np.float32)
    out["frames"] = np.asarray(fr, dtype=np.int32)
    dst = os.path.join(ROOT, "tests", "golden")
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "brax_ant_frames.npz"), **out)
    nf = sum(out[f"o{i}_faces"].shape[0] for i in range(len(instances)))
    nv = sum(out[f"o{i}_verts"].shape[0] for i in range(len(instances)))
    print("brax ant:", len(instances), "objects", nf, "faces", nv, "verts")
    print({k: (out[f"cam_{k}"][0]) for k in CameraParameters._fields})


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference checkout not present; nothing to do")
    extract_shapes()
    extract_brax()
