"""Shared test helpers: scene builders and GPU-vs-oracle comparison."""
from __future__ import annotations

import os
from types import SimpleNamespace as NS
from typing import Any, Dict

import numpy as np
import torch

import jaxrenderer_b200 as jr
from jaxrenderer_b200 import synthetic
from jaxrenderer_b200.shaders import (
    DepthExtraInput, DepthShader, GouraudExtraInput, GouraudShader, GouraudTextureExtraInput,
    GouraudTextureShader, PhongReflectionShadowTextureExtraInput, PhongReflectionShadowTextureShader,
    PhongReflectionTextureExtraInput, PhongReflectionTextureShader, PhongTextureDarbouxExtraInput,
    PhongTextureDarbouxShader, PhongTextureExtraInput, PhongTextureShader,
)
from oracle import jr_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def smoke_scene(width: int, height: int, depth: float = 255.0):
    """Scene of the reference's tests/smoke_test.py:28-100 (5 triangles, Gouraud)."""
    eye = torch.tensor((0.0, 0, 2)); centre = torch.tensor((0.0, 0, 0)); up = torch.tensor((0.0, 1, 0))
    cam = jr.Camera.create(
        view=jr.Camera.view_matrix(eye, centre, up),
        projection=jr.Camera.perspective_projection_matrix(90.0, 1.0, -1.0, 1.0),
        viewport=jr.Camera.viewport_matrix(torch.zeros(2), torch.tensor((width, height)), depth))
    faces = torch.tensor(((0, 1, 2), (1, 3, 2), (0, 2, 4), (0, 4, 3), (2, 5, 1)), dtype=torch.int32)
    pos = torch.tensor(((0.0, 0, 0), (2, 0, 0), (0, 1, 0), (1, 1, 0), (-1, -1, 1), (-2, 0, 0)))
    col = torch.tensor(((1.0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0), (1, 1, 1), (1, 1, 0)))
    light = jr.LightSource(direction=torch.tensor((0.0, 0.0, -1.0)), colour=torch.ones(3))
    extra = GouraudExtraInput(position=pos, colour=col, normal=light.direction.expand(6, 3).contiguous(),
                              light=light)
    return cam, faces, extra


def compare(name: str, got_z, got_c, got_tri, ref: "O.RenderOut", rtol: float = 1e-5) -> Dict[str, Any]:
    """Parity report (BASELINE.json tolerances): triangle ids bit-exact except
    pixels whose competing depths differ by < 1e-6 (counted), colours within
    1e-5 relative."""
    got_tri = got_tri.cpu().to(torch.int64)
    mism = got_tri != ref.tri_id
    excused = mism & (ref.gap < 1e-6)
    hard = mism & ~excused
    ok = ~mism
    z_bad = int((got_z.cpu()[ok] != ref.zbuffer[ok]).sum())
    rep = {"name": name, "pixels": int(mism.numel()), "tri_mismatch": int(mism.sum()),
           "excused_depth_ties": int(excused.sum()), "hard_mismatch": int(hard.sum()),
           "z_not_bit_equal": z_bad}
    if got_c is not None:
        a, b = got_c.cpu()[ok], ref.targets[0][ok]
        err = (a - b).abs() / b.abs().clamp_min(1e-3)
        rep["colour_max_rel_err"] = float(err.max()) if err.numel() else 0.0
    return rep


def assert_parity(rep: Dict[str, Any], rtol: float = 1e-5) -> None:
    assert rep["hard_mismatch"] == 0, rep
    assert rep["z_not_bit_equal"] == 0, rep
    if "colour_max_rel_err" in rep:
        assert rep["colour_max_rel_err"] <= rtol, rep


def load_brax_fixture():
    """tests/golden/brax_ant_frames.npz -> (list[ModelObject] batched over frames, CameraParameters)."""
    d = np.load(os.path.join(GOLDEN, "brax_ant_frames.npz"))
    n = int(d["n_objects"])
    objs = []
    for i in range(n):
        g = lambda k: torch.from_numpy(d[f"o{i}_{k}"].copy())
        m = jr.Model(verts=g("verts"), norms=g("norms"), uvs=g("uvs"), faces=g("faces"),
                     faces_norm=g("faces_norm"), faces_uv=g("faces_uv"),
                     diffuse_map=g("diffuse_map"), specular_map=g("specular_map"))
        objs.append(jr.ModelObject(model=m, local_scaling=g("local_scaling"), transform=g("transform"),
                                   double_sided=g("double_sided")))
    cam = {k: torch.from_numpy(np.asarray(d[f"cam_{k}"]).copy()) for k in jr.CameraParameters._fields}
    return objs, jr.CameraParameters(**cam)


def random_mesh_scene(seed: int, n_tri: int = 60, W: int = 48, H: int = 40, tex: int = 16):
    """Random soup of triangles around the origin with all attributes, a
    perspective camera, light, texture, normal map, specular map."""
    g = torch.Generator().manual_seed(seed)
    V = 3 * n_tri
    centres = (torch.rand(n_tri, 1, 3, generator=g) - 0.5) * 2.0
    pos = (centres + (torch.rand(n_tri, 3, 3, generator=g) - 0.5) * 0.9).reshape(V, 3)
    nrm = torch.randn(V, 3, generator=g)
    uv_texel = torch.rand(V, 2, generator=g) * tex * 1.5 - 2.0
    uv01 = torch.rand(V, 2, generator=g) * 3.0 - 1.0
    col = torch.rand(V, 3, generator=g)
    faces = torch.arange(V, dtype=torch.int32).reshape(n_tri, 3)
    cam = jr.Renderer.create_camera_from_parameters(jr.CameraParameters(
        viewWidth=W, viewHeight=H, position=torch.tensor((2.0, 2.5, 1.5)), target=torch.zeros(3)))
    light = jr.LightSource(direction=torch.tensor((0.3, 0.5, 0.8)), colour=torch.tensor((1.0, 0.9, 0.8)))
    texture = torch.rand(tex, tex + 3, 3, generator=g)
    return NS(W=W, H=H, pos=pos, nrm=nrm, uv_texel=uv_texel, uv01=uv01, col=col, faces=faces, cam=cam,
              light=light, texture=texture, normal_map=torch.randn(tex, tex + 3, 3, generator=g),
              gen=g)


def cam_at(cam, b: int):
    """Un-batch element ``b`` of a (partly) batched Camera."""
    return NS(**{k: (v[b] if v.ndim == 3 else v) for k, v in cam._asdict().items()})
