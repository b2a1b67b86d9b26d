"""Gradients against the reference's own forward code.

`tests/golden/reference_grad.npz` (tools/gen_reference_grad_fixtures.py): central finite differences, in FLOAT64, of
`loss = sum(wz * zbuffer) + sum(wc * canvas)` computed by the UNMODIFIED reference (run on the NumPy stand-in for jax in
double precision) on scene `soup0` of `reference_run.npz`, for single entries of the differentiable inputs of the built-in
shaders (two fixture files: depth / gouraud / phong_reflection_shadow, and the four others).  The CPU test checks the oracle's autograd gradients against them, the GPU test the CUDA backward kernels.

Tolerance: BASELINE.json's 1e-4 for EVERY input, relative to the largest reference entry of the same input array.
Measured: every input of every shader within 4e-6; d z / d position (a cancellation of O(100)-sized terms in fp32) within
7e-5 (oracle) / 8e-5 (CUDA)."""
import os

import numpy as np
import pytest
import torch

import jaxrenderer_b200 as jr
from oracle import jr_oracle as O
from tests import test_reference_run as RR

_DIR = os.path.join(os.path.dirname(__file__), "golden")
G = {}
for _f in ("reference_grad.npz", "reference_grad_more.npz"):   # tools/gen_reference_grad_fixtures.py [more]
    if os.path.exists(os.path.join(_DIR, _f)):
        _z = np.load(os.path.join(_DIR, _f))
        G.update({k: _z[k] for k in _z.files})
if not G:
    pytest.skip("tests/golden/reference_grad*.npz have not been generated", allow_module_level=True)
P = "soup0"
SHADERS = tuple(k.split("/")[0] for k in G if k.endswith("/names"))
RTOL = {}   # every input at BASELINE.json's 1e-4 (d z / d position measured 7e-5..8e-5)
DEFAULT_RTOL = 1e-4
# reference_grad name -> (attribute path used to fetch the gradient from the leaves dict)
LEAF_KEYS = ("position", "normal", "colour", "light_direction", "light_colour", "world_to_clip", "viewport",
             "world_to_eye_norm", "atlas", "specular_map", "light_dir_eye", "ambient", "diffuse", "specular",
             "shadow_strength", "uv_texel", "texture", "normal_map")


def _leaves(dev=None):
    out = {}
    for k in LEAF_KEYS:
        t = RR.T(f"{P}/{k}", dev).clone().requires_grad_(True)
        out[k] = t
    return out


def _extra(name, L, dev=None):
    from types import SimpleNamespace as NS

    light = jr.LightSource(direction=L["light_direction"], colour=L["light_colour"])
    cam = NS(world_to_clip=L["world_to_clip"], viewport=L["viewport"], world_to_eye_norm=L["world_to_eye_norm"])
    S = RR.S
    if name == "depth":
        return cam, S.DepthShader, S.DepthExtraInput(position=L["position"])
    if name == "gouraud":
        return cam, S.GouraudShader, S.GouraudExtraInput(L["position"], L["colour"], L["normal"], light)
    T = RR.T
    if name in ("gouraud_texture", "phong"):
        cls = (S.GouraudTextureShader, S.GouraudTextureExtraInput) if name == "gouraud_texture" else (
            S.PhongTextureShader, S.PhongTextureExtraInput)
        return cam, cls[0], cls[1](L["position"], L["normal"], L["uv_texel"], light, L["texture"])
    if name == "phong_darboux":
        faces = T(P + "/faces", dev)
        i2f = torch.arange(faces.shape[0], dtype=torch.int32, device=dev).repeat_interleave(3)
        return cam, S.PhongTextureDarbouxShader, S.PhongTextureDarbouxExtraInput(
            L["position"], L["normal"], L["uv_texel"], light, L["texture"], L["normal_map"], i2f, faces)
    if name == "phong_reflection":
        return cam, S.PhongReflectionTextureShader, S.PhongReflectionTextureExtraInput(
            position=L["position"], normal=L["normal"], uv=T(P + "/uv01", dev), light=light,
            light_dir_eye=L["light_dir_eye"], texture_shape=T(P + "/texture_shape", dev),
            texture_index=T(P + "/texture_index", dev), texture_offset=int(RR.D[P + "/texture_offset"]),
            texture=L["atlas"], specular_map=L["specular_map"], ambient=L["ambient"], diffuse=L["diffuse"],
            specular=L["specular"])
    shadow_cam = NS(world_to_clip=T(P + "/shadow_world_to_clip", dev), viewport=T(P + "/shadow_viewport", dev))
    shadow = jr.Shadow(shadow_map=T(P + "/shadow_map", dev), strength=L["shadow_strength"], camera=shadow_cam)
    extra = S.PhongReflectionShadowTextureExtraInput(
        position=L["position"], normal=L["normal"], uv=T(P + "/uv01", dev), light=light, light_dir_eye=L["light_dir_eye"],
        texture_shape=T(P + "/texture_shape", dev), texture_index=T(P + "/texture_index", dev),
        texture_offset=int(RR.D[P + "/texture_offset"]), texture=L["atlas"], specular_map=L["specular_map"],
        shadow=shadow, camera=cam, ambient=L["ambient"], diffuse=L["diffuse"], specular=L["specular"])
    return cam, S.PhongReflectionShadowTextureShader, extra


def _compare(tag, name, L):
    names, index, want = G[f"{name}/names"], G[f"{name}/index"], G[f"{name}/grad"]
    assert len(names) > 10
    worst = {}
    for arr in sorted(set(names.tolist())):
        sel = names == arr
        scale = float(np.abs(want[sel]).max())
        if scale == 0.0:
            continue
        g = L[arr].grad
        g = torch.zeros_like(L[arr]) if g is None else g      # an input the shader does not differentiate
        g = g.detach().cpu().double().numpy()
        errs = []
        for idx, w in zip(index[sel], want[sel]):
            got = g[tuple(int(i) for i in idx if i >= 0)]
            errs.append(abs(got - w))
        rel = max(errs) / scale
        worst[arr] = rel
        assert rel <= RTOL.get(arr, DEFAULT_RTOL), f"[{tag}] {name}: d loss / d {arr}: {rel:.3g} of max |ref| {scale:.3g}"
    print(f"[{tag}] {name}: {len(names)} entries; worst error / max|ref| per input: "
          + ", ".join(f"{k} {v:.1e}" for k, v in worst.items()))


@pytest.mark.parametrize("name", SHADERS)
def test_oracle_gradients_match_reference_finite_differences(name):
    L = _leaves()
    cam, _, extra = _extra(name, L)
    W, H = int(RR.D[P + "/W"]), int(RR.D[P + "/H"])
    z0, c0 = torch.ones(W, H), torch.full((W, H, 3), 0.25)
    ref = O.render(cam, name, z0, () if name == "depth" else (c0,), RR.T(P + "/faces"), extra)
    loss = (ref.zbuffer * torch.from_numpy(G["wz"])).sum()
    if name != "depth":
        loss = loss + (ref.targets[0] * torch.from_numpy(G["wc"])).sum()
    assert abs(float(loss.detach()) - float(G[f"{name}/loss"])) <= 1e-4 * abs(float(G[f"{name}/loss"]))
    loss.backward()
    _compare("oracle", name, L)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SHADERS)
def test_cuda_backward_matches_reference_finite_differences(name):
    dev = torch.device("cuda", 0)
    L = _leaves(dev)
    cam, shader, extra = _extra(name, L, dev)
    camera = jr.Camera(*[None] * 8)._replace(world_to_clip=cam.world_to_clip, viewport=cam.viewport,
                                            world_to_eye_norm=cam.world_to_eye_norm)
    W, H = int(RR.D[P + "/W"]), int(RR.D[P + "/H"])
    z0, c0 = torch.ones(W, H, device=dev), torch.full((W, H, 3), 0.25, device=dev)
    out = jr.render(camera, shader, jr.Buffers(z0, () if name == "depth" else (c0,)), RR.T(P + "/faces", dev), extra)
    loss = (out.zbuffer * torch.from_numpy(G["wz"]).to(dev)).sum()
    if name != "depth":
        loss = loss + (out.targets[0] * torch.from_numpy(G["wc"]).to(dev)).sum()
    loss.backward()
    _compare("cuda", name, L)
