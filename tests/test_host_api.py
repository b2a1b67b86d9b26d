"""CPU tests of the host-side mirror of the reference API, the C-ABI surface and the
no-fallback rules."""
import ctypes
import math
import os
import re

import pytest
import torch

import jaxrenderer_b200 as jr
from jaxrenderer_b200 import _native
from jaxrenderer_b200.shaders import BUILTIN_SHADERS, DepthExtraInput, DepthShader, GouraudShader
from oracle import jr_oracle as O
from tests.helpers import load_brax_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_library_exports_every_declared_symbol():
    lib = _native.load()
    header = open(os.path.join(ROOT, "include", "jr_b200.h")).read()
    declared = set(re.findall(r"\b(jr_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
        # every entry point has its ctypes signature declared (an undeclared one would truncate 64-bit pointers)
        assert getattr(lib, name).argtypes is not None or name in ("jr_abi_version", "jr_launch_count"), name
    assert lib.jr_abi_version() == int(re.search(r'#define\s+JR_ABI_VERSION\s+(\d+)', header).group(1))
    assert lib.jr_strerror(-4).decode() == "workspace too small"
    # the per-kernel timing switch needs no device: nothing was launched, nothing is reported
    assert lib.jr_debug_kernel_timing(1) == 0 and lib.jr_debug_kernel_times(None, 0) == 0
    assert lib.jr_debug_kernel_timing(0) == 0 and _native.kernel_times() == []
    # struct layout agreed between Python and C: a NULL args pointer is reported, not crashed on
    assert lib.jr_render_forward(None, None) == -1
    bad = _native.JrRenderArgs()
    bad.shader = 99
    assert lib.jr_render_forward(ctypes.byref(bad), None) == -3
    bad.shader, bad.B, bad.W, bad.H = 0, 1, 0, 4
    assert lib.jr_render_forward(ctypes.byref(bad), None) == -2


def test_custom_shader_is_rejected_without_fallback():
    class MyShader(GouraudShader):  # overriding a stage, like reference tests/smoke_test.py:155
        @staticmethod
        def fragment(*a, **k):
            return None

    bufs = jr.Buffers(torch.zeros(4, 4), (torch.zeros(4, 4, 3),))
    with pytest.raises(jr.UnsupportedShaderError, match="Custom `Shader` subclasses are not supported"):
        jr.render(None, MyShader, bufs, torch.zeros(1, 3, dtype=torch.int32), (torch.zeros(3, 3),))
    with pytest.raises(jr.UnsupportedShaderError):
        jr.render(None, jr.Shader, bufs, torch.zeros(1, 3, dtype=torch.int32), (torch.zeros(3, 3),))
    for s in BUILTIN_SHADERS:  # the stage methods stay callable (host-side tensor code, tests/test_stage_methods.py) ...
        for stage in ("vertex", "primitive_chooser", "interpolate", "fragment", "mix"):
            assert callable(getattr(s, stage))
    # ... but rendering never goes through them: a subclass overriding one is rejected above, not executed


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_no_cpu_fallback():
    cam = jr.Renderer.create_camera_from_parameters(jr.CameraParameters(viewWidth=8, viewHeight=8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        jr.render(cam, DepthShader, jr.Buffers(torch.ones(8, 8), ()), torch.zeros(1, 3, dtype=torch.int32),
                  DepthExtraInput(position=torch.zeros(3, 3)))


def test_camera_builders_match_closed_forms_and_oracle():
    eye, centre, up = torch.tensor((2.0, 4.0, 1.0)), torch.zeros(3), torch.tensor((0.0, 0.0, 1.0))
    view = jr.Camera.view_matrix(eye, centre, up)
    assert torch.allclose(view, O.view_matrix(eye, centre, up), atol=1e-6)
    assert torch.allclose(view @ jr.Camera.view_matrix_inv(eye, centre, up), torch.eye(4), atol=1e-5)
    assert torch.allclose((view @ torch.cat((eye, torch.ones(1))))[:3], torch.zeros(3), atol=1e-5)
    p = jr.Camera.perspective_projection_matrix(90.0, 2.0, 0.5, 10.0)
    f = 1 / math.tan(math.radians(45))
    assert torch.allclose(torch.diagonal(p)[:2], torch.tensor([f / 2, f]))
    assert float(p[3, 2]) == -1.0 and abs(float(p[2, 3]) - 2 * 10 * 0.5 / (0.5 - 10)) < 1e-6
    o = jr.Camera.orthographic_projection_matrix(-1.0, 1.0, -1.0, 1.0, -1.0, 1.0)
    assert torch.allclose(o, O.orthographic(-1.0, 1.0, -1.0, 1.0, -1.0, 1.0))
    vp = jr.Camera.viewport_matrix(torch.zeros(2), torch.tensor((640, 480)), 1.0)
    assert torch.allclose(vp @ torch.tensor([-1.0, -1.0, -1.0, 1.0]), torch.tensor([0.0, 0.0, 0.0, 1.0]))
    assert torch.allclose(vp @ torch.tensor([1.0, 1.0, 1.0, 1.0]), torch.tensor([640.0, 480.0, 1.0, 1.0]))
    cam = jr.Camera.create(view, p, vp)
    assert torch.allclose(cam.world_to_clip, p @ view) and torch.allclose(cam.world_to_eye_norm, torch.linalg.inv(view).T, atol=1e-6)
    assert torch.allclose(cam.screen_to_world @ cam.world_to_screen, torch.eye(4), atol=1e-3)
    # batched construction == stacked un-batched construction
    eyes = torch.stack((eye, eye * 0.5 + 1))
    vb = jr.Camera.view_matrix(eyes, centre, up)
    assert torch.allclose(vb[1], jr.Camera.view_matrix(eyes[1], centre, up), atol=1e-6)
    camp = jr.CameraParameters(position=eyes)
    cb = jr.Renderer.create_camera_from_parameters(camp)
    c1 = jr.Renderer.create_camera_from_parameters(jr.CameraParameters(position=eyes[1]))
    assert cb.world_to_clip.shape == (2, 4, 4) and torch.allclose(cb.world_to_clip[1], c1.world_to_clip, atol=1e-6)


def test_merge_objects_and_shapes_and_fixture():
    tex = torch.rand(4, 4, 3)
    cube = jr.create_cube(torch.tensor((1.0, 2.0, 3.0)), torch.ones(2), tex, torch.ones(4, 4))
    cap = jr.create_capsule(0.1, 0.25, jr.UpAxis.Z, torch.rand(1, 1, 3), torch.ones(1, 1))
    assert cube.verts.shape == (24, 3) and cube.faces.shape == (12, 3)
    assert cap.verts.shape == (576, 3) and cap.faces.shape == (192, 3)
    assert abs(float(cap.verts[:, 2].max()) - 0.35) < 1e-6 and abs(float(cap.verts[:, 0].abs().max()) - 0.1) < 1e-3
    t = torch.eye(4); t[:3, 3] = torch.tensor((1.0, 0.0, 0.5))
    m = jr.merge_objects([jr.ModelObject(model=cube), jr.ModelObject(model=cap, transform=t)])
    assert m.verts.shape == (600, 3) and m.faces.shape == (204, 3) and int(m.faces.max()) == 599
    assert m.diffuse_map.shape == (8, 4, 3) and m.offset == 4 and m.texture_shape.tolist() == [[4, 4], [1, 1]]
    assert torch.equal(m.diffuse_map[:4], tex) and float(m.diffuse_map[5:].abs().max()) == 0.0
    assert m.texture_index[:24].eq(0).all() and m.texture_index[24:].eq(1).all()
    assert torch.allclose(m.verts[24:], cap.verts + t[:3, 3])
    bm = jr.batch_models([m, m])
    assert bm.verts.shape == (2, 600, 3) and bm.offset == 4
    objs, camp = load_brax_fixture()
    mm = jr.merge_objects(objs)
    assert mm.verts.shape == (4, 9816, 3) and mm.faces.shape[-2:] == (3276, 3)
    assert mm.diffuse_map.shape[-3:] == (1800, 100, 3) and mm.specular_map.shape[-2:] == (18, 1)
    assert camp.viewWidth.shape == (4,)


def test_uv_repeat_and_shadow_get_host_twins():
    uv = torch.tensor([[0.25, -0.25], [1.5, 2.0]])
    out = jr.MergedModel.uv_repeat(uv, torch.tensor([10, 20]), torch.tensor(2), 100)
    assert torch.allclose(out, torch.tensor([[202.5, 15.0], [205.0, 0.0]]))
    sh = jr.Shadow(shadow_map=torch.arange(12.0).reshape(3, 4), strength=None, camera=None)
    pos = torch.tensor([[0.5, 1.49], [-1.0, 0.0], [2.6, 0.0], [0.49999997, 3.0]])
    assert sh.get(pos).tolist() == [5.0, 8.0, float("inf"), 3.0]
    assert torch.equal(sh.get(pos), O.shadow_get(sh.shadow_map, pos))


def test_transpose_for_display():
    a = torch.arange(24.0).reshape(2, 3, 4)
    d = jr.transpose_for_display(a)
    assert d.shape == (3, 2, 4) and torch.equal(d[0], a[:, 2]) and torch.equal(jr.transpose_for_display(a, False)[0], a[:, 0])


def test_shape_validation_errors_before_any_launch():
    cam = jr.Renderer.create_camera_from_parameters(jr.CameraParameters(viewWidth=8, viewHeight=8))
    f13 = torch.zeros(1, 3, dtype=torch.int32)
    cases = [
        (dict(position=torch.zeros(3, 2)), f13, torch.ones(8, 8), "must end in shape"),
        (dict(position=torch.zeros(2, 3, 3)), torch.zeros(3, 1, 3, dtype=torch.int32), torch.ones(8, 8), "inconsistent batch"),
        (dict(position=torch.zeros(3, 3)), torch.zeros(1, 4, dtype=torch.int32), torch.ones(8, 8), "must end in shape"),
        (dict(position=torch.zeros(3, 3)), f13, torch.ones(8), "has rank"),
    ]
    for kw, faces, z, msg in cases:
        with pytest.raises(ValueError, match=msg):
            jr.render(cam, DepthShader, jr.Buffers(z, ()), faces, DepthExtraInput(**kw))
    with pytest.raises(ValueError, match="targets must be"):
        jr.render(cam, DepthShader, jr.Buffers(torch.ones(8, 8), (torch.ones(8, 8, 3),)), f13,
                  DepthExtraInput(torch.zeros(3, 3)))
    with pytest.raises(TypeError, match="do not carry the fields"):
        jr.render(cam, GouraudShader, jr.Buffers(torch.ones(8, 8), (torch.ones(8, 8, 3),)), f13,
                  DepthExtraInput(torch.zeros(3, 3)))


@pytest.mark.skipif(not os.path.exists("/root/reference/test_resources/pre-gen-brax/inputs-2.zip"),
                    reason="reference checkout not present (GPU box)")
def test_brax_pregen_loader_matches_committed_fixture():
    from jaxrenderer_b200.brax_io import load_pregen
    objs, cam, targets = load_pregen("/root/reference/test_resources/pre-gen-brax/inputs-30.zip")
    fix_objs, fix_cam = load_brax_fixture()      # frames 0, 7, 15, 29 extracted at build time
    assert len(objs) == 18 and targets.shape == (30, 3)
    assert objs[0].model.verts.shape == (24, 3) and objs[1].model.faces.shape == (192, 3)
    frames = [0, 7, 15, 29]
    for o, f in zip(objs, fix_objs):
        assert torch.equal(o.transform[frames], f.transform) and torch.equal(o.model.verts, f.model.verts)
    assert torch.equal(cam.position[frames], fix_cam.position)
    m = jr.merge_objects(objs)
    assert m.verts.shape == (30, 9816, 3) and m.faces.shape[-2:] == (3276, 3)


def test_brax_unpickler_refuses_everything_but_the_exact_allow_list():
    """the loader's unpickler must not expose builtins / numpy helpers that execute code."""
    import io
    import pickle

    from jaxrenderer_b200 import brax_io

    class Evil:
        def __reduce__(self):
            return (eval, ("__import__('os').getpid()",))

    class EvilNumpy:
        def __reduce__(self):
            import numpy
            return (numpy.load, ("/nonexistent",))

    for payload in (Evil(), EvilNumpy(), getattr, print):
        with pytest.raises(pickle.UnpicklingError, match="refusing"):
            brax_io._Unpickler(io.BytesIO(pickle.dumps(payload))).load()
    # an attacker-chosen constructor handed to the jax array hook is refused too
    with pytest.raises(pickle.UnpicklingError):
        brax_io._reconstruct_array(eval, ("1",), None, None)
    # plain numpy arrays still load
    import numpy as np
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    b = brax_io._Unpickler(io.BytesIO(pickle.dumps(a))).load()
    assert np.array_equal(a, b)


def test_graft_entry_check_matches_header():
    """The driver's build hook verifies ABI version + exports against the header (without recompiling here)."""
    import __graft_entry__ as g

    g.check()


def test_graphed_rejects_what_it_cannot_capture():
    """`jr.graphed` (CUDA-graph counterpart of the reference users' `jax.jit`): argument validation needs no GPU."""
    g = jr.graphed(lambda x: x * 2)
    with pytest.raises(ValueError, match="CUDA tensors"):
        g(torch.ones(3))
    with pytest.raises(ValueError, match="no tensor"):
        g(3.0)
    assert g.cache_size() == 0
    deco = jr.graphed(copy_outputs=True)(lambda x: x)
    assert deco.copy_outputs and deco.cache_size() == 0
