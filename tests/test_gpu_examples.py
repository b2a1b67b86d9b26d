"""The reference's `examples/sphere.py` and `examples/plane_as_mesh.py` (minus the `Scene` registry, which is outside
the path: objects are built with `create_capsule` / `Model.create` + `ModelObject` directly), with the argument idioms
those scripts use -- tuples for camera position / target and for the PyTinyrenderer texture, `Model.create` defaults, a
tuple quaternion, `ShadowParameters(offset=0.05)`, a capsule of half height 0 (a sphere) along X -- rendered by the
CUDA path at a quarter of the examples' 640x480 and compared with the CPU oracle's facade."""
import pytest
import torch

import jaxrenderer_b200 as jr
from oracle import jr_oracle as O

pytestmark = pytest.mark.gpu
W, H = 160, 120
RGBW = (255, 255, 255, 255, 0, 0, 0, 255, 0, 0, 0, 255)     # white, red, green, blue


def _compare(tag, objs_fn, light, cp, sp):
    dev = torch.device("cuda", 0)
    img = jr.Renderer.get_camera_image(objects=objs_fn(dev), light=light, camera=cp, width=W, height=H, shadow_param=sp)
    merged = jr.merge_objects(objs_fn(None))
    cam = jr.Renderer.create_camera_from_parameters(cp)
    res = O.renderer_render(merged, light, cam, torch.ones(W, H), torch.ones(W, H, 3), shadow_param=sp)
    want, gap = res["out"].targets[0], res["out"].gap
    diff = (img.cpu() - want).abs().amax(-1)
    bad = diff > 2e-5
    covered = int((res["out"].zbuffer != 1.0).sum())
    print(f"{tag}: covered {covered} px, max |dcolour| {float(diff.max()):.3g}, pixels off {int(bad.sum())}, "
          f"of which depth ties {int((bad & (gap < 1e-6)).sum())}")
    assert img.shape == (W, H, 3) and covered > 0.03 * W * H
    assert int((bad & ~(gap < 1e-6)).sum()) == 0
    rgb = jr.transpose_for_display(torch.clamp(img * 255, 0.0, 255.0).to(torch.uint8))      # the scripts' last lines
    assert rgb.shape == (H, W, 3) and rgb.dtype == torch.uint8


def test_example_sphere():
    tex = jr.build_texture_from_PyTinyrenderer(torch.tensor(RGBW), 2, 2) / 255.0
    light = jr.LightParameters(direction=torch.tensor([2.0, 4.0, 1.0]), ambient=torch.zeros(3), diffuse=torch.full((3,), 1.0),
                               specular=torch.full((3,), 0.0))
    cp = jr.CameraParameters(viewWidth=W, viewHeight=H, position=torch.tensor([2.0, 4.0, 1.0]), target=torch.tensor([0.0, 0.0, 0.0]))

    def objs(dev):
        m = jr.create_capsule(radius=torch.tensor(1.0), half_height=torch.tensor(0.0), up_axis=jr.UpAxis.X, diffuse_map=tex,
                              specular_map=torch.full(tex.shape[:2], 2.0))      # what Scene.add_capsule passes
        if dev is not None:
            m = type(m)(*[t.to(dev) for t in m])
        return [jr.ModelObject(model=m)]
    _compare("sphere.py", objs, light, cp, jr.ShadowParameters(offset=0.05))


def test_example_plane_as_mesh():
    tex = jr.build_texture_from_PyTinyrenderer(RGBW, 2, 2) / 255.0
    verts = torch.tensor([[100.0, -100.0, 0.0], [100.0, 100.0, 0.0], [-100.0, 100.0, 0.0], [-100.0, -100.0, 0.0]]) * 0.01
    norms = torch.tensor([[0.0, 0.0, 1.0]] * 4)
    uvs = torch.tensor([[1.0, 0.0], [1.0, 1.0], [0.0, 1.0], [0.0, 0.0]])
    faces = torch.tensor([[0, 1, 2], [0, 2, 3]])
    cp = jr.CameraParameters(viewWidth=W, viewHeight=H, position=(2.0, 4.0, 1.0), target=(0.0, 0.0, 0.0))

    def objs(dev):
        m = jr.Model.create(verts=verts, norms=norms, uvs=uvs, faces=faces, diffuse_map=tex)
        if dev is not None:
            m = type(m)(*[t.to(dev) for t in m])
        return [jr.ModelObject(model=m).replace_with_orientation((1.0, 0, 0, 0))]
    _compare("plane_as_mesh.py", objs, jr.LightParameters(), cp, jr.ShadowParameters())


def test_example_axises_and_plane_growing_object_lists():
    """`examples/axises_and_plane.py`: a flat cube (texture scaling 16) and capsules along X, Y, Z, rendered four times
    with a GROWING list of objects -- each list is its own atlas / merged topology (the memoised static half of
    `merge_objects` must not leak from one list into the next)."""
    tex = jr.build_texture_from_PyTinyrenderer(RGBW, 2, 2) / 255.0
    spec = torch.full(tex.shape[:2], 2.0)
    cp = jr.CameraParameters(viewWidth=W, viewHeight=H, position=(2.0, 4.0, 1.0), target=(0.0, 0.0, 0.0))

    def all_objs(dev):
        mv = (lambda m: type(m)(*[t.to(dev) for t in m])) if dev is not None else (lambda m: m)
        cube = mv(jr.create_cube(half_extents=torch.tensor((1.5, 1.5, 0.03)), texture_scaling=torch.tensor((16.0, 16.0)),
                                 diffuse_map=tex, specular_map=spec))
        caps = [mv(jr.create_capsule(radius=torch.tensor(0.1), half_height=torch.tensor(0.4), up_axis=ax, diffuse_map=tex,
                                     specular_map=spec)) for ax in (jr.UpAxis.X, jr.UpAxis.Y, jr.UpAxis.Z)]
        return [jr.ModelObject(model=cube).replace_with_position((0.0, 0.0, -0.5))] + [jr.ModelObject(model=c) for c in caps]

    cuda_objs, cpu_objs = all_objs(torch.device("cuda", 0)), all_objs(None)
    for n in (2, 3, 4, 2):                                   # ... and back to the first list
        _compare(f"axises_and_plane.py, {n} objects", lambda dev, n=n: (cuda_objs if dev is not None else cpu_objs)[:n],
                 jr.LightParameters(), cp, jr.ShadowParameters())
