"""Drop-in check: names, parameter lists, NamedTuple fields and literal defaults of the public API on the
path equal the reference's (``tests/golden/reference_api.json``, extracted from the reference sources by
``tools/extract_reference_api.py``).  Extra keyword-only parameters of this package (``inplace``,
``return_tri_id``, ``device`` ...) are allowed after the reference's own."""
import inspect
import json
import os

import jaxrenderer_b200 as jr
from jaxrenderer_b200 import geometry, model, shaders, shadow, utils

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_api.json")))


def _params(fn):
    fn = fn.__func__ if isinstance(fn, (classmethod, staticmethod)) else fn
    return [p for p in inspect.signature(fn).parameters if p not in ("self", "cls")]


def _check_prefix(name, ours, ref):
    assert ours[:len(ref)] == ref, f"{name}: parameters {ours} do not start with the reference's {ref}"


def test_function_signatures():
    where = {"render": jr, "merge_objects": jr, "batch_models": jr, "create_cube": jr, "create_capsule": jr,
             "transpose_for_display": utils, "build_texture_from_PyTinyrenderer": utils, "normalise": geometry,
             "quaternion": geometry, "quaternion_mul": geometry, "rotation_matrix": geometry}
    for name, ref in GOLDEN["functions"].items():
        ours = _params(getattr(where[name], name))
        if name == "quaternion_mul":          # positional-only in spirit: argument names differ (a, b)
            assert len(ours) == len(ref)
            continue
        _check_prefix(name, ours, ref)


def test_method_signatures():
    classes = {"Renderer": jr.Renderer, "Shadow": shadow.Shadow, "Camera": geometry.Camera,
               "ModelObject": model.ModelObject, "Model": model.Model, "MergedModel": model.MergedModel}
    for qual, ref in GOLDEN["methods"].items():
        cls, meth = qual.split(".")
        fn = inspect.getattr_static(classes[cls], meth)
        _check_prefix(qual, _params(fn), ref)


def test_namedtuple_fields_and_defaults():
    classes = {"CameraParameters": jr.CameraParameters, "LightParameters": jr.LightParameters,
               "ShadowParameters": jr.ShadowParameters, "Model": jr.Model, "MergedModel": jr.MergedModel,
               "ModelObject": jr.ModelObject, "Buffers": jr.Buffers, "LightSource": jr.LightSource,
               "Shadow": jr.Shadow, "Camera": jr.Camera}
    for name in GOLDEN["tuples"]:
        if name.endswith("ExtraInput"):
            classes[name] = getattr(shaders, name)
    for name, fields in GOLDEN["tuples"].items():
        cls = classes[name]
        assert list(cls._fields) == [f for f, _ in fields], (name, cls._fields)
        defaults = getattr(cls, "_field_defaults", {})
        for f, d in fields:
            if d is None:
                continue
            got = defaults.get(f)
            got = list(got) if isinstance(got, tuple) else got
            assert got == d, f"{name}.{f}: default {got!r} != reference {d!r}"


def test_literal_parameter_defaults():
    where = {"render": jr, "transpose_for_display": utils}
    classes = {"Renderer": jr.Renderer, "Shadow": shadow.Shadow, "Camera": geometry.Camera,
               "ModelObject": model.ModelObject, "Model": model.Model, "MergedModel": model.MergedModel}
    for qual, defaults in GOLDEN["defaults"].items():
        if not defaults:
            continue
        if "." in qual:
            cls, meth = qual.split(".")
            fn = inspect.getattr_static(classes[cls], meth)
            fn = fn.__func__ if isinstance(fn, (classmethod, staticmethod)) else fn
        else:
            fn = getattr(where[qual], qual)
        sig = inspect.signature(fn).parameters
        for name, value in defaults.items():
            assert sig[name].default == value, f"{qual}({name}=...): {sig[name].default!r} != reference {value!r}"
