#!/usr/bin/env python
"""Extract the public API surface of the reference (names, parameter lists, NamedTuple fields and literal
defaults) for the drop-in path into ``tests/golden/reference_api.json``.  Parses the reference sources with
``ast`` -- nothing is imported (jax is not installed) and nothing is copied but names.

  python tools/extract_reference_api.py /root/reference
"""
from __future__ import annotations

import ast
import json
import os
import sys

FUNCS = {
    "renderer/pipeline.py": ["render"],
    "renderer/model.py": ["merge_objects", "batch_models"],
    "renderer/shapes/cube.py": ["create_cube"],
    "renderer/shapes/capsule.py": ["create_capsule"],
    "renderer/utils.py": ["transpose_for_display", "build_texture_from_PyTinyrenderer"],
    "renderer/geometry.py": ["normalise", "quaternion", "quaternion_mul", "rotation_matrix"],
}
METHODS = {
    "renderer/renderer.py": {"Renderer": ["create_camera_from_parameters", "create_buffers", "render",
                                          "get_camera_image"]},
    "renderer/shadow.py": {"Shadow": ["render_shadow_map", "get"]},
    "renderer/geometry.py": {"Camera": ["create", "view_matrix", "view_matrix_inv", "perspective_projection_matrix",
                                        "orthographic_projection_matrix", "viewport_matrix", "apply", "apply_pos",
                                        "apply_vec", "to_screen", "to_clip"]},
    "renderer/model.py": {"ModelObject": ["replace_with_position", "replace_with_orientation",
                                          "replace_with_local_scaling", "replace_with_double_sided"],
                          "Model": ["create"],
                          "MergedModel": ["generate_object_vert_info", "merge_verts", "merge_maps", "uv_repeat"]},
}
TUPLES = {
    "renderer/renderer.py": ["CameraParameters", "LightParameters", "ShadowParameters"],
    "renderer/model.py": ["Model", "MergedModel", "ModelObject"],
    "renderer/types.py": ["Buffers", "LightSource"],
    "renderer/shadow.py": ["Shadow"],
    "renderer/geometry.py": ["Camera"],
    "renderer/shaders/depth.py": ["DepthExtraInput"],
    "renderer/shaders/gouraud.py": ["GouraudExtraInput"],
    "renderer/shaders/gouraud_texture.py": ["GouraudTextureExtraInput"],
    "renderer/shaders/phong.py": ["PhongTextureExtraInput"],
    "renderer/shaders/phong_darboux.py": ["PhongTextureDarbouxExtraInput"],
    "renderer/shaders/phong_reflection.py": ["PhongReflectionTextureExtraInput"],
    "renderer/shaders/phong_reflection_shadow.py": ["PhongReflectionShadowTextureExtraInput"],
}


def params(fn: ast.FunctionDef):
    a = fn.args
    names = [p.arg for p in a.posonlyargs + a.args + a.kwonlyargs]
    return [n for n in names if n not in ("self", "cls")]


def param_defaults(fn: ast.FunctionDef):
    """{name: literal default} for parameters whose default is a plain literal."""
    a = fn.args
    pos = a.posonlyargs + a.args
    out = {}
    for p_, d in zip(pos[len(pos) - len(a.defaults):], a.defaults):
        v = literal(d)
        if isinstance(v, tuple):
            v = list(v)
        if isinstance(v, (int, float, bool, list)):
            out[p_.arg] = v
    for p_, d in zip(a.kwonlyargs, a.kw_defaults):
        v = literal(d) if d is not None else None
        if isinstance(v, (int, float, bool)):
            out[p_.arg] = v
    return out


def literal(node):
    try:
        return ast.literal_eval(node)
    except Exception:
        return None


def main(root: str) -> None:
    out = {"functions": {}, "methods": {}, "tuples": {}, "defaults": {}}
    for rel in sorted(set(FUNCS) | set(METHODS) | set(TUPLES)):
        tree = ast.parse(open(os.path.join(root, rel)).read())
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in FUNCS.get(rel, []):
                out["functions"][node.name] = params(node)
                out["defaults"][node.name] = param_defaults(node)
            if isinstance(node, ast.ClassDef):
                if node.name in METHODS.get(rel, {}):
                    for sub in node.body:
                        if isinstance(sub, ast.FunctionDef) and sub.name in METHODS[rel][node.name]:
                            out["methods"][f"{node.name}.{sub.name}"] = params(sub)
                            out["defaults"][f"{node.name}.{sub.name}"] = param_defaults(sub)
                if node.name in TUPLES.get(rel, []):
                    fields = []
                    for sub in node.body:
                        if isinstance(sub, ast.AnnAssign) and isinstance(sub.target, ast.Name):
                            default = literal(sub.value) if sub.value is not None else None
                            if isinstance(default, tuple):
                                default = list(default)
                            fields.append([sub.target.id, default if isinstance(default, (int, float, list, bool))
                                           else None])
                    out["tuples"][node.name] = fields
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                       "reference_api.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print("wrote", dst, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
