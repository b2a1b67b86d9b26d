"""The XLA-FFI side of the boundary (ffi/jr_ffi.cc + ffi/jax_binding.py) cannot be built or run here (no jax / jaxlib).
These tests keep the sources honest: the C++ compiles against a stub of the FFI API, it declares 7 forward + 7 backward
targets, and its operand tables agree with the Python binding's."""
import ast
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CC = os.path.join(ROOT, "ffi", "jr_ffi.cc")
PY = os.path.join(ROOT, "ffi", "jax_binding.py")
SHADERS = ("depth", "gouraud", "gouraud_texture", "phong", "phong_darboux", "phong_reflection", "phong_reflection_shadow")


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_ffi_handlers_compile_against_the_api_stub():
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-DJR_FFI_STUB", "-I", os.path.join(ROOT, "ffi"), CC], check=True)
    # without XLA headers the file is an empty translation unit (no accidental dependency)
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", CC], check=True)


def _cc_tables():
    src = open(CC).read()
    tables = {}
    for m in re.finditer(r"constexpr Operand (k\w+)\[\] = \{(.*?)\};", src, re.S):
        tables[m.group(1)] = [(n, int(r), f == "true", d == "true")
                              for n, r, f, d in re.findall(r'\{"(\w+)",\s*(\d),\s*(true|false),\s*(true|false)\}', m.group(2))]
    return src, tables


def test_ffi_targets_and_operand_tables_agree_with_the_python_binding():
    src, t = _cc_tables()
    for s in SHADERS:
        assert re.search(rf"JR_FFI_TARGETS\({s},", src), s
    import typing
    ns = {"Dict": typing.Dict, "Tuple": typing.Tuple}
    tree = ast.parse(open(PY).read())
    for node in tree.body:      # evaluate only the literal tables
        if isinstance(node, (ast.Assign, ast.AnnAssign)):
            names = [node.target.id] if isinstance(node, ast.AnnAssign) else [x.id for x in node.targets if isinstance(x, ast.Name)]
            if names and names[0] in ("COMMON", "PHONG_REFLECTION", "OPERANDS"):
                exec(compile(ast.Module([node], []), PY, "exec"), ns)
    ops = ns["OPERANDS"]
    assert tuple(ops) == SHADERS
    cc = {"depth": t["kCommon"], "gouraud": t["kCommon"] + t["kGouraud"],
          "gouraud_texture": t["kCommon"] + t["kGouraudTexture"], "phong": t["kCommon"] + t["kPhong"],
          "phong_darboux": t["kCommon"] + t["kPhongDarboux"], "phong_reflection": t["kCommon"] + t["kPhongReflection"],
          "phong_reflection_shadow": t["kCommon"] + t["kPhongReflection"] + t["kShadow"]}
    for s in SHADERS:
        want = [(n, r, d == "f32", g) for n, r, d, g in ops[s]]
        assert cc[s] == want, s
    # every differentiable operand has a gradient slot in the ABI
    header = open(os.path.join(ROOT, "include", "jr_b200.h")).read()
    for s in SHADERS:
        for n, _, _, g in cc[s]:
            if g:
                assert re.search(rf"\bd_{n}\b", header), n
