"""ctypes binding of ``lib/libjr_b200.so`` (C ABI in ``include/jr_b200.h``).

The library is the product: importing it fails loudly when it is missing --
there is no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JR_B200_LIB") or os.path.join(_HERE, "lib", "libjr_b200.so")  # env: A/B builds

JR_DEPTH, JR_GOURAUD, JR_GOURAUD_TEXTURE, JR_PHONG, JR_PHONG_DARBOUX = 0, 1, 2, 3, 4
JR_PHONG_REFLECTION, JR_PHONG_REFLECTION_SHADOW = 5, 6

EXPORTS = (
    "jr_abi_version", "jr_strerror", "jr_workspace_bytes", "jr_backward_workspace_bytes",
    "jr_render_forward", "jr_render_backward",
    "jr_depth_forward", "jr_gouraud_forward", "jr_gouraud_texture_forward", "jr_phong_forward",
    "jr_phong_darboux_forward", "jr_phong_reflection_forward", "jr_phong_reflection_shadow_forward",
    "jr_add_scalar", "jr_canvas_to_uint8_display", "jr_launch_count", "jr_merge_objects", "jr_camera_build",
    "jr_instance_norm_scales", "jr_camera_vjp", "jr_debug_audit_cull",
    "jr_debug_kernel_timing", "jr_debug_kernel_times",
)


class JrF32(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("batch_stride", C.c_longlong)]


class JrI32(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("batch_stride", C.c_longlong)]


class JrRenderArgs(C.Structure):
    _fields_ = [
        ("shader", C.c_int32), ("B", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
        ("T", C.c_int32), ("n_pos", C.c_int32), ("n_nrm", C.c_int32), ("n_uv", C.c_int32),
        ("world_to_clip", JrF32), ("viewport", JrF32), ("world_to_eye_norm", JrF32),
        ("position", JrF32), ("faces", JrI32), ("normal", JrF32), ("faces_norm", JrI32),
        ("uv", JrF32), ("faces_uv", JrI32), ("colour", JrF32),
        ("light_direction", JrF32), ("light_colour", JrF32), ("light_dir_eye", JrF32),
        ("ambient", JrF32), ("diffuse", JrF32), ("specular", JrF32),
        ("texture", JrF32), ("tex_w", C.c_int32), ("tex_h", C.c_int32),
        ("specular_map", JrF32), ("spec_w", C.c_int32), ("spec_h", C.c_int32),
        ("normal_map", JrF32), ("texture_shape", JrI32), ("n_objects", C.c_int32),
        ("texture_index", JrI32), ("faces_tex", JrI32), ("n_texidx", C.c_int32),
        ("texture_offset", C.c_int32),
        ("id_to_face", JrI32), ("faces_indices", JrI32), ("n_faces_indices", C.c_int32),
        ("shadow_map", JrF32), ("shadow_w", C.c_int32), ("shadow_h", C.c_int32),
        ("shadow_strength", JrF32), ("shadow_world_to_clip", JrF32), ("shadow_viewport", JrF32),
        ("zbuffer", C.c_void_p), ("canvas", C.c_void_p), ("tri_id", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("inst_vert_object", JrI32), ("inst_scaling", JrF32), ("inst_transform", JrF32),
        ("inst_norm_object", JrI32), ("inst_normal_matrix", JrF32), ("inst_norm_scale", JrF32),
        ("n_inst", C.c_int32),
        ("canvas_u8", C.c_void_p), ("canvas_u8_background", C.c_float * 3),
        ("depth_offset", C.c_float), ("depth_fill", C.c_int32), ("depth_fill_value", C.c_float),
        ("stats", C.c_void_p),
    ]


class JrGradArgs(C.Structure):
    _fields_ = [
        ("d_zbuffer", C.c_void_p), ("d_canvas", C.c_void_p),
        ("d_position", JrF32), ("d_normal", JrF32), ("d_colour", JrF32),
        ("d_world_to_clip", JrF32), ("d_viewport", JrF32), ("d_world_to_eye_norm", JrF32),
        ("d_light_direction", JrF32), ("d_light_colour", JrF32), ("d_light_dir_eye", JrF32),
        ("d_ambient", JrF32), ("d_diffuse", JrF32), ("d_specular", JrF32),
        ("d_texture", JrF32), ("d_specular_map", JrF32), ("d_shadow_strength", JrF32),
        ("d_uv", JrF32), ("d_normal_map", JrF32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("no_buffer_grads", C.c_int32),
    ]


class JrMergeArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("n_objects", C.c_int32), ("n_verts", C.c_int32), ("n_norms", C.c_int32),
        ("local_verts", JrF32), ("local_norms", JrF32), ("vert_object", JrI32), ("norm_start", JrI32),
        ("scaling", JrF32), ("transform", JrF32), ("normal_matrix", JrF32),
        ("out_verts", C.c_void_p), ("out_norms", C.c_void_p),
    ]


class JrCameraArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("mode", C.c_int32), ("params", JrF32), ("viewport", JrF32),
                ("out", C.c_void_p)]


JR_CAMERA_PERSPECTIVE, JR_CAMERA_LIGHT = 0, 1


class JrKernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 24), ("ms", C.c_float), ("call", C.c_int32)]


JR_KERNEL_TIMES_MAX = 4096

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "jaxrenderer_b200/csrc/build.sh). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.jr_abi_version.restype = C.c_int
    lib.jr_strerror.restype = C.c_char_p
    lib.jr_strerror.argtypes = [C.c_int]
    lib.jr_workspace_bytes.restype = C.c_size_t
    lib.jr_workspace_bytes.argtypes = [C.POINTER(JrRenderArgs)]
    for name in ("jr_render_forward", "jr_depth_forward", "jr_gouraud_forward",
                 "jr_gouraud_texture_forward", "jr_phong_forward", "jr_phong_darboux_forward",
                 "jr_phong_reflection_forward", "jr_phong_reflection_shadow_forward"):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(JrRenderArgs), C.c_void_p]
    if hasattr(lib, "jr_render_backward"):
        lib.jr_render_backward.restype = C.c_int
        lib.jr_render_backward.argtypes = [C.POINTER(JrRenderArgs), C.POINTER(JrGradArgs), C.c_void_p]
        lib.jr_backward_workspace_bytes.restype = C.c_size_t
        lib.jr_backward_workspace_bytes.argtypes = [C.POINTER(JrRenderArgs), C.POINTER(JrGradArgs)]
    lib.jr_add_scalar.restype = C.c_int
    lib.jr_add_scalar.argtypes = [C.c_void_p, C.c_longlong, C.c_float, C.c_void_p]
    lib.jr_canvas_to_uint8_display.restype = C.c_int
    lib.jr_canvas_to_uint8_display.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.jr_launch_count.restype = C.c_longlong
    lib.jr_merge_objects.restype = C.c_int
    lib.jr_merge_objects.argtypes = [C.POINTER(JrMergeArgs), C.c_void_p]
    lib.jr_instance_norm_scales.restype = C.c_int
    lib.jr_instance_norm_scales.argtypes = [C.POINTER(JrMergeArgs), C.c_void_p, C.c_void_p]
    lib.jr_camera_build.restype = C.c_int
    lib.jr_camera_build.argtypes = [C.POINTER(JrCameraArgs), C.c_void_p]
    lib.jr_debug_audit_cull.restype = C.c_int
    lib.jr_debug_audit_cull.argtypes = [C.POINTER(JrRenderArgs), C.c_void_p, C.c_void_p]
    lib.jr_camera_vjp.restype = C.c_int
    lib.jr_camera_vjp.argtypes = [C.POINTER(JrCameraArgs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.jr_debug_kernel_timing.restype = C.c_int
    lib.jr_debug_kernel_timing.argtypes = [C.c_int]
    lib.jr_debug_kernel_times.restype = C.c_int
    lib.jr_debug_kernel_times.argtypes = [C.POINTER(JrKernelTime), C.c_int]
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        raise RuntimeError(f"libjr_b200: {load().jr_strerror(status).decode()} (status {status})")


def stream_ptr(device: Any) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(load().jr_launch_count())


def kernel_timing(enable: bool) -> None:
    """Switch the library's per-kernel event timing (measurement aid, ``jr_debug_kernel_timing``)."""
    check(load().jr_debug_kernel_timing(1 if enable else 0))


def kernel_times() -> list:
    """``[(kernel name, ms, call index), ...]`` in launch order for every launch since the switch-on / the last
    read (waits for them: ``jr_debug_kernel_times``)."""
    buf = (JrKernelTime * JR_KERNEL_TIMES_MAX)()
    n = load().jr_debug_kernel_times(buf, JR_KERNEL_TIMES_MAX)
    if n < 0:
        check(n)
    return [(buf[i].name.decode(), float(buf[i].ms), int(buf[i].call)) for i in range(min(n, JR_KERNEL_TIMES_MAX))]
