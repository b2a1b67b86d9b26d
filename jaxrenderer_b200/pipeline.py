"""``pipeline.render`` -- the drop-in boundary (``renderer/pipeline.py:470-537``).

Same call as the reference::

    render(camera, shader, buffers, face_indices, extra, loop_unroll=1) -> Buffers

but the body is one call into ``libjr_b200.so`` (``jr_render_forward`` /
``jr_render_backward``, ``include/jr_b200.h``): a tiled visibility kernel and a
fused interpolate + fragment + mix kernel per built-in shader.  Differences a
caller can observe:

* batching is native -- any leaf of ``camera`` / ``buffers`` / ``extra`` /
  ``face_indices`` may carry one extra leading axis (what ``jax.vmap`` would
  add, ``examples/batch_rendering.py:87-95``); un-batched leaves broadcast;
* ``shader`` must be one of the seven built-in classes, anything else raises
  ``UnsupportedShaderError`` (no fallback);
* ``loop_unroll`` is accepted and ignored (it never changed results,
  ``changelog.md:57``);
* gradients flow through ``torch.autograd`` (custom backward kernel through
  fixed visibility), the reference's ``jax.grad``;
* host (CPU) tensors are copied to the current CUDA device and the result is
  copied back, like JAX's implicit ``device_put``.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from . import _native
from ._native import JrF32, JrGradArgs, JrI32, JrRenderArgs
from .geometry import Camera
from .model import InstancedArray
from .shader import Shader, UnsupportedShaderError
from .shaders import BUILTIN_SHADERS
from .types import Buffers, Tensor, _f32, _i32, _is_vmapped

# name -> (un-batched rank, is_float)
_SPEC: Dict[str, Tuple[int, bool]] = {
    "world_to_clip": (2, True), "viewport": (2, True), "world_to_eye_norm": (2, True),
    "position": (2, True), "faces": (2, False), "normal": (2, True), "faces_norm": (2, False),
    "uv": (2, True), "faces_uv": (2, False), "colour": (2, True),
    "light_direction": (1, True), "light_colour": (1, True), "light_dir_eye": (1, True),
    "ambient": (1, True), "diffuse": (1, True), "specular": (1, True),
    "texture": (3, True), "specular_map": (2, True), "normal_map": (3, True),
    "texture_shape": (2, False), "texture_index": (1, False), "faces_tex": (2, False),
    "id_to_face": (1, False), "faces_indices": (2, False),
    "shadow_map": (2, True), "shadow_strength": (1, True),
    "shadow_world_to_clip": (2, True), "shadow_viewport": (2, True),
    "zbuffer": (2, True), "canvas": (3, True),
}

# required trailing dimensions (None = free)
_TRAIL: Dict[str, Tuple[Optional[int], ...]] = {
    "world_to_clip": (4, 4), "viewport": (4, 4), "world_to_eye_norm": (4, 4),
    "shadow_world_to_clip": (4, 4), "shadow_viewport": (4, 4),
    "position": (None, 3), "normal": (None, 3), "colour": (None, 3), "uv": (None, 2),
    "faces": (None, 3), "faces_norm": (None, 3), "faces_uv": (None, 3), "faces_tex": (None, 3),
    "faces_indices": (None, 3), "texture_shape": (None, 2),
    "light_direction": (3,), "light_colour": (3,), "light_dir_eye": (3,), "ambient": (3,),
    "diffuse": (3,), "specular": (3,), "shadow_strength": (3,),
    "texture": (None, None, 3), "normal_map": (None, None, 3),
}

# differentiable inputs, in the positional order of _RenderFn
_DIFF = (
    "zbuffer", "canvas", "position", "normal", "colour", "world_to_clip", "viewport",
    "world_to_eye_norm", "light_direction", "light_colour", "light_dir_eye", "ambient",
    "diffuse", "specular", "texture", "specular_map", "shadow_strength",
    "uv", "normal_map",   # phong_darboux only (through the tangent frame); zero for the other shaders
)
_GRAD_FIELD = {n: "d_" + n for n in _DIFF if n not in ("zbuffer", "canvas")}


def _as(t: Any, is_float: bool, device: torch.device, written: bool = False) -> Tensor:
    """Device placement of one input.  ``written`` marks the buffers the kernels update in place: they
    never alias the memoised constants."""
    dt = torch.float32 if is_float else torch.int32
    if written and not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t, dtype=dt)
    if not written and (not isinstance(t, torch.Tensor) or (not t.is_cuda and t.numel() <= 64)):
        return (_f32 if is_float else _i32)(t, device).contiguous()   # small host constants: memoised copy
    if t.dtype != dt:
        t = t.to(dt)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def _collect(shader: type, camera: Any, face_indices: Any, extra: Any) -> Dict[str, Any]:
    """Map the reference's ``extra`` NamedTuple of a built-in shader onto the
    flat array names of ``JrRenderArgs``."""
    sid = shader._jr_shader
    arr: Dict[str, Any] = {
        "world_to_clip": camera.world_to_clip, "viewport": camera.viewport,
        "position": extra.position, "faces": face_indices,
    }
    if sid == _native.JR_DEPTH:
        return arr
    arr["normal"] = extra.normal
    arr["light_colour"] = extra.light.colour
    if sid == _native.JR_GOURAUD:
        arr["colour"] = extra.colour
        arr["light_direction"] = extra.light.direction
        return arr
    arr["uv"] = extra.uv
    arr["texture"] = extra.texture
    if sid >= _native.JR_PHONG:
        arr["world_to_eye_norm"] = camera.world_to_eye_norm
    if sid in (_native.JR_GOURAUD_TEXTURE, _native.JR_PHONG, _native.JR_PHONG_DARBOUX):
        arr["light_direction"] = extra.light.direction
    if sid == _native.JR_PHONG_DARBOUX:
        arr["normal_map"] = extra.normal_map
        arr["id_to_face"] = extra.id_to_face
        arr["faces_indices"] = extra.faces_indices
    if sid >= _native.JR_PHONG_REFLECTION:
        arr["light_dir_eye"] = extra.light_dir_eye
        arr["ambient"], arr["diffuse"], arr["specular"] = extra.ambient, extra.diffuse, extra.specular
        arr["specular_map"] = extra.specular_map
        arr["texture_shape"] = extra.texture_shape
        arr["texture_index"] = extra.texture_index
        arr["texture_offset"] = int(extra.texture_offset)
    if sid == _native.JR_PHONG_REFLECTION_SHADOW:
        sh = extra.shadow
        arr["shadow_map"] = sh.shadow_map
        arr["shadow_strength"] = sh.strength
        arr["shadow_world_to_clip"] = sh.camera.world_to_clip
        arr["shadow_viewport"] = sh.camera.viewport
    return arr


_STATS: Optional[Tensor] = None   # set by `visibility_stats()`: device uint64[8] the kernels add their counters to
STAT_NAMES = ("triangles", "filter_passed", "exact_kept", "n_test", "fragments", "exact_rounds", "batches")


class visibility_stats:
    """Context manager: forward renders inside it run the COUNTING variant of the visibility kernel
    (``JrRenderArgs.stats``, include/jr_b200.h) and accumulate ``n_test`` (edge-function evaluations; the
    reference evaluates W*H*T of them, ``pipeline.py:163-279``), survivors of the filter / exact cull, ...
    ``.read()`` synchronises and returns them as a dict.  Measurement aid only (single-tile canvases)."""

    def __init__(self, device: Any = "cuda"):
        self.buf = torch.zeros(8, dtype=torch.int64, device=device)

    def __enter__(self) -> "visibility_stats":
        global _STATS
        self._prev, _STATS = _STATS, self.buf
        return self

    def __exit__(self, *exc: Any) -> None:
        global _STATS
        _STATS = self._prev

    def read(self) -> Dict[str, int]:
        return dict(zip(STAT_NAMES, (int(v) for v in self.buf.cpu().tolist())))


def audit_cull(camera: Any, face_indices: Any, position: Any) -> Dict[str, int]:
    """Test aid (``jr_debug_audit_cull``, include/jr_b200.h): brute-force every pixel of every triangle with the exact
    edge functions and count what the conservative culls -- the filter phase of the single-tile kernel, the bbox
    margin of the exact phase / binned setup -- would have lost.  ``filter_lost_triangles``, ``pixels_outside_bbox``
    and ``pixels_of_rejected_triangles`` must be 0.  Costs T*W*H per image."""
    lib = _native.load()
    dev = position.device if isinstance(position, (torch.Tensor, InstancedArray)) else torch.device("cuda")
    W = int(round(float(torch.as_tensor(camera.viewport).reshape(-1, 16)[0, 0]) * 2))
    H = int(round(float(torch.as_tensor(camera.viewport).reshape(-1, 16)[0, 5]) * 2))
    arrays = {"world_to_clip": camera.world_to_clip, "viewport": camera.viewport, "position": position, "faces": face_indices}
    raw = {k: (v if isinstance(v, InstancedArray) else _as(torch.as_tensor(v), _SPEC[k][1], dev)) for k, v in arrays.items()}
    batched = {k: v.ndim == _SPEC[k][0] + 1 for k, v in raw.items()}
    B = max([v.shape[0] for k, v in raw.items() if batched[k]] + [1])
    call = _Call(_native.JR_DEPTH, raw, B, W, H, 0, batched)
    z = torch.empty((B, W, H), device=dev)
    args = call.fill(z, None, None)
    counters = torch.zeros(8, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _native.check(lib.jr_debug_audit_cull(C.byref(args), counters.data_ptr(), _native.stream_ptr(dev)))
    c = counters.cpu().tolist()
    return {"filter_lost_triangles": c[0], "pixels_outside_bbox": c[1], "pixels_of_rejected_triangles": c[2],
            "triangles_kept": c[3], "inside_pixels": c[4]}


class _Call:
    """Everything about one render call that is not a differentiable tensor."""

    def __init__(self, sid: int, arrays: Dict[str, Tensor], B: int, W: int, H: int,
                 texture_offset: int, batched: Dict[str, bool]):
        self.sid, self.arrays, self.B, self.W, self.H = sid, arrays, B, W, H
        self.texture_offset = texture_offset
        self.batched = batched
        self.tri_id: Optional[Tensor] = None
        self.depth_epilogue: Optional[Tuple[float, Optional[float]]] = None
        self.u8_out: Optional[Tensor] = None
        self.u8_background: Tuple[float, float, float] = (1.0, 1.0, 1.0)

    def stride(self, name: str) -> int:
        t = self.arrays[name]
        return int(t[0].numel()) if self.batched[name] else 0

    def _fill_instanced(self, a: JrRenderArgs) -> None:
        """``JrRenderArgs.inst_*``: local meshes + per-image object transforms."""
        g = self.arrays["position"].geom
        st = g.st

        def f32(t, r):
            return JrF32(t.data_ptr(), int(t[0].numel()) if t.ndim == r + 1 else 0)

        a.position = f32(st["local_verts"], 2)
        a.inst_vert_object = JrI32(st["vert_object"].data_ptr(), 0)
        a.inst_scaling, a.inst_transform = f32(g.scaling, 2), f32(g.transform, 3)
        a.n_inst = g.n_obj
        if "normal" in self.arrays:
            a.normal = f32(st["local_norms"], 2)
            a.inst_norm_object = JrI32(st["norm_object"].data_ptr(), 0)
            a.inst_normal_matrix = f32(g.nmat, 3)
            a.inst_norm_scale = f32(g.norm_scales(), 2)

    def fill(self, zbuffer: Tensor, canvas: Optional[Tensor], tri_id: Tensor) -> JrRenderArgs:
        a = JrRenderArgs()
        a.shader, a.B, a.W, a.H = self.sid, self.B, self.W, self.H
        A = self.arrays
        faces = A["faces"]
        a.T = faces.shape[-2]
        a.n_pos = A["position"].shape[-2]
        a.n_nrm = A["normal"].shape[-2] if "normal" in A else 0
        a.n_uv = A["uv"].shape[-2] if "uv" in A else 0
        inst = isinstance(A["position"], InstancedArray)
        for name, t in A.items():
            if name in ("zbuffer", "canvas") or isinstance(t, InstancedArray):
                continue
            is_float = _SPEC[name][1]
            setattr(a, name, (JrF32 if is_float else JrI32)(t.data_ptr(), self.stride(name)))
        if "texture" in A:
            a.tex_w, a.tex_h = A["texture"].shape[-3], A["texture"].shape[-2]
        if "specular_map" in A:
            a.spec_w, a.spec_h = A["specular_map"].shape[-2], A["specular_map"].shape[-1]
        if "texture_shape" in A:
            a.n_objects = A["texture_shape"].shape[-2]
            a.n_texidx = A["texture_index"].shape[-1]
            a.texture_offset = self.texture_offset
        if "faces_indices" in A:
            a.n_faces_indices = A["faces_indices"].shape[-2]
        if "shadow_map" in A:
            a.shadow_w, a.shadow_h = A["shadow_map"].shape[-2], A["shadow_map"].shape[-1]
        a.zbuffer = zbuffer.data_ptr()
        a.canvas = canvas.data_ptr() if canvas is not None else None
        a.tri_id = tri_id.data_ptr() if tri_id is not None else None
        if inst:
            self._fill_instanced(a)
        if self.u8_out is not None:
            a.canvas_u8 = self.u8_out.data_ptr()
            a.canvas_u8_background = (C.c_float * 3)(*self.u8_background)
        if self.depth_epilogue is not None and self.sid == _native.JR_DEPTH:
            off, fill = self.depth_epilogue
            a.depth_offset = off
            a.depth_fill, a.depth_fill_value = (0, 0.0) if fill is None else (1, fill)
        a.workspace, a.workspace_bytes = None, 0
        a.stats = _STATS.data_ptr() if (_STATS is not None and _STATS.device == zbuffer.device) else None
        return a


def _forward_native(call: _Call, zbuffer: Tensor, canvas: Optional[Tensor],
                    need_tri: bool = True) -> Optional[Tensor]:
    lib = _native.load()
    dev = zbuffer.device
    tri_id = None
    if need_tri or call.sid != _native.JR_DEPTH:
        tri_id = torch.empty((call.B, call.W, call.H), dtype=torch.int32, device=dev)
    args = call.fill(zbuffer, canvas, tri_id)
    need = lib.jr_workspace_bytes(C.byref(args))
    ws = None
    if need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        args.workspace, args.workspace_bytes = ws.data_ptr(), need
    with torch.cuda.device(dev):
        _native.check(lib.jr_render_forward(C.byref(args), _native.stream_ptr(dev)))
    call.tri_id = tri_id
    return tri_id


class _RenderFn(torch.autograd.Function):
    """custom_vjp: forward = visibility + shading kernels, backward = the
    gradient kernels through the saved triangle-id G-buffer."""

    # forward / setup_context are separate (the form torch.func transforms require: the renderer may be called
    # inside torch.func.vmap, see _render_arrays_vmapped)
    @staticmethod
    def forward(call: _Call, *diff: Optional[Tensor]):  # type: ignore[override]
        named = dict(zip(_DIFF, diff))
        zbuffer = named["zbuffer"].clone()
        canvas = named["canvas"].clone() if named["canvas"] is not None else None
        tri_id = _forward_native(call, zbuffer, canvas)
        if canvas is None:
            return zbuffer, tri_id
        return zbuffer, canvas, tri_id

    @staticmethod
    def vmap(info: Any, in_dims: Any, call: _Call, *diff: Optional[Tensor]):
        """torch.func.vmap rule.  The public entry points re-batch natively BEFORE reaching this Function
        (_render_arrays_vmapped: every mapped input, differentiable or not, becomes a batched array of ONE
        kernel call), so its operands are never mapped at the enclosing level and functorch lowers the call
        without consulting this rule; it exists because functorch insists on one, and to fail clearly if the
        Function is ever applied to mapped tensors directly."""
        raise NotImplementedError("_RenderFn applied to vmapped tensors: call jaxrenderer_b200.render / "
                                  "Renderer.render inside torch.func.vmap instead (they batch natively)")

    @staticmethod
    def setup_context(ctx: Any, inputs: Any, output: Any) -> None:
        call, *diff = inputs
        named = dict(zip(_DIFF, diff))
        ctx.call = call
        ctx.present = [t is not None for t in diff]
        ctx.save_for_backward(*[t for t in diff if t is not None and t is not named["zbuffer"]
                                and t is not named["canvas"]])
        ctx.mark_non_differentiable(output[-1])

    @staticmethod
    def backward(ctx: Any, *grads: Optional[Tensor]):  # type: ignore[override]
        call: _Call = ctx.call
        lib = _native.load()
        has_canvas = call.sid != _native.JR_DEPTH
        d_z = grads[0]
        d_c = grads[1] if has_canvas else None
        dev = call.tri_id.device
        zshape = (call.B, call.W, call.H)
        # The kernels overwrite d_zbuffer / d_canvas in place with the cotangents of the INCOMING buffers
        # (d out / d old = 1 - keep).  When nobody asks for those (the usual case: fresh buffers), the
        # incoming cotangents are passed read-only -- no clone, no zero-fill, no mask launch.
        zi, ci = _DIFF.index("zbuffer") + 1, _DIFF.index("canvas") + 1
        buffer_grads = bool(ctx.needs_input_grad[zi] or (has_canvas and ctx.needs_input_grad[ci]))
        if buffer_grads:
            d_z = torch.zeros(zshape, device=dev) if d_z is None else d_z.reshape(zshape).contiguous().clone()
            if has_canvas:
                d_c = (torch.zeros(zshape + (3,), device=dev) if d_c is None
                       else d_c.reshape(zshape + (3,)).contiguous().clone())
        else:
            d_z = None if d_z is None else d_z.reshape(zshape).contiguous()
            d_c = None if (d_c is None or not has_canvas) else d_c.reshape(zshape + (3,)).contiguous()
        g = JrGradArgs()
        g.no_buffer_grads = 0 if buffer_grads else 1
        g.d_zbuffer = d_z.data_ptr() if d_z is not None else None
        g.d_canvas = d_c.data_ptr() if (has_canvas and d_c is not None) else None
        outs: Dict[str, Tensor] = {}
        wanted = []
        for i, name in enumerate(_DIFF):
            if name in ("zbuffer", "canvas") or not ctx.present[i]:
                continue
            if ctx.needs_input_grad[i + 1]:
                wanted.append(name)
        if "colour" in wanted and "position" not in wanted:
            wanted.append("position")  # the kernel reduces both in one keyed pass
        for name in wanted:
            t = call.arrays[name]
            buf = torch.zeros_like(t)
            outs[name] = buf
            setattr(g, _GRAD_FIELD[name], JrF32(buf.data_ptr(), call.stride(name)))
        args = call.fill(call.tri_id, None, call.tri_id)  # zbuffer/canvas slots are unused by backward
        need = lib.jr_backward_workspace_bytes(C.byref(args), C.byref(g))
        ws = None
        if need:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            g.workspace, g.workspace_bytes = ws.data_ptr(), need
        with torch.cuda.device(dev):
            _native.check(lib.jr_render_backward(C.byref(args), C.byref(g), _native.stream_ptr(dev)))
        result: List[Optional[Tensor]] = [None]
        for i, name in enumerate(_DIFF):
            if not ctx.present[i] or not ctx.needs_input_grad[i + 1]:
                result.append(None)
            elif name == "zbuffer":
                result.append(d_z)
            elif name == "canvas":
                result.append(d_c)
            else:
                result.append(outs[name])
        return tuple(result)


def _render_arrays(sid: int, arrays: Dict[str, Any], zbuffer: Any, canvas: Optional[Any],
                   inplace: bool = False, return_tri_id: bool = False,
                   depth_epilogue: Optional[Tuple[float, Optional[float]]] = None,
                   display_u8: Optional[Sequence[float]] = None):
    """Internal entry: flat C-ABI array names -> (zbuffer, canvas, tri_id).

    ``display_u8`` (a background colour): the shading kernels write the display image -- ``(B?, H, W, 3)`` uint8,
    clamped, transposed, flipped (``utils.py:79-98``) -- instead of the fp32 canvas, which is returned in its
    place; ``canvas`` may then be ``None`` (pixels nothing covers show the background colour) or the incoming
    canvas (read only).  Forward only."""
    _native.load()  # fail loudly before anything else when the extension is missing
    arrays = dict(arrays)
    if _is_vmapped(zbuffer, canvas, *arrays.values()):
        if display_u8 is not None:
            raise NotImplementedError("display_uint8 inside torch.func.vmap")
        return _render_arrays_vmapped(sid, arrays, zbuffer, canvas, return_tri_id, depth_epilogue)
    texture_offset = int(arrays.pop("texture_offset", 0))
    # ---- validate shapes first (host side, before any transfer or launch)
    raw: Dict[str, Tensor] = {}
    for name, v in list(arrays.items()) + [("zbuffer", zbuffer), ("canvas", canvas)]:
        if v is None:
            continue
        raw[name] = v if isinstance(v, (torch.Tensor, InstancedArray)) else torch.as_tensor(
            v, dtype=torch.float32 if _SPEC[name][1] else torch.int32)
    # Instanced geometry (merge_objects on CUDA) stays factored only where the kernels can instance it themselves:
    # forward, no gradient anywhere, position AND normal from the same merge, not the Darboux shader.
    inst = [n for n in ("position", "normal") if isinstance(raw.get(n), InstancedArray)]
    if inst:
        same = len({id(raw[n].geom) for n in inst}) == 1 and isinstance(raw["position"], InstancedArray) and (
            "normal" not in raw or isinstance(raw["normal"], InstancedArray))
        grads = torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in raw.values())
        on_dev = isinstance(zbuffer, torch.Tensor) and zbuffer.is_cuda and zbuffer.device == raw[inst[0]].device
        if not same or grads or not on_dev or sid == _native.JR_PHONG_DARBOUX:
            for n in inst:
                raw[n] = raw[n].materialise()
    B = None
    batched: Dict[str, bool] = {}
    for name, t in raw.items():
        r = _SPEC[name][0]
        if t.ndim == r:
            batched[name] = False
        elif t.ndim == r + 1:
            batched[name] = True
            if B is None:
                B = t.shape[0]
            elif B != t.shape[0]:
                raise ValueError(f"inconsistent batch size for `{name}`: {t.shape[0]} vs {B}")
        else:
            raise ValueError(f"`{name}` has rank {t.ndim}, expected {r} or {r + 1}")
        trail = _TRAIL.get(name)
        if trail is not None:
            got = tuple(t.shape[t.ndim - len(trail):])
            if any(w is not None and w != g for w, g in zip(trail, got)):
                raise ValueError(f"`{name}` must end in shape {trail}, got {tuple(t.shape)}")
    if "colour" in raw and raw["colour"].shape[-2] != raw["position"].shape[-2]:
        raise ValueError("`colour` must have one row per `position` row")
    if "id_to_face" in raw and raw["id_to_face"].shape[-1] != raw["position"].shape[-2]:
        raise ValueError("`id_to_face` must have one entry per `position` row")
    if "normal_map" in raw and raw["normal_map"].shape[-3:-1] != raw["texture"].shape[-3:-1]:
        raise ValueError("`normal_map` must have the texture's width and height")
    W, H = raw["zbuffer"].shape[-2], raw["zbuffer"].shape[-1]
    if canvas is not None and tuple(raw["canvas"].shape[-3:]) != (W, H, 3):
        raise ValueError(f"canvas must be (..., {W}, {H}, 3), got {tuple(raw['canvas'].shape)}")
    # ---- device placement
    host_in = not (isinstance(zbuffer, torch.Tensor) and zbuffer.is_cuda)
    if host_in:
        if not torch.cuda.is_available():
            raise RuntimeError("jaxrenderer_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
    else:
        dev = zbuffer.device
    tens: Dict[str, Tensor] = {n: (t if isinstance(t, InstancedArray) else _as(t, _SPEC[n][1], dev))
                               for n, t in raw.items() if n not in ("zbuffer", "canvas")}
    z = _as(raw["zbuffer"], True, dev, written=True)
    c = _as(raw["canvas"], True, dev, written=True) if canvas is not None else None
    squeeze = B is None
    B = 1 if B is None else B
    # buffers always get a batch axis (vmap would broadcast them on output)
    if not batched["zbuffer"]:
        z = z.unsqueeze(0).expand(B, W, H)
    if c is not None and not batched["canvas"]:
        c = c.unsqueeze(0).expand(B, W, H, 3)
    call = _Call(sid, tens, B, W, H, texture_offset, batched)
    call.depth_epilogue = depth_epilogue     # (offset, fill value or None): JrRenderArgs.depth_*, depth shader only
    if display_u8 is not None:
        if needs_grad_any(tens, z, c) or sid in (_native.JR_DEPTH, _native.JR_PHONG_DARBOUX) or host_in:
            raise NotImplementedError("display_uint8 needs CUDA buffers, a canvas shader other than the Darboux "
                                      "shader, and no gradient")
        call.u8_background = tuple(float(v) for v in display_u8)
        call.u8_out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
    needs_grad = torch.is_grad_enabled() and any(
        t.requires_grad for t in list(tens.values()) + [z] + ([c] if c is not None else []))
    if needs_grad:
        diff = []
        for name in _DIFF:
            if name == "zbuffer":
                diff.append(z.contiguous())
            elif name == "canvas":
                diff.append(c.contiguous() if c is not None else None)
            else:
                diff.append(tens.get(name))
        out = _RenderFn.apply(call, *diff)
        if c is None:
            z_out, tri = out
            c_out = None
        else:
            z_out, c_out, tri = out
    else:
        def _own(t: Tensor, src: Any) -> Tensor:
            tc = t.contiguous()
            fresh = tc.data_ptr() != t.data_ptr() or host_in or not (
                isinstance(src, torch.Tensor) and src.data_ptr() == tc.data_ptr())
            return tc if (fresh or inplace) else tc.clone()

        z_out = _own(z, zbuffer)
        if call.u8_out is not None:
            c_in = c.contiguous() if c is not None else None      # read only: the background
            tri = _forward_native(call, z_out, c_in, need_tri=return_tri_id)
            c_out = call.u8_out
        else:
            c_out = _own(c, canvas) if c is not None else None
            tri = _forward_native(call, z_out, c_out, need_tri=return_tri_id)
    if squeeze:
        z_out = z_out[0]
        c_out = c_out[0] if c_out is not None else None
        tri = tri[0] if tri is not None else None
    if host_in:
        z_out = z_out.cpu()
        c_out = c_out.cpu() if c_out is not None else None
        tri = tri.cpu() if (return_tri_id and tri is not None) else tri
    return z_out, c_out, tri


def needs_grad_any(tens: Dict[str, Any], z: Tensor, c: Optional[Tensor]) -> bool:
    return torch.is_grad_enabled() and any(
        getattr(t, "requires_grad", False) for t in list(tens.values()) + [z] + ([c] if c is not None else []))


def _render_arrays_vmapped(sid: int, arrays: Dict[str, Any], zbuffer: Any, canvas: Optional[Any],
                           return_tri_id: bool, depth_epilogue: Any):
    """``torch.func.vmap`` support (the reference's batching idiom: ``jax.vmap(lambda m, b: Renderer.render(...))``,
    ``examples/batch_rendering.py:87-95``).  Inside a vmap every mapped input is a functorch BatchedTensor; the
    kernels batch NATIVELY, so the mapped axis of each input is moved to the front and handed to the ordinary
    (batched) call -- mapped inputs become batched arrays, un-mapped ones stay shared (``in_axes=None`` = batch
    stride 0) -- and the outputs are re-wrapped for the enclosing vmap.  One vmap level; gradients taken OUTSIDE the
    vmap flow through the underlying tensors as usual."""
    f = torch._C._functorch
    level = None

    def unwrap(name: str, v: Any) -> Any:
        nonlocal level
        if not (isinstance(v, torch.Tensor) and f.is_batchedtensor(v)):
            if isinstance(v, torch.Tensor) and v.ndim == _SPEC[name][0] + 1:
                raise NotImplementedError(f"`{name}` carries its own batch axis inside torch.func.vmap: nested "
                                          "batching is not supported (map that axis with the vmap instead)")
            return v
        lvl, bdim = f.maybe_get_level(v), f.maybe_get_bdim(v)
        raw = f.get_unwrapped(v)
        if f.is_batchedtensor(raw):
            raise NotImplementedError("nested torch.func.vmap around the renderer is not supported")
        if level not in (None, lvl):
            raise NotImplementedError("inputs mapped by different vmap levels")
        level = lvl
        if raw.ndim - 1 != _SPEC[name][0]:
            raise ValueError(f"`{name}` has rank {raw.ndim - 1} inside vmap, expected {_SPEC[name][0]}")
        return raw.movedim(bdim, 0)

    raw_arrays = {k: (unwrap(k, v) if k in _SPEC else v) for k, v in arrays.items()}
    z = unwrap("zbuffer", zbuffer)
    c = unwrap("canvas", canvas) if canvas is not None else None
    z_out, c_out, tri = _render_arrays(sid, raw_arrays, z, c, inplace=False, return_tri_id=return_tri_id,
                                       depth_epilogue=depth_epilogue)
    wrap = lambda t: None if t is None else f._add_batch_dim(t, 0, level)  # noqa: E731
    return wrap(z_out), wrap(c_out), wrap(tri)


def render(camera: Camera, shader: type, buffers: Buffers, face_indices: Any, extra: Any,
           loop_unroll: int = 1, *, inplace: bool = False, return_tri_id: bool = False):
    """Render a scene with a built-in shader (reference ``pipeline.py:470-537``).

    ``inplace=True`` is the reference's buffer donation (``pipeline.py:466``):
    the given CUDA buffers are updated in place and returned.
    ``return_tri_id=True`` additionally returns the triangle-id G-buffer.
    """
    del loop_unroll
    if not (isinstance(shader, type) and shader in BUILTIN_SHADERS):
        name = getattr(shader, "__name__", repr(shader))
        raise UnsupportedShaderError(
            f"shader `{name}` is not one of the built-in shaders "
            f"({', '.join(s.__name__ for s in BUILTIN_SHADERS)}). Custom `Shader` subclasses are "
            "not supported by jaxrenderer_b200: the five stages are fused into CUDA kernels and "
            "there is no Python fallback.")
    if len(extra) == 0:
        raise ValueError("`extra` must have the per-vertex array as its first field (pipeline.py:488-491)")
    sid = shader._jr_shader
    targets = tuple(buffers.targets)
    if sid == _native.JR_DEPTH:
        if len(targets) != 0:
            raise ValueError("DepthShader renders to the z-buffer only: buffers.targets must be ()")
        canvas = None
    else:
        if len(targets) != 1:
            raise ValueError(f"{shader.__name__} renders to exactly one target (canvas)")
        canvas = targets[0]
    try:
        arrays = _collect(shader, camera, face_indices, extra)
    except AttributeError as e:
        raise TypeError(f"`extra` / `camera` do not carry the fields {shader.__name__} reads: {e}") from None
    z, c, tri = _render_arrays(sid, arrays, buffers.zbuffer, canvas, inplace=inplace,
                               return_tri_id=return_tri_id)
    out = Buffers(zbuffer=z, targets=() if c is None else (c,))
    return (out, tri) if return_tri_id else out


__all__ = ["render", "Shader"]
