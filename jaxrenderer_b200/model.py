"""Mesh containers and object merging (host side), mirroring ``renderer/model.py``.

``Model`` (:37-143), ``MergedModel`` (:155-339), ``ModelObject`` (:342-421),
``batch_models`` (:425-442), ``merge_objects`` (:447-555).  ``uv_repeat``
(:306-339) additionally lives inside the phong_reflection* fragment kernels.

Batch convention: any tensor leaf may carry one extra leading batch axis (the
reference gets this from ``jax.vmap``); ``merge_objects`` broadcasts them.
"""
from __future__ import annotations

import collections

from typing import Any, List, NamedTuple, Optional, Sequence, Tuple

import torch

from .geometry import to_cartesian, to_homogeneous, transform_matrix_from_rotation
from .types import Tensor, _f32, _i32, _is_vmapped


class Model(NamedTuple):
    """``model.py:37-53``."""

    verts: Tensor
    norms: Tensor
    uvs: Tensor
    faces: Tensor
    faces_norm: Tensor
    faces_uv: Tensor
    diffuse_map: Tensor
    specular_map: Tensor

    @classmethod
    def create(cls, verts: Tensor, norms: Tensor, uvs: Tensor, faces: Tensor,
               diffuse_map: Tensor, specular_map: Optional[Tensor] = None) -> "Model":
        """``model.py:55-90``: same index buffer for all attributes; default
        specular map is a constant 2.0."""
        diffuse_map = _f32(diffuse_map)
        if specular_map is None:
            specular_map = torch.full(diffuse_map.shape[:2], 2.0, dtype=torch.float32,
                                      device=diffuse_map.device)
        faces = _i32(faces)
        return cls(verts=_f32(verts), norms=_f32(norms), uvs=_f32(uvs), faces=faces,
                   faces_norm=faces, faces_uv=faces, diffuse_map=diffuse_map,
                   specular_map=_f32(specular_map))

    def value_checks(self) -> None:
        """Index-range validation (``model.py:105-143``); raises ``ValueError``."""
        for name, idx, n in (("faces", self.faces, self.verts.shape[-2]),
                             ("faces_norm", self.faces_norm, self.norms.shape[-2]),
                             ("faces_uv", self.faces_uv, self.uvs.shape[-2])):
            lo, hi = int(idx.min()), int(idx.max())
            if lo < 0 or hi >= n:
                raise ValueError(f"{name} out of bound, expected [0, {n}), got [{lo}, {hi}].")


class MergedModel(NamedTuple):
    """``model.py:155-178``."""

    verts: Tensor
    norms: Tensor
    uvs: Tensor
    faces: Tensor
    faces_norm: Tensor
    faces_uv: Tensor
    texture_index: Tensor
    double_sided: Tensor
    texture_shape: Tensor
    offset: int
    diffuse_map: Tensor
    specular_map: Tensor

    @staticmethod
    def generate_object_vert_info(counts: Sequence[int], values: Sequence[Any]) -> Tensor:
        """``model.py:180-213``: repeat one value per object over its vertices."""
        parts = []
        for count, value in zip(counts, values):
            v = torch.as_tensor(value)
            parts.append(v.expand(count, *v.shape).clone() if v.ndim else v.repeat(count))
        return torch.cat(parts, dim=0)

    @staticmethod
    def merge_verts(vs: Sequence[Tensor], fs: Sequence[Tensor]) -> Tuple[Tensor, Tensor]:
        """``model.py:215-247``: concatenate attribute arrays, offset indices."""
        cumsum = [0]
        for v in vs[:-1]:
            cumsum.append(cumsum[-1] + v.shape[-2])
        nb = max(v.ndim for v in vs)
        batch = torch.broadcast_shapes(*[v.shape[:-2] for v in vs])
        verts = torch.cat([v.expand(*batch, *v.shape[-2:]) if nb > 2 else v for v in vs], dim=-2)
        fb = torch.broadcast_shapes(*[f.shape[:-2] for f in fs])
        faces = torch.cat([(f + cumsum[i]).expand(*fb, *f.shape[-2:]) for i, f in enumerate(fs)],
                          dim=-2)
        return verts, faces.to(torch.int32)

    @staticmethod
    def merge_maps(maps: Sequence[Tensor], rank: Optional[int] = None) -> Tuple[Tensor, Tuple[int, int]]:
        """``model.py:249-301``: zero-pad every map to the max shape and
        concatenate along the first (width) axis.  ``base`` rank = rank of an
        un-batched map (2 for specular, 3 for diffuse)."""
        if rank is None:
            rank = min(m.ndim for m in maps)
        # a map is batched iff it has rank+1 dims; all maps share `rank`
        single = tuple(max(m.shape[m.ndim - rank + i] for m in maps) for i in range(rank))
        batch = torch.broadcast_shapes(*[m.shape[: m.ndim - rank] for m in maps])
        padded = []
        for m in maps:
            pad: List[int] = []
            for i in reversed(range(rank)):
                pad += [0, single[i] - m.shape[m.ndim - rank + i]]
            p = torch.nn.functional.pad(_f32(m), pad)
            padded.append(p.expand(*batch, *p.shape[p.ndim - rank:]))
        return torch.cat(padded, dim=len(batch)), (single[0], single[1])

    @staticmethod
    def uv_repeat(uv: Tensor, shape: Tensor, map_index: Tensor, offset: Any) -> Tensor:
        """``model.py:306-339`` (host twin of the in-kernel version)."""
        frac = uv - torch.trunc(uv)
        frac = torch.where(frac < 0, frac + 1, frac)
        out = frac * shape
        out[..., 0] = out[..., 0] + map_index * offset
        return out


class ModelObject(NamedTuple):
    """``model.py:342-421``."""

    model: Model
    local_scaling: Any = (1.0, 1.0, 1.0)
    transform: Any = ((1.0, 0.0, 0.0, 0.0), (0.0, 1.0, 0.0, 0.0), (0.0, 0.0, 1.0, 0.0), (0.0, 0.0, 0.0, 1.0))
    double_sided: Any = False

    def replace_with_position(self, position: Tensor) -> "ModelObject":
        t = _f32(self.transform).clone()
        t[..., :3, 3] = _f32(position)
        return self._replace(transform=t)

    def replace_with_orientation(self, orientation: Optional[Tensor] = None,
                                 rotation_matrix: Optional[Tensor] = None) -> "ModelObject":
        if rotation_matrix is None:
            if orientation is None:
                orientation = torch.tensor((0.0, 0.0, 0.0, 1.0))
            rotation_matrix = transform_matrix_from_rotation(_f32(orientation))
        t = _f32(self.transform).clone()
        t[..., :3, :3] = rotation_matrix
        return self._replace(transform=t)

    def replace_with_local_scaling(self, local_scaling: Tensor) -> "ModelObject":
        return self._replace(local_scaling=local_scaling)

    def replace_with_double_sided(self, double_sided: Any) -> "ModelObject":
        return self._replace(double_sided=double_sided)


def batch_models(models: Sequence[MergedModel]) -> MergedModel:
    """Stack several ``MergedModel`` along a new leading axis (``model.py:425-442``).
    ``offset`` (a Python int) must agree and stays un-batched."""
    fields = []
    for i, name in enumerate(MergedModel._fields):
        if name == "offset":
            assert all(m.offset == models[0].offset for m in models)
            fields.append(models[0].offset)
        else:
            # (a CUDA merge_objects hands out un-materialised verts / norms: torch.as_tensor would walk them as sequences)
            vals = [m[i].materialise() if isinstance(m[i], InstancedArray) else torch.as_tensor(m[i]) for m in models]
            fields.append(torch.stack(vals, dim=0))
    return MergedModel._make(fields)


def _frob_normalise(x: Tensor) -> Tensor:
    """``normalise`` applied by the reference to a whole ``(N, 3)`` array."""
    return x / torch.linalg.norm(x, dim=(-2, -1), keepdim=True)


_STATIC_FIELDS = ("verts", "norms", "uvs", "faces", "faces_norm", "faces_uv", "diffuse_map", "specular_map")
_STATIC_RANK = (2, 2, 2, 2, 2, 2, 3, 2)
_STATIC_CACHE: "collections.OrderedDict[tuple, tuple]" = collections.OrderedDict()
_STATIC_CACHE_SIZE = 8


def _merge_static(models: Sequence[Model], dev: torch.device) -> dict:
    """The part of ``merge_objects`` that depends only on the objects' meshes and maps, not on their
    per-frame transforms: concatenated index buffers / UVs / local vertices, the texture atlas and the
    per-vertex object tables (``model.py:447-555``).  The reference gets this for free from ``jax.jit``
    (constant folding at trace time); here the result is memoised on the identity + in-place version of
    the model tensors (entries keep the tensors alive, so an id cannot be recycled).  Batched or
    gradient-carrying mesh tensors bypass the cache."""
    tensors = [getattr(m, f) for m in models for f in _STATIC_FIELDS]
    cacheable = all(isinstance(t, torch.Tensor) and not t.requires_grad and t.ndim == r
                    for t, r in zip(tensors, _STATIC_RANK * len(models))) and not _is_vmapped(*tensors)
    key = (dev, tuple((id(t), t._version) for t in tensors)) if cacheable else None
    if key is not None:
        hit = _STATIC_CACHE.get(key)
        if hit is not None:
            _STATIC_CACHE.move_to_end(key)
            return hit[1]
    n_obj = len(models)
    counts_v = [m.verts.shape[-2] for m in models]
    counts_n = [m.norms.shape[-2] for m in models]
    counts_uv = [m.uvs.shape[-2] for m in models]

    def starts(counts):
        out = [0]
        for c in counts:
            out.append(out[-1] + c)
        return out

    def cat_faces(fs, counts):
        off = starts(counts)
        fb = torch.broadcast_shapes(*[f.shape[:-2] for f in fs])
        return torch.cat([(f.to(dev) + off[i]).expand(*fb, *f.shape[-2:]) for i, f in enumerate(fs)],
                         dim=-2).to(torch.int32)

    def cat_attr(parts):
        parts = [_f32(p, dev) for p in parts]
        batch = torch.broadcast_shapes(*[p.shape[:-2] for p in parts])
        return torch.cat([p.expand(*batch, *p.shape[-2:]) for p in parts], dim=-2).contiguous()

    diffuse_map, single = MergedModel.merge_maps([m.diffuse_map for m in models], rank=3)
    st = {
        "counts": counts_v,
        "vert_object": torch.repeat_interleave(torch.arange(n_obj, dtype=torch.int32),
                                               torch.tensor(counts_v)).to(dev),
        "norm_start": torch.tensor(starts(counts_n), dtype=torch.int32).to(dev),
        "norm_object": torch.repeat_interleave(torch.arange(n_obj, dtype=torch.int32),
                                               torch.tensor(counts_n)).to(dev),
        "local_verts": cat_attr([m.verts for m in models]),
        "local_norms": cat_attr([m.norms for m in models]),
        "uvs": cat_attr([m.uvs for m in models]),
        "faces": cat_faces([m.faces for m in models], counts_v),
        "faces_norm": cat_faces([m.faces_norm for m in models], counts_n),
        "faces_uv": cat_faces([m.faces_uv for m in models], counts_uv),
        "texture_shape": torch.tensor([tuple(m.diffuse_map.shape[-3:-1]) for m in models],
                                      dtype=torch.int32).to(dev),
        "diffuse_map": diffuse_map.to(dev), "offset": int(single[0]),
        "specular_map": MergedModel.merge_maps([m.specular_map for m in models], rank=2)[0].to(dev),
    }
    if key is not None:
        _STATIC_CACHE[key] = (tensors, st)
        while len(_STATIC_CACHE) > _STATIC_CACHE_SIZE:
            _STATIC_CACHE.popitem(last=False)
    return st


class InstancedGeometry:
    """The geometry half of ``merge_objects`` kept in FACTORED form: shared local meshes
    (``st``, memoised) + per-image object ``scaling`` / ``transform`` / ``normal_matrix``.  The render kernels
    evaluate the world-space merge on the fly from these (``JrRenderArgs.inst_*``), so a batch of B images moves
    ``B * n_objects * 35`` floats instead of ``B * (Nv + Nn) * 3``; ``materialise()`` runs the fused
    ``jr_merge_objects`` kernels (the same arithmetic, bit-identical results) for everything else."""

    def __init__(self, objects: Sequence["ModelObject"], dev: torch.device, st: dict):
        self.st, self.dev = st, dev
        self.n_obj = len(objects)

        def stack(vals, base_rank, shape):
            ts = [_f32(v, dev) for v in vals]
            if any(t.ndim > base_rank for t in ts):
                batch = torch.broadcast_shapes(*[t.shape[: t.ndim - base_rank] for t in ts])
                ts = [t.expand(*batch, *shape) for t in ts]
            return torch.stack(ts, dim=-(base_rank + 1)).contiguous()

        self.scaling = stack([o.local_scaling for o in objects], 1, (3,))
        self.transform = stack([o.transform for o in objects], 2, (4, 4))
        self._nmat: Optional[Tensor] = None   # inverse-transpose of the transforms: only normals need it
        lv, ln = st["local_verts"], st["local_norms"]
        B = None
        for t, r in ((lv, 2), (ln, 2), (self.scaling, 2), (self.transform, 3)):
            if t.ndim == r + 1:
                B = t.shape[0]
        self.batch = B                      # None = un-batched
        self._verts: Optional[Tensor] = None
        self._norms: Optional[Tensor] = None
        self._scales: Optional[Tensor] = None

    @property
    def nmat(self) -> Tensor:
        if self._nmat is None:
            # inv_ex: no host synchronisation on the singularity check (a singular transform gives inf / nan
            # normals, as in the reference)
            self._nmat = torch.linalg.inv_ex(self.transform, check_errors=False).inverse.transpose(-1, -2).contiguous()
        return self._nmat

    # ---- C-ABI plumbing
    def _merge_args(self):
        from ._native import JrF32, JrI32, JrMergeArgs

        st = self.st
        lv, ln = st["local_verts"], st["local_norms"]

        def f32(t, r):
            return JrF32(t.data_ptr(), int(t[0].numel()) if t.ndim == r + 1 else 0)

        a = JrMergeArgs()
        a.B, a.n_objects = (1 if self.batch is None else self.batch), self.n_obj
        a.n_verts, a.n_norms = lv.shape[-2], ln.shape[-2]
        a.local_verts, a.local_norms = f32(lv, 2), f32(ln, 2)
        a.vert_object = JrI32(st["vert_object"].data_ptr(), 0)
        a.norm_start = JrI32(st["norm_start"].data_ptr(), 0)
        a.scaling, a.transform, a.normal_matrix = f32(self.scaling, 2), f32(self.transform, 3), f32(self.nmat, 3)
        return a

    def materialise(self) -> Tuple[Tensor, Tensor]:
        """World-space ``(verts, norms)``: two launches of ``jr_merge_objects`` (cached)."""
        if self._verts is None:
            import ctypes as C

            from . import _native

            a = self._merge_args()
            out_v = torch.empty((a.B, a.n_verts, 3), dtype=torch.float32, device=self.dev)
            out_n = torch.empty((a.B, a.n_norms, 3), dtype=torch.float32, device=self.dev)
            a.out_verts, a.out_norms = out_v.data_ptr(), out_n.data_ptr()
            lib = _native.load()
            with torch.cuda.device(self.dev):
                _native.check(lib.jr_merge_objects(C.byref(a), _native.stream_ptr(self.dev)))
            self._verts, self._norms = (out_v[0], out_n[0]) if self.batch is None else (out_v, out_n)
        return self._verts, self._norms

    def norm_scales(self) -> Tensor:
        """``(B?, n_objects, 2)``: the two Frobenius normalisation constants of the normal transform
        (``model.py:517-530``) per (image, object): one launch of ``jr_instance_norm_scales`` (cached)."""
        if self._scales is None:
            import ctypes as C

            from . import _native

            a = self._merge_args()
            out = torch.empty((a.B, self.n_obj, 2), dtype=torch.float32, device=self.dev)
            lib = _native.load()
            with torch.cuda.device(self.dev):
                _native.check(lib.jr_instance_norm_scales(C.byref(a), out.data_ptr(), _native.stream_ptr(self.dev)))
            self._scales = out[0] if self.batch is None else out
        return self._scales


class InstancedArray:
    """``MergedModel.verts`` / ``.norms`` of a CUDA ``merge_objects`` call, not yet materialised.  Passed on to
    ``Renderer.render`` / ``pipeline.render`` / the shadow pass it makes the kernels instance the geometry
    themselves; touched in any other way (indexing, arithmetic, ``.cpu()``, any ``torch.*`` function) it turns
    into the ordinary world-space tensor, computed once by the fused merge kernels."""

    def __init__(self, kind: str, geom: InstancedGeometry):
        assert kind in ("verts", "norms")
        self.kind, self.geom = kind, geom

    # ---- what the host code asks without needing values
    @property
    def shape(self) -> torch.Size:
        n = self.geom.st["local_verts" if self.kind == "verts" else "local_norms"].shape[-2]
        return torch.Size(((self.geom.batch,) if self.geom.batch is not None else ()) + (n, 3))

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def dim(self) -> int:
        return self.ndim

    @property
    def device(self) -> torch.device:
        return self.geom.dev

    dtype = torch.float32
    requires_grad = False
    is_cuda = True

    def detach(self) -> "InstancedArray":
        return self

    # ---- everything else: the materialised tensor
    def materialise(self) -> Tensor:
        v, n = self.geom.materialise()
        return v if self.kind == "verts" else n

    def __getattr__(self, name: str):       # .cpu(), .to(), .abs(), ... (only reached for names not defined above)
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.materialise(), name)

    def __getitem__(self, idx):
        return self.materialise()[idx]

    def __len__(self) -> int:
        return self.shape[0]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        conv = lambda x: x.materialise() if isinstance(x, InstancedArray) else x  # noqa: E731
        args = tuple(conv(a) if not isinstance(a, (list, tuple)) else type(a)(conv(b) for b in a) for a in args)
        kwargs = {k: conv(v) for k, v in (kwargs or {}).items()}
        return func(*args, **kwargs)

    def _binary(name):  # noqa: N805
        def op(self, other):
            other = other.materialise() if isinstance(other, InstancedArray) else other
            return getattr(self.materialise(), name)(other)
        return op

    __add__, __radd__, __sub__, __rsub__ = _binary("__add__"), _binary("__radd__"), _binary("__sub__"), _binary("__rsub__")
    __mul__, __rmul__, __truediv__, __matmul__ = _binary("__mul__"), _binary("__rmul__"), _binary("__truediv__"), _binary("__matmul__")
    __eq__, __ne__ = _binary("__eq__"), _binary("__ne__")
    __hash__ = object.__hash__
    del _binary

    def __neg__(self):
        return -self.materialise()

    def __repr__(self) -> str:
        return f"InstancedArray({self.kind}, shape={tuple(self.shape)}, device={self.device})"


def _merge_geometry_fused(objects: Sequence[ModelObject], dev: torch.device, st: dict):
    """CUDA fast path of the vertex / normal part of ``merge_objects``: two launches of
    ``jr_merge_objects`` (``csrc/jr_forward.cu``) instead of ~15 framework ops per object."""
    return InstancedGeometry(objects, dev, st).materialise()


_DOUBLE_SIDED_CACHE: dict = {}


def _double_sided_per_vertex(objects: Sequence[ModelObject], st: dict, dev: torch.device) -> Tensor:
    """One flag per merged vertex (``model.py:492-497``), without reading device values on the host."""
    flags = [o.double_sided for o in objects]
    if any(isinstance(f, torch.Tensor) and f.is_cuda for f in flags):
        if all(isinstance(f, torch.Tensor) and f.shape == flags[0].shape and f.device == flags[0].device
               for f in flags):
            per_obj = torch.stack(flags).reshape(len(flags), -1)[:, 0].to(device=dev, dtype=torch.bool)
        else:
            per_obj = torch.stack([torch.as_tensor(f, device=dev).reshape(-1)[0].to(torch.bool) for f in flags])
        return per_obj[st["vert_object"].long()]
    # host flags (Python bools / host tensors): memoised by value on the device (no per-call host->device copy,
    # which would also break CUDA-graph capture of the facade)
    vals = tuple(bool(torch.as_tensor(f).reshape(-1)[0]) for f in flags)
    key = (str(dev), tuple(st["counts"]), vals)
    hit = _DOUBLE_SIDED_CACHE.get(key)
    if hit is None:
        hit = MergedModel.generate_object_vert_info(st["counts"], list(vals)).to(dev)
        if len(_DOUBLE_SIDED_CACHE) > 64:
            _DOUBLE_SIDED_CACHE.clear()
        _DOUBLE_SIDED_CACHE[key] = hit
    return hit


def merge_objects(objects: Sequence[ModelObject]) -> MergedModel:
    """World-space merge of all objects into one mesh + texture atlas
    (``model.py:447-555``).  On CUDA inputs (and when no gradient is requested through the
    transforms) the vertex / normal transforms run in the fused ``jr_merge_objects`` kernels; the
    transform-independent part is memoised (``_merge_static``)."""
    models = [obj.model for obj in objects]
    dev = models[0].verts.device
    st = _merge_static(models, dev)
    double_sided = _double_sided_per_vertex(objects, st, dev)

    def transform_vert(verts: Tensor, local_scaling: Any, transform: Any) -> Tensor:
        ls = _f32(local_scaling, dev)
        tf = _f32(transform, dev)
        scaled = _f32(verts) * ls[..., None, :]
        return to_cartesian(to_homogeneous(scaled) @ tf.transpose(-1, -2))

    def transform_normals(normals: Tensor, transform: Any) -> Tensor:
        tf = _f32(transform, dev)
        m = torch.linalg.inv(tf).transpose(-1, -2)
        n = _frob_normalise(_f32(normals))
        t = (to_homogeneous(n, 0.0) @ m.transpose(-1, -2))[..., :3]
        return _frob_normalise(t)

    needs_grad = torch.is_grad_enabled() and any(
        isinstance(t, torch.Tensor) and t.requires_grad
        for o in objects for t in (o.model.verts, o.model.norms, o.local_scaling, o.transform))
    vmapped = _is_vmapped(*[t for o in objects for t in (o.model.verts, o.model.norms, o.local_scaling, o.transform)])
    if dev.type == "cuda" and not needs_grad and not vmapped:
        # factored form: the render kernels instance the geometry themselves; any other use of .verts / .norms
        # materialises them with the fused merge kernels (InstancedArray)
        geom = InstancedGeometry(objects, dev, st)
        verts, norms = InstancedArray("verts", geom), InstancedArray("norms", geom)
    else:
        verts, _ = MergedModel.merge_verts(
            [transform_vert(o.model.verts, o.local_scaling, o.transform) for o in objects],
            [m.faces for m in models])
        norms, _ = MergedModel.merge_verts(
            [transform_normals(o.model.norms, o.transform) for o in objects],
            [m.faces_norm for m in models])
    return MergedModel(
        verts=verts, norms=norms, uvs=st["uvs"], faces=st["faces"], faces_norm=st["faces_norm"],
        faces_uv=st["faces_uv"], texture_shape=st["texture_shape"], texture_index=st["vert_object"],
        double_sided=double_sided, offset=st["offset"], diffuse_map=st["diffuse_map"],
        specular_map=st["specular_map"])
