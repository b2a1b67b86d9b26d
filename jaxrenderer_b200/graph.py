"""`graphed(fn)`: capture a rendering function into a CUDA graph once, replay it per call.

The reference's users wrap their render function in `jax.jit` (`examples/batch_rendering.py:87-95`,
`notebooks/32x32/A100.ipynb:341`): one trace, then one executable launch per call.  The counterpart here is a CUDA
graph: after warm-up a call of `Renderer.get_camera_image` / `pipeline.render` neither allocates outside the caching
allocator nor synchronises, so the whole call sequence (camera construction, instancing constants, shadow pass,
visibility, shading, display epilogue) captures into ONE graph and a replay costs one launch from the host instead of
seven plus ~0.4 ms of Python (1.21 ms -> 0.91 ms per 1024 Brax frames, 0.60 -> 0.11 ms per 640x480 frame).

    def frame(transforms, eye):                                   # closes over the meshes, the light, the camera
        objs = [o._replace(transform=t) for o, t in zip(objects, transforms)]
        return jr.Renderer.get_camera_image(objs, light, cam._replace(position=eye), 84, 84, shadow_param=sp)
    render = jr.graphed(frame)
    img = render(transforms, eye)     # first call with these shapes: warm-up + capture
    img = render(transforms2, eye2)   # later calls: inputs copied into the captured buffers, one replay

Floating-point tensors anywhere inside the arguments (NamedTuples, lists, dicts: pytrees) are the graph's inputs;
everything else (ints, floats, strings, None, shader classes) is a static argument and part of the cache key, like
`static_argnums`; what the function closes over is a constant of the capture, as under `jax.jit`.  Index buffers and
other integer tensors must be closed over, not passed (rejected with an error): the facade memoises the merged form of
the meshes and maps per tensor identity (`model._merge_static`, the counterpart of jit's constant folding) on the HOST,
which a replay does not re-run -- topology is a constant of a captured call, and so are the mesh attributes and maps
of a `Model` (vertices, normals, uvs, textures) even when they are floating point: pass what changes from frame to
frame (object transforms and scalings, camera and light parameters, buffers), close over the rest.  One graph is kept
per (argument structure, shapes, dtypes, static values).  The returned tensors are
the graph's own output buffers: they are overwritten by the next call with the same signature -- clone what has to
outlive it (`copy_outputs=True` does so for you).  Forward only: a captured call records no autograd tape (tensors that
require grad are rejected; use the eager call to differentiate)."""
from __future__ import annotations

from typing import Any, Callable, Dict, Tuple

import torch
from torch.utils import _pytree as pytree


class _Entry:
    __slots__ = ("graph", "inputs", "outputs", "out_spec")


class Graphed:
    """Callable returned by :func:`graphed`."""

    def __init__(self, fn: Callable[..., Any], warmup: int = 2, copy_outputs: bool = False):
        self.fn = fn
        self.warmup = max(1, int(warmup))
        self.copy_outputs = bool(copy_outputs)
        self._cache: Dict[Tuple, _Entry] = {}

    @staticmethod
    def _split(args: tuple, kwargs: dict):
        leaves, spec = pytree.tree_flatten((args, kwargs))
        tensors = [(i, v) for i, v in enumerate(leaves) if isinstance(v, torch.Tensor)]
        static = tuple((i, v if isinstance(v, (int, float, bool, str, type(None), type)) else id(v))
                       for i, v in enumerate(leaves) if not isinstance(v, torch.Tensor))
        return leaves, spec, tensors, static

    def __call__(self, *args: Any, **kwargs: Any) -> Any:
        leaves, spec, tensors, static = self._split(args, kwargs)
        if not tensors:
            raise ValueError("graphed(fn): no tensor among the arguments -- nothing to replay on")
        dev = tensors[0][1].device
        for _, t in tensors:
            if t.device.type != "cuda":
                raise ValueError("graphed(fn) needs CUDA tensors (there is no CPU path); got a tensor on " + str(t.device))
            if t.device != dev:
                raise ValueError("graphed(fn): all tensors must live on one device")
            if t.requires_grad and torch.is_grad_enabled():
                raise ValueError("graphed(fn) is forward only: call the function eagerly to differentiate")
            if not t.dtype.is_floating_point:
                raise ValueError("graphed(fn): integer tensors (index buffers, texture indices) are constants of a captured "
                                 "call -- close over them instead of passing them (see the module docstring)")
        key = (str(spec), static, tuple((i, tuple(t.shape), t.dtype, tuple(t.stride())) for i, t in tensors), dev.index)
        entry = self._cache.get(key)
        if entry is None:
            entry = self._capture(leaves, spec, tensors, dev)
            self._cache[key] = entry
        else:
            for (_, src), dst in zip(tensors, entry.inputs):
                if src.data_ptr() != dst.data_ptr():
                    dst.copy_(src, non_blocking=True)
        entry.graph.replay()
        outs = [o.clone() for o in entry.outputs] if self.copy_outputs else list(entry.outputs)
        return pytree.tree_unflatten(outs, entry.out_spec)

    def _capture(self, leaves, spec, tensors, dev) -> _Entry:
        e = _Entry()
        with torch.no_grad():
            # the graph reads its inputs from buffers it owns: later calls copy into them
            e.inputs = [torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, device=dev).copy_(t) for _, t in tensors]
            own = list(leaves)
            for (i, _), buf in zip(tensors, e.inputs):
                own[i] = buf
            c_args, c_kwargs = pytree.tree_unflatten(own, spec)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(self.warmup):          # allocator pools, memoised constants, kernel attributes
                    self.fn(*c_args, **c_kwargs)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            e.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(e.graph):
                out = self.fn(*c_args, **c_kwargs)
        out_leaves, e.out_spec = pytree.tree_flatten(out)
        if not all(isinstance(o, torch.Tensor) for o in out_leaves):
            raise TypeError("graphed(fn): the function must return tensors (or pytrees of tensors)")
        e.outputs = out_leaves
        return e

    def cache_size(self) -> int:
        return len(self._cache)


def graphed(fn: Callable[..., Any] = None, *, warmup: int = 2, copy_outputs: bool = False):
    """Wrap `fn` (see the module docstring).  Usable as a decorator: ``@jr.graphed`` or ``@jr.graphed(copy_outputs=True)``."""
    if fn is None:
        return lambda f: Graphed(f, warmup=warmup, copy_outputs=copy_outputs)
    return Graphed(fn, warmup=warmup, copy_outputs=copy_outputs)
