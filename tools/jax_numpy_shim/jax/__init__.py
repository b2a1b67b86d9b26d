"""NumPy-backed stand-in for the slice of `jax` the reference renderer uses (see ../README.md)."""
from __future__ import annotations

import contextlib
import functools
import os as _os

import numpy as _np

from . import numpy  # noqa: F401  (jax.numpy)
from . import lax, tree_util, dtypes, stages, profiler  # noqa: F401
from . import experimental  # noqa: F401
from .numpy import _wrap, Array  # noqa: F401
from .tree_util import tree_map, tree_flatten, tree_unflatten

__version__ = "0.4.13-numpy-shim"


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def jit(f=None, **kwargs):
    """Eager: returns the function itself (static_argnames / donate_argnums are irrelevant)."""
    if f is None:
        return lambda g: jit(g, **kwargs)

    @functools.wraps(f)
    def wrapper(*a, **k):
        return f(*a, **k)

    wrapper.lower = lambda *a, **k: None
    return wrapper


@contextlib.contextmanager
def ensure_compile_time_eval():
    yield


@contextlib.contextmanager
def named_scope(name):
    yield


_VMAP_DEPTH = 0


def _forked_map(one, size, procs):
    """`[one(i) for i in range(size)]` over `procs` forked workers (JAX_SHIM_PROCS): the stand-in's `vmap` is a Python
    loop, and a render of the real Brax frame at 84x84 is hours of it.  Same calls, same order of the results, same
    arithmetic -- the workers are forks of this very process (closures and all), each takes every procs-th index and
    pickles its results back through a pipe."""
    import pickle

    global _VMAP_DEPTH
    pipes = []
    for w in range(procs):
        r, wfd = _os.pipe()
        pid = _os.fork()
        if pid == 0:                      # worker
            _os.close(r)
            code = 0
            try:
                _VMAP_DEPTH = 1           # nested vmaps inside a worker stay plain loops
                res = [(i, one(i)) for i in range(w, size, procs)]
                with _os.fdopen(wfd, "wb") as fh:
                    pickle.dump(res, fh, protocol=pickle.HIGHEST_PROTOCOL)
            except BaseException:         # noqa: BLE001  (report and die: never return into the parent's stack)
                import traceback
                traceback.print_exc()
                code = 1
            _os._exit(code)
        _os.close(wfd)
        pipes.append((pid, r))
    outs = [None] * size
    for pid, r in pipes:
        with _os.fdopen(r, "rb") as fh:
            data = fh.read()
        _, status = _os.waitpid(pid, 0)
        if status != 0 or not data:
            raise RuntimeError("a forked vmap worker failed (see its traceback above)")
        for i, v in pickle.loads(data):
            outs[i] = v
    return outs


def vmap(f, in_axes=0, out_axes=0):
    """`jax.vmap` as a Python loop: slices every mapped leaf along its axis, calls `f`, stacks the results."""

    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(axes) == len(args), (len(axes), len(args))
        flat_args, flat_axes, size = [], [], None
        for arg, ax in zip(args, axes):
            leaves, treedef = tree_flatten(arg)
            if isinstance(ax, (tuple, list)) or (ax is not None and not isinstance(ax, int)):
                # a pytree of axes matching the argument (one entry per top-level field)
                ax_leaves = _broadcast_axes(arg, ax)
            else:
                ax_leaves = [ax] * len(leaves)
            for leaf, a in zip(leaves, ax_leaves):
                if a is not None:
                    n = _np.shape(leaf)[a]
                    assert size is None or size == n, (size, n)
                    size = n
            flat_args.append((leaves, treedef))
            flat_axes.append(ax_leaves)
        assert size is not None, "vmap needs at least one mapped argument"

        def one(i):
            call_args = []
            for (leaves, treedef), ax_leaves in zip(flat_args, flat_axes):
                sl = [leaf if a is None else _wrap(_np.take(_np.asarray(leaf), i, axis=a))
                      for leaf, a in zip(leaves, ax_leaves)]
                call_args.append(tree_unflatten(treedef, sl))
            return f(*call_args)

        global _VMAP_DEPTH
        procs = int(_os.environ.get("JAX_SHIM_PROCS", "1"))
        if procs > 1 and _VMAP_DEPTH == 0 and size >= 2 * procs:
            outs = _forked_map(one, size, procs)     # the OUTERMOST loop only, split over forked workers
        else:
            _VMAP_DEPTH += 1
            try:
                outs = [one(i) for i in range(size)]
            finally:
                _VMAP_DEPTH -= 1
        return tree_map(lambda *xs: _wrap(_np.stack([_np.asarray(x) for x in xs], axis=out_axes)), *outs)

    return mapped


def _broadcast_axes(arg, ax):
    """Expand an in_axes pytree prefix to one axis per leaf of `arg`."""
    if ax is None or isinstance(ax, int):
        return [ax] * len(tree_flatten(arg)[0])
    out = []
    if isinstance(arg, tuple):
        fields = list(arg)
        axes = list(ax)
        assert len(fields) == len(axes)
        for fld, a in zip(fields, axes):
            out += _broadcast_axes(fld, a)
        return out
    raise TypeError(f"unsupported in_axes structure for {type(arg)}")
