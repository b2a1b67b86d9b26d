"""Loader for the reference's pre-generated Brax scenes.

``test_resources/pre-gen-brax/inputs-{2,30}.zip`` of the reference hold a pickle of
``(list[ModelObject] batched over frames, CameraParameters batched over frames, targets)`` made of
jax arrays (``notebooks/Generate Data.ipynb``).  ``load_pregen`` un-pickles them WITHOUT jax (the
array reconstruction hook is mapped onto numpy) into this package's types, every leaf a torch tensor
with the leading frame axis -- ready for ``Renderer.get_camera_image`` (native batching).
"""
from __future__ import annotations

import io
import pickle
import zipfile
from typing import Any, List, Tuple

import numpy as np
import torch

from .model import Model, ModelObject
from .renderer import CameraParameters


def _np_reconstruct(*args: Any):
    from numpy._core.multiarray import _reconstruct   # numpy >= 2 (``numpy.core`` in older pickles)

    return _reconstruct(*args)


def _reconstruct_array(fun, args, arr_state, aval_state):  # jax._src.array._reconstruct_array
    # `fun` comes out of the pickle stream: only numpy's own array reconstructor may be called
    if fun is not _np_reconstruct:
        raise pickle.UnpicklingError("refusing to call a non-numpy array constructor from a pickle")
    a = fun(*args)
    if not isinstance(a, np.ndarray):
        raise pickle.UnpicklingError("array reconstruction did not yield a numpy array")
    a.__setstate__(arr_state)
    return a


# Exact (module, name) pairs the pre-generated files reference; everything else is refused.  (Whole-module
# allow-lists would expose builtins.eval / numpy helpers that execute code.)
_ALLOWED = {
    ("renderer.model", "Model"): Model,
    ("renderer.model", "ModelObject"): ModelObject,
    ("renderer.renderer", "CameraParameters"): CameraParameters,
    ("jax._src.array", "_reconstruct_array"): _reconstruct_array,
    ("numpy", "ndarray"): np.ndarray,
    ("numpy", "dtype"): np.dtype,
    ("numpy.core.multiarray", "_reconstruct"): _np_reconstruct,
    ("numpy._core.multiarray", "_reconstruct"): _np_reconstruct,
    ("collections", "OrderedDict"): __import__("collections").OrderedDict,
}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        try:
            return _ALLOWED[(module, name)]
        except KeyError:
            raise pickle.UnpicklingError(f"refusing to load {module}.{name}") from None


def _to_torch(x: Any) -> Any:
    if isinstance(x, np.ndarray):
        if x.dtype == np.float64:
            x = x.astype(np.float32)
        if x.dtype == np.int64:
            x = x.astype(np.int32)
        return torch.from_numpy(np.ascontiguousarray(x))
    if isinstance(x, tuple) and hasattr(x, "_fields"):
        return type(x)(*[_to_torch(v) for v in x])
    if isinstance(x, (list, tuple)):
        return type(x)(_to_torch(v) for v in x)
    return x


def load_pregen(path: str) -> Tuple[List[ModelObject], CameraParameters, torch.Tensor]:
    """Load ``inputs-*.zip`` -> ``(objects, camera_parameters, targets)``; geometry and maps (constant
    over frames in these files) are reduced to their first frame, transforms / camera keep the frame axis."""
    with zipfile.ZipFile(path) as z:
        payload = z.read(z.namelist()[0])
    objects, camera, targets = _Unpickler(io.BytesIO(payload)).load()
    objs: List[ModelObject] = []
    for o in objects:
        m = _to_torch(o.model)
        m = Model(*[(t[0] if isinstance(t, torch.Tensor) and bool((t == t[:1]).all()) else t) for t in m])
        objs.append(ModelObject(model=m, local_scaling=_to_torch(np.asarray(o.local_scaling)),
                                transform=_to_torch(np.asarray(o.transform)),
                                double_sided=_to_torch(np.asarray(o.double_sided))))
    cam = CameraParameters(*[_to_torch(np.asarray(v)) for v in camera])
    return objs, cam, _to_torch(np.asarray(targets))
