import numpy as _np

from .numpy import _canon


def result_type(*args):
    return _canon(_np.result_type(*[(_np.asarray(a).dtype if not isinstance(a, (type, _np.dtype)) else a)
                                    for a in args]))
