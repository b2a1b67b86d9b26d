"""`checkify.checkify(f)` -> f returning (error, value); `check` raises eagerly."""
import functools


class _NoError:
    def throw(self):
        pass

    def get(self):
        return None


def checkify(f, errors=None):
    @functools.wraps(f)
    def wrapper(*a, **k):
        return _NoError(), f(*a, **k)

    return wrapper


def check(pred, msg, *fmt_args, **fmt_kwargs):
    import numpy as np

    if not bool(np.all(np.asarray(pred))):
        raise ValueError(msg.format(*fmt_args, **fmt_kwargs) if (fmt_args or fmt_kwargs) else msg)


user_checks = index_checks = float_checks = frozenset()
