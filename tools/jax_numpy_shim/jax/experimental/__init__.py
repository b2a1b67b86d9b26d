from . import checkify  # noqa: F401
