"""Constant meshes (cube, capsule); tables live in ``_data/*.npz``."""
from .capsule import UpAxis, create_capsule
from .cube import create_cube

__all__ = ["UpAxis", "create_capsule", "create_cube"]
