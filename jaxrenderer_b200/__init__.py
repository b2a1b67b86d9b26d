"""jaxrenderer_b200 -- B200-native rasterisation path behind the jaxrenderer API.

Public names mirror ``renderer/__init__.py:1-73`` of the reference.
"""
from .geometry import Camera, normalise, quaternion, quaternion_mul, rotation_matrix
from .graph import graphed
from .model import MergedModel, Model, ModelObject, batch_models, merge_objects
from .pipeline import render
from .renderer import CameraParameters, LightParameters, Renderer, ShadowParameters
from .shader import MixerOutput, PerFragment, PerVertex, Shader, UnsupportedShaderError
from .shadow import Shadow
from .shapes import UpAxis, create_capsule, create_cube
from .types import Buffers, Colour, LightSource, SpecularMap, Texture, Vec3f
from .utils import build_texture_from_PyTinyrenderer, canvas_to_uint8_display, transpose_for_display

__all__ = [
    "Buffers", "Camera", "CameraParameters", "Colour", "SpecularMap", "Texture", "Vec3f", "LightParameters", "LightSource", "MergedModel",
    "MixerOutput", "Model", "ModelObject", "PerFragment", "PerVertex", "Renderer", "Shader",
    "Shadow", "ShadowParameters", "UnsupportedShaderError", "UpAxis", "batch_models",
    "build_texture_from_PyTinyrenderer", "canvas_to_uint8_display", "graphed", "create_capsule", "create_cube",
    "merge_objects", "normalise", "quaternion", "quaternion_mul", "render", "rotation_matrix",
    "transpose_for_display",
]
