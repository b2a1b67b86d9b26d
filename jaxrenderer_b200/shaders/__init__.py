"""The seven built-in shaders (``renderer/shaders/``)."""
from .depth import DepthExtraInput, DepthShader
from .gouraud import GouraudExtraInput, GouraudShader
from .gouraud_texture import GouraudTextureExtraInput, GouraudTextureShader
from .phong import PhongTextureExtraInput, PhongTextureShader
from .phong_darboux import PhongTextureDarbouxExtraInput, PhongTextureDarbouxShader
from .phong_reflection import PhongReflectionTextureExtraInput, PhongReflectionTextureShader
from .phong_reflection_shadow import (
    PhongReflectionShadowTextureExtraInput,
    PhongReflectionShadowTextureShader,
)

BUILTIN_SHADERS = (
    DepthShader, GouraudShader, GouraudTextureShader, PhongTextureShader,
    PhongTextureDarbouxShader, PhongReflectionTextureShader, PhongReflectionShadowTextureShader,
)
