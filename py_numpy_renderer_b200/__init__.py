"""B200-native frame pipeline behind the py-numpy-renderer Python API.

    from py_numpy_renderer_b200 import Model, Camera, Light, Scene, Lightning, scale, translation, rotate

Host side = this package (pure Python / NumPy, mirrors obj/core.py); device side = `csrc/` reached through the
C ABI of `include/b2r.h`.  There is no CPU fallback: `Scene.render()` raises if the CUDA library is missing.
"""
from .constants import PROJECTION_TYPE, SUBSYSTEM, SYSTEM  # noqa: F401
from .lightning import Lightning  # noqa: F401
from .materials import Material, Texture  # noqa: F401
from .transformation import scale, translation, rotate, rotate_xyz  # noqa: F401
from .cube_map import CubeMap  # noqa: F401
from .core import Model, Face, Camera, Light, Scene, TextureMaps  # noqa: F401

__all__ = ["Model", "Face", "Camera", "Light", "Scene", "TextureMaps", "CubeMap", "Material", "Texture", "Lightning",
           "PROJECTION_TYPE", "SUBSYSTEM", "SYSTEM", "scale", "translation", "rotate", "rotate_xyz"]
