"""`from core import Camera, Light, Model, Scene` (README.md of the reference) -> the B200 implementation.
Put this directory on sys.path instead of the reference's `obj/`."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from py_numpy_renderer_b200.core import *  # noqa: F401,F403,E402
from py_numpy_renderer_b200.core import Camera, Face, Light, Model, Scene, TextureMaps, Bound  # noqa: F401,E402
from py_numpy_renderer_b200._native import Errors  # noqa: F401,E402
