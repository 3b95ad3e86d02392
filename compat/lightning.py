"""`from lightning import Lightning` -- the SAME class object as py_numpy_renderer_b200.Lightning."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from py_numpy_renderer_b200.lightning import Lightning  # noqa: F401,E402
