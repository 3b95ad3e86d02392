import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from py_numpy_renderer_b200.constants import *  # noqa: F401,F403,E402
