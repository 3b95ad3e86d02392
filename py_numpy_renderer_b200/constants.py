"""Index aliases and enums of the reference's L0 layer (reference: obj/constants.py:5-31).

The reference indexes homogeneous point arrays with `(..., k)` tuples (`pts[X]` == `pts[..., 0]`) and models its
enums as plain int-valued classes; both are part of the public surface (`from constants import SYSTEM`), so the
same names and values are exported here.
"""
import numpy as np


def _col(k):
    return (Ellipsis, k)


X = U = _col(0)
Y = V = _col(1)
Z = _col(2)
W = _col(3)
W_COL = _col([3])
XY = _col((0, 1))
XZ = _col((0, 2))
YZ = _col((1, 2))
XYZ = _col(slice(None, 3))
XYZW = None
mat3x3 = (slice(None, 3), slice(None, 3))
add_dim = _col(np.newaxis)


class PROJECTION_TYPE:
    PERSPECTIVE = 1
    ORTHOGRAPHIC = 2


class SUBSYSTEM:
    DIRECTX = 1
    OPENGL = 2


class SYSTEM:
    LH = -1
    RH = 1
