"""Host-side matrix builders (reference: obj/transformation.py).

These stay NumPy on the host on purpose: the kernels consume the *float64 matrices themselves*, so the only way
to be value-identical to the reference is to evaluate the same scalar expressions with the same library.  Row
vector convention throughout (`v' = v @ M`), so translations live in row 3.

Quirks preserved: `scale(2)` is an int64 matrix, `rotate_xyz` is float32 and its
per-axis matrices are named after the *other* axis' angle, `opengl_orthographicLH` is float32.
"""
import math

import numpy as np

from .constants import PROJECTION_TYPE, SUBSYSTEM, SYSTEM, X, Y, mat3x3


def normalize(a, axis=-1, order=2):
    """a / |a| with zero norms replaced by 1 (transformation.py:46-49)."""
    length = np.atleast_1d(np.linalg.norm(a, order, axis))
    length[length == 0] = 1
    return a / np.expand_dims(length, axis)


def barycentric(a, b, c, p):
    """Dot-product barycentrics with float32 intermediates (transformation.py:12-32).  Host helper used by the
    tests; the device and the oracle restate it scalar-wise."""
    e0, e1, e2 = b - a, c - a, p - a
    d00, d01, d11 = np.float32(e0 @ e0), np.float32(e0 @ e1), np.float32(e1 @ e1)
    d20, d21 = np.float32(e2 @ e0), np.float32(e2 @ e1)
    den = d00 * d11 - d01 * d01
    if den == 0:
        return None
    inv = 1.0 / den
    bv = (d11 * d20 - d01 * d21) * inv
    bw = (d00 * d21 - d01 * d20) * inv
    return np.array([1.0 - bv - bw, bv, bw]).T


def bound_box(vert, height, width):
    """Clamped, ceil-ed pixel box `[ceil(min), ceil(max))` (transformation.py:35-43)."""
    lo_x = vert[X].min().max(initial=0)
    hi_x = vert[X].max().min(initial=width)
    lo_y = vert[Y].min().max(initial=0)
    hi_y = vert[Y].max().min(initial=height)
    if lo_x > hi_x or lo_y > hi_y:
        return None
    return np.ceil((lo_x, hi_x, lo_y, hi_y)).astype(np.int32)


# ---- camera matrices (transformation.py:77-98, 123-136) -------------------------------------------------------
def looka_at_translate(eye):
    m = np.eye(4)
    m[3, :3] = -eye
    return m


def _unit3(x, y, z):
    """normalize() of one 3-vector with Python floats: sqrt((x*x + y*y) + z*z), zero length -> 1.  Bit-identical to
    the NumPy path (same IEEE operations in the same order, no FMA anywhere) at a fraction of the call overhead."""
    length = math.sqrt((x * x + y * y) + z * z)
    if length == 0:
        length = 1.0
    return x / length, y / length, z / length


def _look_at_basis(eye, center, up):
    """forward = normalize(center - eye); right = normalize(cross(up, forward)); new_up = cross(forward, right)
    (transformation.py:83-98), evaluated on Python floats -- same values as the NumPy expressions, ~15x faster,
    which matters when a camera-orbit batch builds hundreds of cameras per second (tests/test_host_api.py checks
    equality against the NumPy formulation)."""
    d = np.asarray(center) - np.asarray(eye)
    fx, fy, fz = _unit3(float(d[0]), float(d[1]), float(d[2]))
    ux, uy, uz = float(up[0]), float(up[1]), float(up[2])
    rx, ry, rz = _unit3(uy * fz - uz * fy, uz * fx - ux * fz, ux * fy - uy * fx)
    nx, ny, nz = fy * rz - fz * ry, fz * rx - fx * rz, fx * ry - fy * rx
    return (rx, ry, rz), (nx, ny, nz), (fx, fy, fz)


def look_at_rotate_lh(eye, center, up):
    r, u, f = _look_at_basis(eye, center, up)
    return np.array([[r[0], u[0], -f[0], 0.0], [r[1], u[1], -f[1], 0.0], [r[2], u[2], -f[2], 0.0], [0.0, 0.0, 0.0, 1.0]])


def look_at_rotate_rh(eye, center, up):
    r, u, f = _look_at_basis(eye, center, up)
    return np.array([[r[0], u[0], f[0], 0.0], [r[1], u[1], f[1], 0.0], [r[2], u[2], f[2], 0.0], [0.0, 0.0, 0.0, 1.0]])


def ViewPort(resolution, far, near, x_offset=0, y_offset=0):
    height, width = resolution
    depth = far - near
    return np.array([[width / 2, 0, 0, 0],
                     [0, height / 2, 0, 0],
                     [0, 0, depth / 2, 0],
                     [width / 2 + x_offset, height / 2 + y_offset, depth / 2, 1]])


# ---- projections (transformation.py:139-204) --------------------------------------------------------------------
def opengl_orthographicLH(fov, aspect_ratio, z_near, z_far):
    top = np.tan(np.radians(fov / 2.0)) * z_near
    right = top * aspect_ratio
    return np.array([[1 / right, 0, 0, 0],
                     [0, 1 / top, 0, 0],
                     [0, 0, -2 / (z_far - z_near), 0],
                     [0, 0, (z_far + z_near) / (z_far - z_near), 1]], dtype=np.float32)


def _perspective(fovy, aspect, m22, m32, m23):
    f = 1.0 / np.tan(np.radians(fovy) / 2.0)
    m = np.zeros((4, 4))
    m[0, 0] = f / aspect
    m[1, 1] = f
    m[2, 2] = m22
    m[3, 2] = m32
    m[2, 3] = m23
    return m


def opengl_perspectiveLH(fovy, aspect, z_near, z_far):
    return _perspective(fovy, aspect, -(z_far + z_near) / (z_far - z_near),
                        2.0 * z_far * z_near / (z_far - z_near), 1.0)


def opengl_perspectiveRH(fovy, aspect, z_near, z_far):
    return _perspective(fovy, aspect, -(z_far + z_near) / (z_far - z_near),
                        -2.0 * z_far * z_near / (z_far - z_near), -1.0)


def directx_perspectiveRH(fovy, aspect, z_near, z_far):
    return _perspective(fovy, aspect, z_far / (z_near - z_far), z_near * z_far / (z_near - z_far), -1.0)


def directx_perspectiveLH(fovy, aspect, z_near, z_far):
    return _perspective(fovy, aspect, -z_far / (z_far - z_near), z_near * z_far / (z_far - z_near), 1.0)


# Only OPENGL/LH has an orthographic builder; other combinations raise KeyError like the reference
# (transformation.py:346-361).
perspectives = {
    SUBSYSTEM.DIRECTX: {
        PROJECTION_TYPE.PERSPECTIVE: {SYSTEM.LH: directx_perspectiveLH, SYSTEM.RH: directx_perspectiveRH},
        PROJECTION_TYPE.ORTHOGRAPHIC: {}},
    SUBSYSTEM.OPENGL: {
        PROJECTION_TYPE.PERSPECTIVE: {SYSTEM.LH: opengl_perspectiveLH, SYSTEM.RH: opengl_perspectiveRH},
        PROJECTION_TYPE.ORTHOGRAPHIC: {SYSTEM.LH: opengl_orthographicLH}},
}


# ---- model transform chain factors (transformation.py:207-263) ------------------------------------------------
def scale(factor):
    m = np.array([[factor, 0, 0, 0],
                  [0, factor, 0, 0],
                  [0, 0, factor, 0],
                  [0, 0, 0, 1]])
    return m


def translation(vec):
    x, y, z = vec
    return np.array([[1, 0, 0, x],
                     [0, 1, 0, y],
                     [0, 0, 1, z],
                     [0, 0, 0, 1]]).T


def rotate_xyz(a):
    x, y, z = np.deg2rad(a)
    # NB the reference builds the matrix it calls rotate_x from angle y and vice versa (transformation.py:230-251).
    about_x = np.array([[1, 0, 0, 0],
                        [0, np.cos(y), -np.sin(y), 0],
                        [0, np.sin(y), np.cos(y), 0],
                        [0, 0, 0, 1]], dtype=np.float32).T
    about_y = np.array([[np.cos(x), 0, np.sin(x), 0],
                        [0, 1, 0, 0],
                        [-np.sin(x), 0, np.cos(x), 0],
                        [0, 0, 0, 1]], dtype=np.float32).T
    about_z = np.array([[np.cos(z), np.sin(z), 0, 0],
                        [-np.sin(z), np.cos(z), 0, 0],
                        [0, 0, 1, 0],
                        [0, 0, 0, 1]], dtype=np.float32).T
    return about_z @ about_y @ about_x


rotate = rotate_xyz  # README.md:15-16 calls it `rotate`; only `rotate_xyz` exists in the reference


def extract_frustum_planes(matrix):
    """left, right, bottom, top, near, far; each `(col3 +/- col_i) / |.|_2` over all 4 coefficients
    (plane_intersection.py:43-56).  np.linalg.norm of a 1-D vector is sqrt(dot(x, x)); the dot stays a BLAS call
    because its fused accumulation order is part of the value."""
    matrix = np.asarray(matrix, dtype=np.float64)
    planes = np.empty((6, 4))
    col_w = matrix[:, 3]
    for axis in range(3):
        col = matrix[:, axis]
        for k, p in enumerate((col_w + col, col_w - col)):
            planes[2 * axis + k] = p / math.sqrt(p.dot(p))
    return planes
