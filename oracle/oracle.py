"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of the CPU oracle (`oracle/oracle.c`).

Only tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline / --impl reference legs may import this.
It renders the same host `Scene` objects the product API takes (py_numpy_renderer_b200.core) on the CPU with the
reference's algorithm, returning rgb + the z / stencil / winner planes used for parity checks.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.orc_render.restype = C.c_int
        _lib.orc_render_view.restype = C.c_int
    return _lib


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def render_packed(packed, fp, views, threads=1, planes=True):
    """views: ctypes array of View.  Returns dict(rgb, z, stencil, winner, face_status, n_silhouette)."""
    n = len(views)
    H, W = fp.height, fp.width
    rgb = np.empty((n, H, W, 3), np.uint8)
    z = np.empty((n, H, W), np.float64) if planes else None
    st = np.empty((n, H, W), np.int16) if planes else None
    win = np.empty((n, H, W), np.int32) if planes else None
    status = np.zeros((n, max(1, packed.total_faces)), np.uint8)
    nsil = np.zeros((n, max(1, packed.n_models)), np.int32)
    rc = lib().orc_render(packed.models, C.c_int32(packed.n_models), packed.textures, C.c_int32(packed.n_textures),
                          packed.sky_ptr, C.byref(fp), views, C.c_int32(n), _ptr(rgb), _ptr(z), _ptr(st), _ptr(win),
                          _ptr(status), _ptr(nsil), C.c_int32(threads))
    if rc == 2:
        raise IndexError("oracle: texture lookup out of range (the reference raises IndexError here)")
    if rc != 0:
        raise RuntimeError("oracle failed")
    return dict(rgb=rgb, z=z, stencil=st, winner=win, face_status=status[:, :packed.total_faces],
                n_silhouette=nsil[:, :packed.n_models])


def render_scene(scene, cameras=None, threads=1, planes=True, extra=False):
    """Render a host-side `py_numpy_renderer_b200.Scene` with the oracle (fresh silhouette per view)."""
    from py_numpy_renderer_b200 import _abi
    from py_numpy_renderer_b200.cube_map import CubeMap
    cams = list(cameras) if cameras is not None else [scene.camera]
    sky = scene.skybox if isinstance(scene.skybox, CubeMap) else None
    packed = _abi.PackedScene(scene.models, sky)
    fp = _abi.pack_frame_params(scene.light, scene.resolution, scene._background(), False,
                                shading=getattr(scene, 'shading', 'general'))
    for cam in cams:
        cam.scene = scene
    views = (_abi.View * len(cams))(*[_abi.pack_view(c, scene.debug_camera, scene.system, sky is not None)
                                      for c in cams])
    if not extra:
        return render_packed(packed, fp, views, threads=threads, planes=planes)
    # single view with the float frame and the pass-1 winner plane (golden comparison against the reference)
    H, W = fp.height, fp.width
    rgb = np.empty((H, W, 3), np.uint8); z = np.empty((H, W)); st = np.empty((H, W), np.int16)
    win = np.empty((H, W), np.int32); win1 = np.empty((H, W), np.int32); frame = np.empty((H, W, 3), np.float32)
    status = np.zeros(max(1, packed.total_faces), np.uint8); nsil = np.zeros(max(1, packed.n_models), np.int32)
    rc = lib().orc_render_view(packed.models, C.c_int32(packed.n_models), packed.textures, C.c_int32(packed.n_textures),
                          packed.sky_ptr, C.byref(fp), C.byref(views[0]), _ptr(rgb), _ptr(z), _ptr(st), _ptr(win),
                          _ptr(status), _ptr(nsil), _ptr(frame), _ptr(win1))
    if rc == 2:
        raise IndexError("oracle: texture lookup out of range (the reference raises IndexError here)")
    return dict(rgb=rgb, z=z, stencil=st, winner=win, winner1=win1, frame_f32=frame,
                face_status=status[:packed.total_faces], n_silhouette=nsil[:packed.n_models])
