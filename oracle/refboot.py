"""TEST INFRASTRUCTURE ONLY -- boots the *unmodified* Python reference (Denizantip/py-numpy-renderer)
from /root/reference inside the build container and renders instrumented frames with it.

Nothing here is product code; nothing here runs on the GPU box (/root/reference does not exist there).
It is used by `oracle/make_golden.py` to produce the committed fixtures under `tests/golden/`.

Instrumentation is observation-only: `core.rasterize` / `triangular.general_shading` are wrapped so that the
frame / z / stencil arrays that `Scene.render` (core.py:587-640) allocates internally, and the identity of
the face that coloured every pixel in pass 1 and pass 3, can be read back.  No arithmetic is changed.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("B2R_REFERENCE_ROOT", "/root/reference")
ASSETS = os.path.join(REF_ROOT, "obj")

_booted = None


def available() -> bool:
    return os.path.isfile(os.path.join(ASSETS, "core.py"))


def boot():
    """Import the reference (SURVEY.md Appendix C recipe).  Returns a namespace of its public names."""
    global _booted
    if _booted is not None:
        return _booted
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    sys.path[:0] = [REF_ROOT, ASSETS]
    if "matplotlib" not in sys.modules:  # import-time-only dependency of triangular.py:2
        mp, mpp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.path")
        mpp.Path = object
        mp.path = mpp
        sys.modules["matplotlib"], sys.modules["matplotlib.path"] = mp, mpp
    import warnings
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import core  # noqa
    import triangular  # noqa
    from obj.constants import PROJECTION_TYPE, SYSTEM, SUBSYSTEM
    from obj.cube_map import CubeMap
    from obj.lightning import Lightning
    import transformation
    ns = types.SimpleNamespace(
        core=core, triangular=triangular, transformation=transformation,
        Camera=core.Camera, Light=core.Light, Model=core.Model, Scene=core.Scene,
        PROJECTION_TYPE=PROJECTION_TYPE, SYSTEM=SYSTEM, SUBSYSTEM=SUBSYSTEM,
        CubeMap=CubeMap, Lightning=Lightning,
        scale=transformation.scale, translation=transformation.translation,
        rotate_xyz=transformation.rotate_xyz)
    _booted = ns
    return ns


def instrumented_render(scene, quiet=True):
    """scene.render() of the reference, returning dict(rgb, frame_f32, z, stencil, winner1, winner3).

    winnerK[row, col] = global face index (models in add_model order, faces in array order) of the face whose
    `general_shading` call last wrote that pixel in pass K (1 = ambient pass, 3 = lit pass), -1 if none.
    """
    ref = boot()
    core, tri = ref.core, ref.triangular
    H, W = scene.resolution
    state = dict(frame=None, z=None, stencil=None, face_no=-1, seen_pass3=False)
    w1 = np.full((H, W), -1, np.int32)
    w3 = np.full((H, W), -1, np.int32)
    n_faces_total = sum(len(m._faces) for m in scene.models)

    orig_rasterize = core.rasterize
    orig_shading = tri.general_shading
    counters = dict(p1=0, p3=0)

    def rasterize_probe(face, frame, z_buffer, light, camera, stencil_buffer=None, debug_camera=None):
        state["frame"], state["z"] = frame, z_buffer
        if stencil_buffer is None:
            state["face_no"] = counters["p1"]
            counters["p1"] += 1
        else:
            state["stencil"] = stencil_buffer
            state["face_no"] = counters["p3"]
            counters["p3"] += 1
        return orig_rasterize(face, frame, z_buffer, light, camera, stencil_buffer, debug_camera=debug_camera)

    def shading_probe(face, bar, light, camera, frame, x, y, first_pass):
        (w1 if first_pass else w3)[x, y] = state["face_no"]
        return orig_shading(face, bar, light, camera, frame, x, y, first_pass)

    core.rasterize = rasterize_probe
    tri.general_shading = shading_probe
    try:
        sink = io.StringIO()
        with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
            rgb = scene.render()
    finally:
        core.rasterize = orig_rasterize
        tri.general_shading = orig_shading
    assert counters["p1"] == n_faces_total and counters["p3"] == n_faces_total
    stencil = state["stencil"] if state["stencil"] is not None else np.zeros((H, W), np.int16)
    return dict(rgb=rgb, frame_f32=state["frame"].copy(), z=state["z"].copy(), stencil=stencil.copy(),
                winner1=w1, winner3=w3, log=sink.getvalue())
