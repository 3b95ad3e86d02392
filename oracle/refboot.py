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
# The reference's own .py files, unmodified, zipped by __graft_entry__.build() into the git-ignored baseline/_ref/ (which
# travels to the GPU box): lets bench.py time the NumPy reference itself there.  Python imports straight from the zip.
REF_ZIP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "ref_src.zip")

_booted = None


def available() -> bool:
    return os.path.isfile(os.path.join(ASSETS, "core.py"))


def importable() -> bool:
    """The reference can be imported: from /root/reference (build container) or from the staged zip (GPU box)."""
    return available() or os.path.isfile(REF_ZIP)


def boot():
    """Import the reference (SURVEY.md Appendix C recipe).  Returns a namespace of its public names."""
    global _booted
    if _booted is not None:
        return _booted
    if available():
        sys.path[:0] = [REF_ROOT, ASSETS]
    elif os.path.isfile(REF_ZIP):
        sys.path[:0] = [REF_ZIP, REF_ZIP + "/obj"]
    else:
        raise RuntimeError(f"reference not found under {REF_ROOT} nor staged as {REF_ZIP}")
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


def to_reference_scene(scene, camera=None, debug_camera=None):
    """Rebuild a product-side `py_numpy_renderer_b200.Scene` (host objects only) with the reference's OWN classes: fresh
    Models (model.silhouette persists in the reference, SURVEY B-3), float32 texture arrays as TextureMaps.register
    stores them (core.py:90-105), cameras / light with the same constructor arguments.  Used by bench.py's NumPy-reference
    arm and by tests; `camera` / `debug_camera` override the scene's own."""
    ref = boot()
    from py_numpy_renderer_b200.cube_map import CubeMap as MyCubeMap
    from py_numpy_renderer_b200.materials import Texture

    def ref_texture(t):
        arr = t.texels / 255
        if t.signed:
            arr = arr * 2 - 1
        return np.array(arr, dtype=np.dtype(np.float32, metadata={'tangent': bool(t.tangent)}))

    def ref_camera(c):
        return ref.Camera(tuple(np.asarray(c.position).tolist()), center=np.array(c.center), show=False,
                          backface_culling=c.backface_culling, x_offset=c.x_offset, y_offset=c.y_offset,
                          projection_type=c.projection_type, up=c.up,
                          near=c.near, far=c.far, fovy=c.fovy)

    models = []
    for m in scene.models:
        rm = ref.Model(np.array(m.vertices, copy=True), None if m.uv is None else np.array(m.uv, copy=True),
                       None if m.normals is None else np.array(m.normals, copy=True), np.array(m._faces, copy=True),
                       clip=m.clip, depth_test=m.depth_test)
        rm.material_group = list(m.material_group)
        mats = {}
        for name, mat in m.materials.items():
            rmat = type(rm.materials['default'])()
            for key in ("Kd", "Ks", "Ns"):
                object.__setattr__(rmat, key, getattr(mat, key))
            for key in ("map_Kd", "map_Ks", "norm"):
                t = getattr(mat, key, None)
                if isinstance(t, Texture):
                    object.__setattr__(rmat, key, ref_texture(t))
            mats[name] = rmat
        rm.materials = mats
        models.append(rm)
    L = scene.light
    light = ref.Light(tuple(np.asarray(L.position).tolist()), light_type=getattr(ref.Lightning, L.light_type.name),
                      center=tuple(np.asarray(L.center).tolist()), color=tuple(np.asarray(L.color).tolist()),
                      ambient_strength=0, diffuse=L.diffuse, specular_strength=L.specular_strength,
                      constant=L.constant, linear=L.linear, quadratic=L.quadratic)
    light.ambient = np.array(L.ambient, dtype=float)
    sky = scene.skybox
    if isinstance(sky, MyCubeMap):
        rsky = ref.CubeMap.__new__(ref.CubeMap)       # textures as CubeMap.__init__ leaves them (cube_map.py:22-61)
        rsky.textures = np.asarray(sky.texels) / 255
        rsky.faces = [np.array([[-1, 1, 1, 1], [1, 1, 1, 1], [-1, -1, 1, 1]]),      # cube_map.py:46-54
                      np.array([[1, 1, 1, 1], [1, -1, 1, 1], [-1, -1, 1, 1]])]
        sky = rsky
    rs = ref.Scene(ref_camera(camera or scene.camera), light, debug_camera=ref_camera(debug_camera or scene.debug_camera),
                   resolution=tuple(scene.resolution), system=scene.system, subsystem=scene.subsystem, skymap=sky)
    for rm in models:
        rs.add_model(rm)
    return rs
