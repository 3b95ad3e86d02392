"""ctypes mirror of `include/b2r.h` + packing of host `Model` / `Camera` / `Light` objects into its structs.

Pure host code (NumPy + ctypes), no CUDA: it is shared by the product binding (`_native.py`) and by the test-side
wrapper of the CPU oracle (`oracle/oracle.py`), so both consume byte-identical descriptors.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .constants import SYSTEM
from .lightning import Lightning
from .materials import Texture

B2R_ABI_VERSION = 6
B2R_ERR_INDEX = 2
B2R_F32, B2R_F64 = 0, 1
B2R_TEX_UNORM, B2R_TEX_SNORM = 0, 1
B2R_BG_COLOR, B2R_BG_CUBEMAP = 0, 1
B2R_FACE_RENDERED, B2R_FACE_BACK_FACE_CULLING, B2R_FACE_WRONG_MIN_MAX = 0, 1, 2
B2R_FACE_EMPTY_B, B2R_FACE_EMPTY_Z, B2R_FACE_CLIPPED = 4, 8, 16
B2R_MAX_STAGES = 64

_d = C.c_double
_i = C.c_int32


class TextureDesc(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("height", _i), ("width", _i), ("decode", _i), ("tangent", _i)]


class MaterialDesc(C.Structure):
    _fields_ = [("Kd", _d * 3), ("Ks", _d * 3), ("Ns", _d), ("map_Kd", _i), ("map_Ks", _i), ("norm", _i),
                ("reserved", _i), ("Pm", _d), ("Pr", _d), ("Ka", _d * 3)]


class ModelDesc(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("uv", C.c_void_p), ("normals", C.c_void_p), ("faces", C.c_void_p),
                ("materials", C.POINTER(MaterialDesc)),
                ("n_vertices", _i), ("n_uv", _i), ("n_normals", _i), ("n_faces", _i), ("n_materials", _i),
                ("vertex_dtype", _i), ("uv_dtype", _i), ("normal_dtype", _i), ("clip", _i), ("depth_test", _i)]


class CubemapDesc(C.Structure):
    _fields_ = [("faces", C.c_void_p), ("size", _i), ("reserved", _i)]


class View(C.Structure):
    _fields_ = [("mvp", _d * 16), ("mvp_dbg", _d * 16), ("viewport", _d * 16), ("planes", _d * 24),
                ("sky_inv", _d * 16), ("cam_pos", _d * 3), ("near_", _d), ("far_", _d), ("system", _i),
                ("backface_culling", _i)]


class LightDesc(C.Structure):
    _fields_ = [("position", _d * 3), ("direction", _d * 3), ("color", _d * 3), ("ambient", _d * 3),
                ("specular_strength", _d), ("constant", _d), ("linear", _d), ("quadratic", _d),
                ("spot_cos_outer", _d), ("spot_cos_inner", _d), ("type", _i), ("reserved", _i)]


class FrameParams(C.Structure):
    _fields_ = [("light", LightDesc), ("background", C.c_float * 3), ("bg_mode", _i), ("height", _i), ("width", _i),
                ("row_begin", _i), ("row_end", _i), ("persist_silhouette", _i), ("shading", _i)]


class DebugOut(C.Structure):
    _fields_ = [("z", C.c_void_p), ("stencil", C.c_void_p), ("winner", C.c_void_p), ("face_status", C.c_void_p),
                ("n_silhouette", C.c_void_p), ("frame_f32", C.c_void_p)]


def _real(arr, what):
    """Keep float32 / float64 storage (the kernels need to know which: arithmetic in the vertex / uv dtype is
    part of the reference's semantics, core.py:127-130,191-218); anything else is promoted like NumPy would."""
    arr = np.asarray(arr)
    if arr.dtype == np.float32:
        return np.ascontiguousarray(arr), B2R_F32
    if arr.dtype != np.float64:
        arr = arr.astype(np.float64)
    return np.ascontiguousarray(arr), B2R_F64


def _vec3(dst, src):
    src = np.asarray(src, dtype=np.float64).ravel()
    if src.size == 1:
        src = np.repeat(src, 3)
    for k in range(3):
        dst[k] = float(src[k])


class PackedScene:
    """Descriptor arrays for `b2r_scene_create` / the oracle, with the NumPy buffers they point at kept alive."""

    def __init__(self, models, skybox=None):
        self.keep = []
        self.textures_py: list[Texture] = []
        tex_index = {}

        def tex_id(t):
            if not isinstance(t, Texture):
                return -1
            if id(t) not in tex_index:
                tex_index[id(t)] = len(self.textures_py)
                self.textures_py.append(t)
            return tex_index[id(t)]

        self.n_models = len(models)
        self.models = (ModelDesc * max(1, self.n_models))()
        self.face_counts = []
        for mi, m in enumerate(models):
            md = self.models[mi]
            v, md.vertex_dtype = _real(m.vertices, 'vertices')
            if v.ndim != 2 or v.shape[1] != 4:
                raise ValueError("Model.vertices must have shape (V, 4)")
            faces = np.ascontiguousarray(m._faces, dtype=np.int32)
            if faces.ndim != 3 or faces.shape[1:] != (3, 4):
                raise ValueError("Model faces must have shape (F, 3, 4): [v, vt, vn, material] per corner")
            self.keep += [v, faces]
            md.vertices, md.n_vertices = v.ctypes.data, v.shape[0]
            md.faces, md.n_faces = faces.ctypes.data, faces.shape[0]
            self.face_counts.append(faces.shape[0])
            if m.uv is not None:
                uv, md.uv_dtype = _real(m.uv, 'uv')
                self.keep.append(uv)
                md.uv, md.n_uv = uv.ctypes.data, uv.shape[0]
            if m.normals is not None:
                nr, md.normal_dtype = _real(m.normals, 'normals')
                self.keep.append(nr)
                md.normals, md.n_normals = nr.ctypes.data, nr.shape[0]
            # Face.material = materials.get(material_group[slot], materials['default'])  (core.py:125)
            slots = list(m.material_group)
            mats = (MaterialDesc * len(slots))()
            for si, name in enumerate(slots):
                mat = m.materials.get(name, m.materials['default'])
                _vec3(mats[si].Kd, mat.Kd)
                _vec3(mats[si].Ks, mat.Ks)
                mats[si].Ns = float(mat.Ns)
                mats[si].Pm, mats[si].Pr = float(getattr(mat, 'Pm', 0.5)), float(getattr(mat, 'Pr', 0.5))
                _vec3(mats[si].Ka, getattr(mat, 'Ka', (0.3, 0, 0)))
                mats[si].map_Kd = tex_id(getattr(mat, 'map_Kd', None))
                mats[si].map_Ks = tex_id(getattr(mat, 'map_Ks', None))
                mats[si].norm = tex_id(getattr(mat, 'norm', None))
                if (mats[si].map_Kd >= 0 or mats[si].map_Ks >= 0 or mats[si].norm >= 0) and m.uv is None:
                    raise TypeError("textured material on a model without uv coordinates")
                if mats[si].norm >= 0 and self.textures_py[mats[si].norm].tangent and m.normals is None:
                    raise TypeError("tangent-space normal map needs vertex normals (core.py:193)")
            self.keep.append(mats)
            md.materials, md.n_materials = mats, len(slots)
            md.clip, md.depth_test = int(bool(m.clip)), int(bool(m.depth_test))
        self.n_textures = len(self.textures_py)
        self.textures = (TextureDesc * max(1, self.n_textures))()
        for ti, t in enumerate(self.textures_py):
            td = self.textures[ti]
            td.rgb = t.texels.ctypes.data
            td.height, td.width = t.texels.shape[:2]
            td.decode = B2R_TEX_SNORM if t.signed else B2R_TEX_UNORM
            td.tangent = int(t.tangent)
            self.keep.append(t.texels)
        self.sky = None
        if skybox is not None:
            self.sky = CubemapDesc()
            self.sky.faces = skybox.texels.ctypes.data
            self.sky.size = skybox.texels.shape[1]
            self.keep.append(skybox.texels)
        self.total_faces = int(sum(self.face_counts))

    @property
    def sky_ptr(self):
        return C.byref(self.sky) if self.sky is not None else None


VIEW_DTYPE = np.dtype([("mvp", "<f8", 16), ("mvp_dbg", "<f8", 16), ("viewport", "<f8", 16), ("planes", "<f8", 24),
                       ("sky_inv", "<f8", 16), ("cam_pos", "<f8", 3), ("near_", "<f8"), ("far_", "<f8"),
                       ("system", "<i4"), ("backface_culling", "<i4")], align=True)
assert VIEW_DTYPE.itemsize == C.sizeof(View)


def _fill_view(rec, camera, debug_camera, system, with_sky):
    """Everything the kernels read from a camera, evaluated with the reference's own expressions
    (core.py:394-429).  `camera.scene` / `debug_camera.scene` must already be bound."""
    rec["mvp"] = np.asarray(camera.MVP, dtype=np.float64).ravel()  # cached_property like the reference (B-3)
    rec["mvp_dbg"] = np.asarray(debug_camera.MVP, dtype=np.float64).ravel()
    rec["viewport"] = np.asarray(camera.viewport, dtype=np.float64).ravel()
    rec["planes"] = camera.frustum_planes.ravel()
    if with_sky:
        # cube_map.py:94-97: lookat with its translation row zeroed, times the projection, inverted
        look = np.array(camera.lookat, dtype=np.float64, copy=True)
        look[3, :3] = 0
        rec["sky_inv"] = np.linalg.inv(look @ camera.projection).ravel()
    pos = np.asarray(camera.position, dtype=np.float64).ravel()
    rec["cam_pos"] = pos if pos.size == 3 else np.repeat(pos, 3)
    rec["near_"], rec["far_"] = float(camera.near), float(camera.far)
    rec["system"] = 1 if system == SYSTEM.RH else -1
    rec["backface_culling"] = int(bool(camera.backface_culling))


def pack_views(cameras, debug_cameras, system, with_sky):
    """-> ctypes array of View (backed by one NumPy record array, filled field-wise)."""
    n = len(cameras)
    arr = np.zeros(n, dtype=VIEW_DTYPE)
    for k in range(n):
        _fill_view(arr[k], cameras[k], debug_cameras[k], system, with_sky)
    views = (View * n).from_buffer(arr)
    views._keep = arr
    return views


def pack_view(camera, debug_camera, system, with_sky: bool) -> View:
    return pack_views([camera], [debug_camera], system, with_sky)[0]


def pack_light(light) -> LightDesc:
    ld = LightDesc()
    _vec3(ld.position, light.position)
    _vec3(ld.direction, light.direction)
    _vec3(ld.color, light.color)
    _vec3(ld.ambient, light.ambient)
    ld.specular_strength = float(light.specular_strength)
    ld.constant, ld.linear, ld.quadratic = float(light.constant), float(light.linear), float(light.quadratic)
    ld.spot_cos_outer = float(np.cos(np.deg2rad(20)))   # triangular.py:158
    ld.spot_cos_inner = float(np.cos(np.deg2rad(10)))   # triangular.py:159
    lt = light.light_type
    ld.type = int(lt.value) if isinstance(lt, Lightning) else int(getattr(lt, 'value', lt))
    return ld


SHADING = {'general': 0, 'flat': 1, 'gouraud': 2, 'pbr': 3}


def pack_frame_params(light, resolution, background, persist_silhouette, band=None, shading='general') -> FrameParams:
    fp = FrameParams()
    if shading not in SHADING:
        raise ValueError(f"shading must be one of {sorted(SHADING)}")
    fp.shading = SHADING[shading]
    fp.light = pack_light(light)
    mode, color = background
    fp.bg_mode = mode
    for k in range(3):
        fp.background[k] = color[k]
    fp.height, fp.width = int(resolution[0]), int(resolution[1])
    fp.row_begin, fp.row_end = (0, fp.height) if band is None else (int(band[0]), int(band[1]))
    fp.persist_silhouette = int(bool(persist_silhouette))
    return fp
