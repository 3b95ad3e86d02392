"""The reference-side binding as working code: keep the reference's OWN classes
(`obj/core.py`: `Model`, `Camera`, `Light`, `Scene`, `CubeMap`) and replace only the body of `Scene.render`
(`obj/core.py:587-640`) by a call through the C ABI of include/b2r.h.

    import core                                   # the unmodified reference
    from py_numpy_renderer_b200 import integration
    integration.patch_reference(core)             # core.Scene.render now runs on the B200
    frame = scene.render()                        # same uint8 (*resolution, 3) array, same printed lines

What the adapter has to bridge, and nothing else:
* textures: the reference stores float32 arrays `u8/255` (or `u8/255*2-1`) with a `tangent` dtype metadata flag
  (core.py:90-105, 321-348); the kernels fetch uint8 texels and rebuild that value per fetch, so the arrays are turned
  back into their uint8 source (exact: checked by a round trip, anything else raises);
* the cubemap's float64 faces (cube_map.py:22-61) likewise;
* `model.silhouette` lives on the reference's Model (core.py:251); here the persistent set lives with the device scene
  of the `Scene` object, which is equivalent as long as a Model is rendered through one Scene.
Cameras and lights cross unchanged: `_abi.pack_views` / `pack_frame_params` only read attributes the reference's own
objects have (`MVP`, `viewport`, `frustum_planes`, `position`, `near`, `far`, `backface_culling`, `light_type`, ...).
"""
from __future__ import annotations

import types
from typing import Iterable

import numpy as np

from . import _abi
from .materials import Texture


def _texture_from_reference(arr) -> Texture:
    """float32 (H, W, 3) array of the reference -> uint8 texels + decode flags (exact inverse of core.py:96-104)."""
    a = np.asarray(arr)
    meta = a.dtype.metadata or {}
    a64 = a.astype(np.float64)
    for signed in (False, True):
        u8 = np.rint(((a64 + 1) / 2 if signed else a64) * 255)
        if u8.min() < 0 or u8.max() > 255:
            continue
        u8 = u8.astype(np.uint8)
        back = u8 / 255
        back = np.array(back * 2 - 1 if signed else back, dtype=np.float32)
        if np.array_equal(back, a.astype(np.float32)):
            return Texture(u8, signed=signed, tangent=bool(meta.get('tangent', False)))
    raise ValueError("texture array is not f32(u8 / 255) nor f32(u8 / 255 * 2 - 1): cannot be handed to the uint8 kernels")


class _SkyShim:
    """What `_abi.PackedScene` reads of a cubemap: uint8 (6, S, S, 3) in the reference's slot order."""

    def __init__(self, ref_cubemap):
        tex = np.asarray(ref_cubemap.textures, dtype=np.float64)
        u8 = np.rint(tex * 255).astype(np.uint8)
        if not np.array_equal(u8 / 255, tex):
            raise ValueError("cubemap textures are not u8 / 255")
        self.texels = np.ascontiguousarray(u8)


def _model_shim(m, cache):
    """The attributes `_abi.PackedScene` reads of a Model, with the float textures replaced by `Texture` objects."""
    mats = {}
    for name, mat in m.materials.items():
        shim = types.SimpleNamespace(Kd=mat.Kd, Ks=mat.Ks, Ns=mat.Ns)
        for attr in ('map_Kd', 'map_Ks', 'norm'):
            if hasattr(mat, attr):
                arr = getattr(mat, attr)
                if id(arr) not in cache:
                    cache[id(arr)] = (arr, _texture_from_reference(arr))   # the array is kept alive next to its id
                setattr(shim, attr, cache[id(arr)][1])
        mats[name] = shim
    return types.SimpleNamespace(vertices=m.vertices, uv=m.uv, normals=m.normals, _faces=np.asarray(m._faces),
                                 materials=mats, material_group=list(m.material_group), clip=m.clip,
                                 depth_test=m.depth_test)


def render_reference_scene(scene, verbose=True) -> np.ndarray:
    """`Scene.render()` of the reference (core.py:587-640) for one of the reference's own Scene objects."""
    from . import _native, overlay
    is_cubemap = hasattr(scene.skybox, 'textures') and hasattr(scene.skybox, 'faces')
    state = scene.__dict__.get('_b2r_state')
    key = (tuple(id(m) for m in scene.models), tuple(id(m.vertices) for m in scene.models),
           tuple(id(getattr(mat, a, None)) for m in scene.models for mat in m.materials.values()
                 for a in ('map_Kd', 'map_Ks', 'norm')), id(scene.skybox) if is_cubemap else None)
    if state is None or state['key'] != key:         # models / textures / cubemap -> device (rebuilt when they change)
        if state is not None:
            state['dev'].close()
        cache = {}
        shims = [_model_shim(m, cache) for m in scene.models]
        dev = _native.DeviceScene(shims, _SkyShim(scene.skybox) if is_cubemap else None)
        state = scene.__dict__['_b2r_state'] = dict(key=key, dev=dev, shims=shims, cache=cache)
    dev = state['dev']
    scene.camera.scene = scene.debug_camera.scene = scene
    if is_cubemap:
        bg = (_abi.B2R_BG_CUBEMAP, (0.0, 0.0, 0.0))
    elif isinstance(scene.skybox, Iterable):
        bg = (_abi.B2R_BG_COLOR, tuple(np.array(scene.skybox, dtype=np.float32).tolist()))
    else:
        bg = (_abi.B2R_BG_COLOR, tuple(np.array([64 / 255, 0.5, 198 / 255], dtype=np.float32).tolist()))   # core.py:600
    lines = overlay.segments(scene.camera, scene.debug_camera)           # frustum overlay of core.py:638
    mode = 'overlay' if lines else ('status' if verbose else False)
    frames, info = dev.render([scene.camera], [scene.debug_camera], scene.light, scene.resolution, scene.system, bg,
                              persist_silhouette=True, want_debug=mode)
    if is_cubemap:
        scene.camera.lookat[3, :3] = 0                                    # side effect of cube_map.py:94-96
    if verbose:
        for line in _native.status_report(state['shims'], info['face_status'][0]):
            print(line)
    frame = frames[0]
    if lines:
        from .constants import SYSTEM
        canvas = overlay.apply(info['frame_f32'][0], info['z'][0], scene.camera, scene.debug_camera,
                               1 if scene.system == SYSTEM.RH else -1, lines)
        frame = overlay.tonemap(canvas)
    return frame


def patch_reference(core_module):
    """Replace `core.Scene.render` of the (imported, unmodified) reference by the B200 path.  Returns the original
    method so that a caller can restore it."""
    original = core_module.Scene.render
    core_module.Scene.render = lambda self: render_reference_scene(self)
    return original
