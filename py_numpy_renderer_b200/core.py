"""Scene API -- the drop-in boundary (reference: obj/core.py).

Same names, argument meaning and error behaviour as the reference's `Model`, `TextureMaps`, `Camera`, `Light`,
`Scene`; everything below `Scene.render()` is replaced by sm_100a kernels reached through the C ABI in
`include/b2r.h` (see `_native.py`).  The classes here only (1) parse / hold host data exactly like the reference
does and (2) evaluate the reference's own NumPy expressions for the float64 camera / light matrices, which are
the kernels' inputs.

Quirks kept on purpose: `Scene(debug_camera=None)` raises AttributeError (B-1);
`shadows=` / `shadowing=` are dead flags (B-2); `model.silhouette` persists between `render()` calls so a second
render of the same models has no shadows (B-3, opt out with `Scene.persist_silhouette = False`);
`MVP`/`lookat` are cached; `Model.__matmul__` leaves normals alone (B-12).
"""
from __future__ import annotations

import os
from functools import cached_property
from typing import Iterable, List

import numpy as np
from PIL import Image

from .constants import PROJECTION_TYPE, SUBSYSTEM, SYSTEM
from .cube_map import CubeMap
from .lightning import Lightning
from .materials import Material, Texture
from .transformation import (ViewPort, extract_frustum_planes, look_at_rotate_lh, look_at_rotate_rh,
                             looka_at_translate, normalize, perspectives, scale)


class TextureMaps:
    """`model.textures.register(kind, path, normalize=True, tangent=False)` (core.py:77-105)."""
    texture_map = {'diffuse': 'map_Kd', 'ambient': 'map_Ka', 'specular': 'map_Ks', 'shininess': 'map_Ns',
                   'transparency': 'map_d', 'normals': 'norm'}

    def __init__(self, model):
        self.model = model

    def register(self, attr_name, path, normalize=True, tangent=False):
        if attr_name not in self.texture_map:
            raise ValueError(f"{attr_name} not recognized.\nSupported: {self.texture_map.keys()}")
        tex = Texture(self.load_texels(path), signed=bool(normalize), tangent=tangent)
        setattr(self.model.materials['default'], self.texture_map[attr_name], tex)
        self.model._device_dirty()

    @staticmethod
    def load_texels(name) -> np.ndarray:
        """uint8 RGB texels; the reference divides by 255 here (core.py:100-105), the kernels do it per fetch."""
        return np.asarray(Image.open(name).convert('RGB'))

    @staticmethod
    def load_texture(name):
        """Reference-shaped return value (float64 in [0,1]) for callers that used the static helper."""
        return TextureMaps.load_texels(name) / 255


def _fan(corners):
    """Fan triangulation of one OBJ polygon (core.py:72-74)."""
    return [[corners[0], corners[k], corners[k + 1]] for k in range(1, len(corners) - 1)]


class Face:
    """Read-only host view of one triangle (reference: `Face`, core.py:108-136).  The reference drives its three
    passes by iterating such objects; here they only serve inspection (`model.faces`) -- the kernels work on the packed
    arrays.  `material` follows the reference's lookup: the face's group name, falling back to 'default'."""

    def __init__(self, model, vi, ti=None, ni=None, mtl=(0,)):
        self.model = model
        self._vi, self._ti, self._ni = vi, ti, ni
        self.vertices = model.vertices[vi]
        self.world_vertices = self.vertices.copy()
        self.uv = None if model.uv is None else model.uv[ti]
        self.normals = None if model.normals is None else model.normals[ni]
        self.textures = model.textures
        self.material = model.materials.get(model.material_group[mtl[0]], model.materials['default'])

    def _unit_normal(self, pts):
        a, b, c = pts[..., :3]
        return normalize(np.cross(b - a, c - a)).squeeze()

    @property
    def unit_normal_world_space(self):
        """normalize((b - a) x (c - a)) of the untransformed copy, in the vertex dtype (core.py:127-130)."""
        return self._unit_normal(self.world_vertices)

    @property
    def unit_normal_current_space(self):
        """Same on `vertices`, which the reference overwrites with screen coordinates while rasterising (132-136)."""
        return self._unit_normal(self.vertices)


class Model:
    def __init__(self, vertices, uv, normals, faces, shadowing: bool = False, materials: dict = None,
                 material_group: list = None, clip=True, depth_test=True):
        self.vertices = vertices
        self.clip = clip
        self.depth_test = depth_test
        self.shadowing = shadowing  # dead flag in the reference too (core.py:243)
        self.uv = uv
        self.normals = normals
        self._faces = faces
        self.materials = materials or {'default': Material()}
        self.material_group = material_group or ['default']
        self.textures = TextureMaps(self)
        self.shape = None
        # persistent silhouette state, one int8 per undirected edge of the static edge table that lives with
        # the device scene; mirrors `model.silhouette` (core.py:251).  None = empty set.
        self._scene_refs = []

    # -- reference API -----------------------------------------------------------------------------------------
    @property
    def faces(self):
        """Iterator over `Face` views in file order (core.py:253-255)."""
        return (Face(self, *f.T) for f in self._faces)

    @classmethod
    def load_model(cls, name, shadowing=True, native=None):
        """Wavefront OBJ (+MTL) loader producing the arrays of core.py:257-318: vertices f32 (V,4) with w=1,
        uv f32 (T,3), normals f32 (N,3), faces int32 (F,3,4) = [v, vt, vn, material slot], 0-based, -1 = absent.

        `native`: True = the C++ tokenizer of libb2r.so (~50x faster on million-triangle files,
        identical arrays), False = the pure-Python parser below, None = native when the library is built."""
        if native is None:
            from . import _native
            native = os.path.exists(_native.LIB_PATH)
        if native:
            from . import _native
            vertices, uv, normals, faces, slot_names, libs = _native.load_obj(name)
            materials = {'default': Material()}
            for lib in libs:
                path = os.path.join(os.path.dirname(name), lib)
                if os.path.exists(path):
                    materials |= cls.parse_mtl(path)
            return Model(vertices, uv, normals, faces, shadowing, materials=materials, material_group=slot_names)
        verts, uvs, norms, faces = [], [], [], []
        slot_names = ['default']
        current = 'default'
        materials = {'default': Material()}
        here = os.path.dirname(name)
        with open(name) as fh:
            for line in fh:
                tok = line.split()
                if not tok:
                    continue
                head = tok[0]
                if head == 'v':
                    xyz = tok[1:]
                    verts.append(xyz + [1] if len(xyz) == 3 else xyz)
                elif head == 'vt':
                    st = tok[1:]
                    uvs.append(st + [0] if len(st) == 2 else st)
                elif head == 'vn':
                    norms.append(tok[1:])
                elif head == 'f':
                    slot = slot_names.index(current) + 1
                    corners = []
                    for ref in tok[1:]:
                        idx = [int(s) if s else -1 for s in ref.split('/')]
                        corners.append(idx + [slot])
                    faces.extend(_fan(corners))
                elif head == 'usemtl':
                    current = tok[1]
                    if current not in slot_names:
                        slot_names.append(current)
                elif head == 'mtllib':
                    lib = os.path.join(here, tok[1])
                    if os.path.exists(lib):
                        materials |= cls.parse_mtl(lib)
        vertices = np.array(verts, dtype=np.float32)
        faces = np.array(faces, dtype=np.int32)
        faces = np.where(faces > 0, faces - 1, faces)
        normals = np.array(norms, dtype=np.float32) if norms else None
        uv = np.array(uvs, dtype=np.float32) if uvs else None
        return Model(vertices, uv, normals, faces, shadowing, materials=materials, material_group=slot_names)

    @staticmethod
    def parse_mtl(mtllib) -> dict:
        """core.py:320-348.  Texture statements load next to the .mtl; `map_bump` becomes a tangent-space `norm`
        map that is NOT rescaled to [-1,1] (the reference only rescales in `register`)."""
        lib = {}
        here = os.path.dirname(mtllib)
        material = None
        with open(mtllib) as fh:
            for line in fh:
                if line.startswith('#') or line == '\n':
                    continue
                key, *val = line.split()
                if key == 'newmtl':
                    material = lib[val[0]] = Material()
                elif key.startswith('map') or key == 'disp':
                    path = os.path.join(here, val[0])
                    if os.path.exists(path):
                        tangent = key == 'map_bump'
                        setattr(material, 'norm' if tangent else key,
                                Texture(TextureMaps.load_texels(path), signed=False, tangent=tangent))
                    else:
                        print(f"{key} {path} is not found. Recommend manually assign texture by descriptor "
                              f"Model.texture.register")
                else:
                    setattr(material, key, val)
        return lib

    def __matmul__(self, other):
        self.vertices = self.vertices @ other
        self._device_dirty()
        return self

    # -- device mirror bookkeeping -----------------------------------------------------------------------------
    def _device_dirty(self):
        for scene in self._scene_refs:
            scene._invalidate_device()

    @property
    def silhouette(self):
        """Set of (a, b) vertex-index pairs currently in the persistent silhouette (core.py:251); read back from
        the device scene this model was last rendered with."""
        for scene in self._scene_refs:
            got = scene._silhouette_of(self)
            if got is not None:
                return got
        return set(getattr(self, '_silhouette_saved', None) or ())


_MEMO = {}


def _memo(build, *args):
    """Projection / viewport matrices depend on a handful of scalars shared by every camera of an orbit batch:
    build once per distinct argument tuple, hand out copies (callers may mutate their matrix like in the reference)."""
    try:
        key = (build, *[float(a) if not isinstance(a, tuple) else a for a in args])
        hit = _MEMO.get(key)
    except TypeError:
        return build(*args)
    if hit is None:
        if len(_MEMO) > 256:
            _MEMO.clear()
        hit = _MEMO[key] = build(*args)
    return hit.copy()


class PositionedObject:
    def __init__(self, position, center=np.array([0, 0, 0])):
        self.scene = None
        self.position = position
        self.center = center

    def direction_to(self, other):
        return normalize(self.direction - other)

    @property
    def direction(self):
        return normalize(self.position - self.center).ravel()

    def set_position(self, new_position):
        self.position = new_position
        return self


class TransformationMatrixMixin:
    """Camera / light matrices (core.py:373-429): projection table lookup, swapped-argument look-at, cached MVP."""

    def __init__(self, x_offset=0, y_offset=0, projection_type=PROJECTION_TYPE.PERSPECTIVE,
                 up=np.array([0, 1, 0]), near=0.001, far=6, fovy=90):
        self.up = up
        self.projection_type = projection_type
        self.near = np.linalg.norm(self.position) if projection_type == PROJECTION_TYPE.ORTHOGRAPHIC else near
        self.far = far
        self.fovy = fovy
        self.x_offset = x_offset
        self.y_offset = y_offset
        self.scene = None

    @property
    def projection(self):
        height, width = self.scene.resolution
        build = perspectives[self.scene.subsystem][self.projection_type][self.scene.system]
        return _memo(build, self.fovy, width / height, self.near, self.far)

    @property
    def rotate(self):
        if self.scene.system == SYSTEM.LH:
            return look_at_rotate_lh(self.center, self.position, self.up)
        elif self.scene.system == SYSTEM.RH:
            return look_at_rotate_rh(self.center, self.position, self.up)

    @property
    def translate(self):
        return looka_at_translate(self.position)

    @cached_property
    def lookat(self):
        return self.translate @ self.rotate

    @cached_property
    def MVP(self):
        return self.lookat @ self.projection

    @property
    def frustum_planes(self):
        return extract_frustum_planes(self.MVP)

    @property
    def viewport(self):
        return _memo(ViewPort, tuple(self.scene.resolution), self.far, self.near, self.x_offset, self.y_offset)


class Camera(PositionedObject, TransformationMatrixMixin):
    def __init__(self, position, center, show=False, backface_culling=True, **kwargs):
        PositionedObject.__init__(self, np.array(position), center)
        TransformationMatrixMixin.__init__(self, **kwargs)
        self.show = show
        self.backface_culling = backface_culling


class Light(PositionedObject, TransformationMatrixMixin):
    def __init__(self, position, light_type=Lightning.POINT_LIGHTNING, center=(0, 0, 0), color=(1., 1., 1.),
                 ambient_strength=0, diffuse=1, specular_strength=0.5, show=False, constant=1, linear=0.14,
                 quadratic=0.07, **kwargs):
        self.color = np.array(color)
        self.light_type = light_type
        PositionedObject.__init__(self, np.array(position), np.array(center))
        self.ambient = ambient_strength * self.color
        self.show = show
        self.diffuse = diffuse
        self.specular_strength = specular_strength
        self.constant = constant
        self.linear = linear
        self.quadratic = quadratic
        TransformationMatrixMixin.__init__(self, **kwargs)

    @staticmethod
    def smoothstep(edge0, edge1, x_array):
        t = np.clip((x_array - edge0) / (edge1 - edge0), 0.0, 1.0)
        return t * t * (3 - 2 * t)

    @staticmethod
    def reflect(I, N):
        """Mirror the rows of `I` about the rows of `N`, normalised (core.py:493-495; unused by render())."""
        return normalize(I - 2.0 * np.sum(N * I, axis=1)[..., np.newaxis] * N)

    def attenuation(self, fragment_position):
        distance = np.linalg.norm((self.position - fragment_position), axis=1)
        return 1.0 / (self.constant + distance * (self.linear + self.quadratic * distance))[..., np.newaxis]


class Bound:
    """Descriptor binding a camera / light to its scene (core.py:527-555).  State sits on the class-level
    descriptor like in the reference, so `Scene(debug_camera=None)` raises AttributeError here too.  The
    `show=True` gizmo models need OBJ files the reference does not ship -> NotImplementedError."""

    def __set__(self, instance, value):
        self.obj = value
        self.obj.scene = instance
        if getattr(value, 'show', False):
            raise NotImplementedError("show=True gizmos need obj_loader_test/{sphere,camera}.obj, which the "
                                      "reference does not ship; out of scope (SURVEY.md section 2, #17)")

    def __get__(self, instance, owner):
        return self.obj


class Scene:
    camera = Bound()
    light = Bound()
    debug_camera = Bound()

    def __init__(self, camera=None, light=None, shadows=False, debug_camera=None, resolution=(1500, 1500),
                 system=SYSTEM.RH, subsystem=SUBSYSTEM.DIRECTX, skymap=None):
        self.system = system
        self.subsystem = subsystem
        self.models: List[Model] = []
        self.camera = camera if camera is not None else Camera(position=(0, 0, 1), center=(0, 0, 0))
        self.light = light if light is not None else Light(position=(1, 1, 1))
        self.debug_camera = debug_camera
        self.resolution = resolution
        self.skybox = skymap
        # extensions (not in the reference)
        self.persist_silhouette = True   # Appendix B-3 behaviour; False = every render starts from an empty set
        self.verbose = True              # print the three per-model lines of core.py:634-636
        # 'general' = general_shading, what the reference's render() runs.  'flat' / 'gouraud' / 'pbr' = the functions of
        # triangular.py:174-263 whose calls sit commented out next to it (triangular.py:120-130): what the reference
        # renders with that line swapped in
        self.shading = 'general' 
        self._device = None
        self._persist_dirty = False      # the device scene holds a persistent silhouette the host has not seen

    def add_model(self, model: Model):
        self.models.append(model)
        model._scene_refs.append(self)
        self._invalidate_device()

    # -- device plumbing ---------------------------------------------------------------------------------------
    def _invalidate_device(self):
        """The device mirror is rebuilt after `model @ M`, `textures.register`, `add_model`.  In the reference
        `model.silhouette` lives on the Model and survives all of these (core.py:251), so the persistent set is read
        back here and restored into the new device scene."""
        if self._device is not None:
            if self._persist_dirty:
                for i, m in enumerate(self._device.models_py):
                    m._silhouette_saved = self._device.silhouette_of(i)
                self._persist_dirty = False
            self._device.close()
            self._device = None

    def _device_scene(self):
        from . import _native
        if self._device is None:
            self._device = _native.DeviceScene(self.models, self.skybox if isinstance(self.skybox, CubeMap) else None)
            saved = [getattr(m, '_silhouette_saved', None) for m in self.models]
            if any(saved):
                self._device.restore_silhouette(saved)
                self._persist_dirty = True
        return self._device

    def _silhouette_of(self, model):
        if self._device is None:
            return None
        return self._device.silhouette_of(self.models.index(model))

    def _background(self):
        from . import _abi
        if isinstance(self.skybox, CubeMap):
            return _abi.B2R_BG_CUBEMAP, (0.0, 0.0, 0.0)
        if isinstance(self.skybox, Iterable):
            return _abi.B2R_BG_COLOR, tuple(np.array(self.skybox, dtype=np.float32).tolist())
        return _abi.B2R_BG_COLOR, tuple(np.array([64 / 255, 0.5, 198 / 255], dtype=np.float32).tolist())

    def render(self, debug=None) -> np.ndarray:
        """`Scene.render()` (core.py:587-640): uint8 (*resolution, 3), rows flipped, `** 0.8 * 255` tonemap.

        `debug`, if a dict, receives the z / stencil / winner planes and per-face status of this frame."""
        from . import _native, overlay
        dev = self._device_scene()
        self.camera.scene = self.debug_camera.scene = self
        lines = overlay.segments(self.camera, self.debug_camera)  # frustum overlay of core.py:638; [] = nothing drawn
        mode = 'overlay' if lines else (True if debug is not None else ('status' if self.verbose else False))
        frames, info = dev.render([self.camera], [self.debug_camera], self.light, self.resolution, self.system,
                                  self._background(), persist_silhouette=self.persist_silhouette, want_debug=mode,
                                  shading=self.shading)
        self._persist_dirty = self._persist_dirty or bool(self.persist_silhouette)
        if isinstance(self.skybox, CubeMap):
            # fill_frame_from_skybox zeroes the translation row of the *cached* camera.lookat in place
            # (cube_map.py:94-96, Appendix B-3); keep the side effect for callers that look at it afterwards.
            self.camera.lookat[3, :3] = 0
        if self.verbose:
            for line in _native.status_report(self.models, info['face_status'][0]):
                print(line)
        frame = frames[0]
        if lines:  # rare path: the debug frustum pokes into the view -- replay the reference's line pass on the host
            canvas = overlay.apply(info['frame_f32'][0], info['z'][0], self.camera, self.debug_camera,
                                   1 if self.system == SYSTEM.RH else -1, lines)
            frame = overlay.tonemap(canvas)
        if debug is not None:
            debug.update({k: v[0] for k, v in info.items()})
        return frame

    def render_batch(self, cameras, debug_cameras=None, debug=None, out=None, band=None):
        """Extension: one frame per camera in `cameras` with this scene's models / light /
        resolution.  `debug_cameras`: one per camera (the second clip frustum of triangular.py:39,83-87), default
        = the scene's debug camera for all.  Every view starts from an empty silhouette set (fresh-Model semantics)
        and nothing is printed.  Returns uint8 (len(cameras), H, W, 3): a new NumPy array, or `out` when given --
        a (pinned) NumPy array is filled synchronously, a CUDA tensor is written asynchronously on the library
        stream (`_native.sync()` or stream ordering before use).  `band=(row0,row1)` restricts work to buffer rows
        [row0,row1) (multi-GPU screen split); other rows of the output are left untouched."""
        dev = self._device_scene()
        cameras = list(cameras)
        dcams = [self.debug_camera] * len(cameras) if debug_cameras is None else list(debug_cameras)
        for cam in cameras + dcams:
            cam.scene = self
        frames, info = dev.render(cameras, dcams, self.light, self.resolution, self.system, self._background(),
                                  persist_silhouette=False, want_debug=debug is not None, out=out, band=band,
                                  shading=self.shading)
        if debug is not None:
            debug.update(info)
        return frames

    def render_batch_async(self, cameras, debug_cameras=None, out=None, band=None):
        """`render_batch` without the final wait: the camera maths runs here, the kernels and the device-to-host
        copies are only ENQUEUED (library streams), and the call returns a handle whose `result()` blocks until the
        frames are in `out` (a pinned NumPy array the caller must leave alone until then).  Keep two batches in
        flight, each with its own `out`, and the host work of batch k+1 and the PCIe transfer of batch k overlap the
        rendering of batch k+1."""
        from . import _native
        dev = self._device_scene()
        cameras = list(cameras)
        dcams = [self.debug_camera] * len(cameras) if debug_cameras is None else list(debug_cameras)
        for cam in cameras + dcams:
            cam.scene = self
        if out is None:
            out = np.empty((len(cameras), int(self.resolution[0]), int(self.resolution[1]), 3), np.uint8)
        fp, views = dev.pack(cameras, dcams, self.light, self.resolution, self.system, self._background(),
                             persist_silhouette=False, band=band, shading=self.shading)
        frames, info = dev.render_packed(fp, views, out=out, wait=False)
        return _native.PendingFrames(dev.lib, info.get('ticket'), frames, info.get('keep'))
