"""Rebuild a scene through the PRODUCT host API from a tests/golden/*.npz fixture (written by
oracle/make_golden.py from the unmodified Python reference) and expose the reference's outputs."""
import glob
import json
import os

import numpy as np

import py_numpy_renderer_b200 as b2r
from py_numpy_renderer_b200.materials import Material, Texture

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixture_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


class SkyFromTexels(b2r.CubeMap):
    def __init__(self, texels):
        self.texels = np.ascontiguousarray(texels, dtype=np.uint8)


def _camera(kw):
    kw = dict(kw)
    pos, center = kw.pop('position'), kw.pop('center')
    if 'projection_type' in kw:
        kw['projection_type'] = getattr(b2r.PROJECTION_TYPE, kw['projection_type'])
    return b2r.Camera(tuple(pos), center=np.array(center), **kw)


def _light(kw):
    kw = dict(kw)
    pos = kw.pop('position')
    kw['light_type'] = getattr(b2r.Lightning, kw.pop('light_type', 'POINT_LIGHTNING'))
    if 'center' in kw:
        kw['center'] = tuple(kw['center'])
    return b2r.Light(tuple(pos), **kw)


def load(name, resolution=None):
    """-> (scene, expected dict, meta).  `resolution` overrides the fixture's (expected outputs then don't apply)."""
    data = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(data['meta']))
    models = []
    for mi, mm in enumerate(meta['models']):
        p = f"m{mi}_"
        mats = {'default': Material()}
        for si, (gname, rec) in enumerate(zip(mm['groups'], mm['mats'])):
            mat = mats['default'] if gname == 'default' else Material()
            object.__setattr__(mat, 'Kd', np.array(rec['Kd']))
            object.__setattr__(mat, 'Ks', np.array(rec['Ks']))
            object.__setattr__(mat, 'Ns', rec['Ns'])
            for attr in ('map_Kd', 'map_Ks', 'norm'):
                if attr in rec:
                    t = rec[attr]
                    setattr(mat, attr, Texture(data[t['key']], signed=t['signed'], tangent=t['tangent']))
            mats[gname] = mat
        model = b2r.Model(data[p + 'vertices'], data[p + 'uv'] if p + 'uv' in data else None,
                          data[p + 'normals'] if p + 'normals' in data else None, data[p + 'faces'],
                          materials=mats, material_group=list(mm['groups']), clip=mm['clip'],
                          depth_test=mm.get('depth_test', True))
        models.append(model)
    skymap = meta.get('skymap')
    if 'sky_texels' in data:
        skymap = SkyFromTexels(data['sky_texels'])
    scene = b2r.Scene(_camera(meta['camera']), _light(meta['light']), debug_camera=_camera(meta['debug_camera']),
                      resolution=tuple(resolution or meta['resolution']), system=getattr(b2r.SYSTEM, meta['system']),
                      subsystem=getattr(b2r.SUBSYSTEM, meta['subsystem']), skymap=skymap)
    scene.verbose = False
    scene.shading = meta.get('shading', 'general')
    for m in models:
        scene.add_model(m)
    expected = {k: data[k] for k in ('rgb', 'z', 'stencil', 'winner1', 'winner3', 'ref_mvp', 'ref_mvp_dbg',
                                     'ref_viewport', 'ref_planes')}
    w3, w1 = expected['winner3'], expected['winner1']
    expected['winner'] = np.where(w3 >= 0, w3, w1)
    return scene, expected, meta


def compare_planes(got, exp, label=""):
    """Summarise z / stencil / winner / rgb differences between two result dicts."""
    rep = {}
    zg, ze = got['z'], exp['z']
    rep['z_mismatch'] = int((~((zg == ze) | (np.isnan(zg) & np.isnan(ze)))).sum())
    rep['stencil_mismatch'] = int((got['stencil'] != exp['stencil']).sum())
    rep['winner_mismatch'] = int((got['winner'] != exp['winner']).sum())
    d = np.abs(got['rgb'].astype(np.int16) - exp['rgb'].astype(np.int16)).max(axis=-1)
    rep['rgb_px_diff'] = int((d > 0).sum())
    rep['rgb_px_gt1'] = int((d > 1).sum())
    rep['rgb_max'] = int(d.max())
    rep['pixels'] = int(d.size)
    return rep


def oracle_frame(oracle, scene):
    """Oracle render of `scene` (one view) + the host-side frustum overlay of core.py:638 when it draws anything --
    i.e. what `Scene.render()` of the reference returns, as dict(rgb, z, stencil, winner, face_status, n_silhouette)."""
    from py_numpy_renderer_b200 import overlay, SYSTEM
    got = oracle.render_scene(scene, extra=True)
    scene.camera.scene = scene.debug_camera.scene = scene
    lines = overlay.segments(scene.camera, scene.debug_camera)
    if lines:
        frame = overlay.apply(got['frame_f32'], got['z'], scene.camera, scene.debug_camera,
                              1 if scene.system == SYSTEM.RH else -1, lines)
        got['rgb'] = overlay.tonemap(frame)
    return got
