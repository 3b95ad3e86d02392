"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED Python reference.

Runs only in the build container (needs /root/reference).  Every fixture holds the scene INPUTS (mesh arrays,
uint8 texels, camera / light parameters) and the reference's OUTPUTS (uint8 frame, z-buffer, stencil buffer,
per-pixel winner face of pass 1 and pass 3, and the float64 camera matrices), so the tests can rebuild the scene
through the product API anywhere and compare.

    python oracle/make_golden.py            # all fixtures
    python oracle/make_golden.py g1_diablo  # one

Textures are decimated (nearest, every k-th texel) and written as PNG to a scratch directory before the
reference loads them, so the fixtures stay small; the reference accepts any texture size.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import time

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refboot  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
A = refboot.ASSETS
SCRATCH = tempfile.mkdtemp(prefix="b2r_golden_")


def decimate(path, step, name):
    im = np.asarray(Image.open(path).convert('RGB'))[::step, ::step]
    out = os.path.join(SCRATCH, name + ".png")
    Image.fromarray(np.ascontiguousarray(im)).save(out)
    return out


FLOOR_OBJ = """v -2 -1 -2
v 2 -1 -2
v 2 -1 2
v -2 -1 2
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 1 0
f 1/1/1 4/4/1 3/3/1
f 1/1/1 3/3/1 2/2/1
"""


def floor_path():
    p = os.path.join(SCRATCH, "floor.obj")
    if not os.path.exists(p):
        with open(p, "w") as fh:
            fh.write(FLOOR_OBJ)
    return p


def cube_path(step):
    """obj_loader_test/cube.{obj,mtl} copied next to decimated copies of its two PNG maps."""
    d = os.path.join(SCRATCH, f"cube{step}")
    if not os.path.isdir(d):
        os.makedirs(d)
        src = A + '/obj_loader_test/'
        for name in ('cube.obj', 'cube.mtl'):
            with open(src + name) as fi, open(os.path.join(d, name), 'w') as fo:
                fo.write(fi.read())
        for name in ('container2.png', 'container2_specular.png'):
            im = np.asarray(Image.open(src + name).convert('RGB'))[::step, ::step]
            Image.fromarray(np.ascontiguousarray(im)).save(os.path.join(d, name))
    return os.path.join(d, 'cube.obj')


def torus_arrays(nu, nv, R=1.0, amp=0.05):
    """SURVEY.md Appendix C generator (deterministic, closed, displaced torus)."""
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing='ij')
    u, v = 2 * np.pi * i / nu, 2 * np.pi * j / nv
    r = 0.4 + amp * np.sin(12 * u) * np.sin(9 * v)
    x, y, z = (R + r * np.cos(v)) * np.cos(u), r * np.sin(v), (R + r * np.cos(v)) * np.sin(u)
    verts = np.stack([x, y, z, np.ones_like(x)], -1).reshape(-1, 4).astype(np.float32)
    uv = np.stack([u / (2 * np.pi), v / (2 * np.pi), np.zeros_like(u)], -1).reshape(-1, 3).astype(np.float32)
    c = np.stack([R * np.cos(u), np.zeros_like(u), R * np.sin(u)], -1)
    n = np.stack([x, y, z], -1) - c
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    normals = n.reshape(-1, 3).astype(np.float32)
    idx = lambda a, b: ((a % nu) * nv + (b % nv))
    a, b, c2, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
    tris = np.stack([np.stack([a, d, c2], -1), np.stack([a, c2, b], -1)], 2).reshape(-1, 3)
    faces = np.zeros((tris.shape[0], 3, 4), np.int32)
    faces[..., 0] = faces[..., 1] = faces[..., 2] = tris
    return verts, uv, normals, faces


def cam_kwargs(position, center=(0, 0, 0), **kw):
    d = dict(position=list(position), center=list(center))
    d.update(kw)
    return d


def make_camera(ref, kw):
    kw = dict(kw)
    pos, center = kw.pop('position'), kw.pop('center')
    if 'projection_type' in kw:
        kw['projection_type'] = getattr(ref.PROJECTION_TYPE, kw['projection_type'])
    return ref.Camera(tuple(pos), center=np.array(center), **kw)


def make_light(ref, kw):
    kw = dict(kw)
    pos = kw.pop('position')
    kw['light_type'] = getattr(ref.Lightning, kw.pop('light_type', 'POINT_LIGHTNING'))
    if 'center' in kw:
        kw['center'] = tuple(kw['center'])
    return ref.Light(tuple(pos), **kw)


# ---- fixture definitions ----------------------------------------------------------------------------------------
CAM = cam_kwargs((0.5, 1.5, 3), fovy=60, near=0.1, far=10, backface_culling=True)
DCAM = cam_kwargs((0.5, 1.5, 3), fovy=90, near=0.05, far=20, backface_culling=True)
LIGHT = dict(position=[2, 3, 3], light_type='POINT_LIGHTNING', ambient_strength=0.1, linear=0.01, quadratic=0.001)


def diablo(ref, textured, tex_step=4):
    m = ref.Model.load_model(A + '/diablo3_pose/diablo3_pose.obj')
    if textured:
        m.textures.register('normals', decimate(A + '/diablo3_pose/diablo3_pose_nm_tangent.tga', tex_step, 'd_nm'),
                            tangent=True)
        m.textures.register('diffuse', decimate(A + '/diablo3_pose/diablo3_pose_diffuse.tga', tex_step, 'd_kd'),
                            normalize=False)
    return m


def floor(ref, tex_step=4):
    m = ref.Model.load_model(floor_path())
    m.textures.register('diffuse', decimate(A + '/floor_diffuse.tga', tex_step, 'f_kd'), normalize=False)
    return m


def skybox(ref, step):
    names = dict(back='back', bottom='bottom', front='front', left='left', right='right', top='top')
    return ref.CubeMap(**{k: decimate(f"{A}/skybox/{v}.jpg", step, 'sky_' + k) for k, v in names.items()})


def fixtures(ref):
    T = ref.transformation
    yield dict(name='g1_diablo_plain', resolution=(200, 200), system='LH', subsystem='OPENGL', camera=CAM,
               debug_camera=DCAM, light=LIGHT, models=lambda: [diablo(ref, False)])
    yield dict(name='g2_diablo_floor_point', resolution=(270, 480), system='LH', subsystem='OPENGL', camera=CAM,
               debug_camera=DCAM, light=LIGHT, models=lambda: [diablo(ref, True), floor(ref)])
    yield dict(name='g3_diablo_floor_spot', resolution=(216, 384), system='LH', subsystem='OPENGL', camera=CAM,
               debug_camera=DCAM, light=dict(LIGHT, light_type='SPOT_LIGHTNING', center=[0, 0, 0]),
               models=lambda: [diablo(ref, True), floor(ref)])
    yield dict(name='g4_diablo_floor_directional', resolution=(216, 384), system='LH', subsystem='OPENGL',
               camera=CAM, debug_camera=DCAM,
               light=dict(LIGHT, light_type='DIRECTIONAL_LIGHTNING', center=[0, 0.5, 0.5], specular_strength=0.1),
               models=lambda: [diablo(ref, True), floor(ref)])
    sky_cam = cam_kwargs((0.3, 0.5, 1.5), center=(0, 0, -2), fovy=60, near=0.1, far=10, backface_culling=True)
    sky_dcam = cam_kwargs((0.3, 0.5, 1.5), center=(0, 0, -2), fovy=90, near=0.05, far=20, backface_culling=True)
    yield dict(name='g5_skybox_perspective', resolution=(180, 320), system='LH', subsystem='OPENGL', camera=sky_cam,
               debug_camera=sky_dcam, light=LIGHT, skybox=lambda: skybox(ref, 16),
               models=lambda: [diablo(ref, False) @ T.translation((0, 0, -2.))])
    yield dict(name='g6_skybox_orthographic', resolution=(135, 240), system='LH', subsystem='OPENGL',
               camera=dict(sky_cam, projection_type='ORTHOGRAPHIC'),
               debug_camera=dict(sky_dcam, projection_type='ORTHOGRAPHIC'),
               light=LIGHT, skybox=lambda: skybox(ref, 16),
               models=lambda: [diablo(ref, False) @ T.translation((0, 0, -2.))])
    cube_cam = cam_kwargs((2.5, 2.0, 4.0), fovy=50, near=0.5, far=20, backface_culling=True)
    cube_dcam = cam_kwargs((2.5, 2.0, 4.0), fovy=80, near=0.25, far=40, backface_culling=True)
    for system, subsystem in (('RH', 'DIRECTX'), ('LH', 'DIRECTX'), ('RH', 'OPENGL')):
        yield dict(name=f'g7_cube_mtl_{system.lower()}_{subsystem.lower()}', resolution=(160, 200), system=system,
                   subsystem=subsystem, camera=cube_cam, debug_camera=cube_dcam,
                   light=dict(position=[3, 4, 2], light_type='POINT_LIGHTNING', ambient_strength=0.2,
                              linear=0.02, quadratic=0.002),
                   models=lambda: [ref.Model.load_model(cube_path(5)) @ T.scale(0.5)
                                   @ T.rotate_xyz((20, 30, 10)), floor(ref)],
                   )
    yield dict(name='g8_torus_flat', resolution=(192, 256), system='LH', subsystem='OPENGL',
               camera=cam_kwargs((0.0, 1.5, 2.9), fovy=60, near=0.1, far=10, backface_culling=True),
               debug_camera=cam_kwargs((0.0, 1.5, 2.9), fovy=90, near=0.05, far=20, backface_culling=True),
               light=LIGHT, skymap=[0.1, 0.2, 0.3],
               models=lambda: [ref.Model(*torus_arrays(96, 48)), floor(ref)])
    # debug cameras whose frustum does NOT contain the camera frustum: the model is cut by the second clip test
    # (triangular.py:83-87) and the frustum overlay (core.py:638, frustums.py:46-103) draws its red lines
    yield dict(name='g10_overlay_top_camera', resolution=(240, 320), system='LH', subsystem='OPENGL', camera=CAM,
               debug_camera=cam_kwargs((0, 3, 0.01), fovy=80, near=1, far=3, backface_culling=True), light=LIGHT,
               models=lambda: [diablo(ref, False), floor(ref)])
    yield dict(name='g11_overlay_inner_frustum', resolution=(200, 260), system='RH', subsystem='DIRECTX',
               camera=cam_kwargs((2.5, 2.0, 4.0), fovy=50, near=0.5, far=20, backface_culling=True),
               debug_camera=cam_kwargs((1.0, 2.5, 3.0), fovy=35, near=1.5, far=5.5, backface_culling=True),
               light=dict(position=[3, 4, 2], light_type='POINT_LIGHTNING', ambient_strength=0.2, linear=0.02,
                          quadratic=0.002),
               models=lambda: [ref.Model.load_model(cube_path(5)) @ T.scale(0.5) @ T.rotate_xyz((20, 30, 10)),
                               floor(ref)])
    # Model(depth_test=False): decals in front of / behind the figure, before / between / after the writing models in
    # add_model order (the face that colours a pixel depends on that order)
    yield dict(name='g12_depth_test_false', resolution=(216, 288), system='LH', subsystem='OPENGL', camera=CAM,
               debug_camera=DCAM, light=LIGHT,
               models=lambda: [decal(ref, -0.6, 0.2, -0.3, 0.6, 0.9), diablo(ref, False),
                               decal(ref, -1.0, 1.0, -0.8, 1.0, -0.9), floor(ref), decal(ref, 0.3, 1.2, -0.9, 0.1, 0.3)])
    yield dict(name='g13_depth_test_false_rh', resolution=(150, 200), system='RH', subsystem='DIRECTX',
               camera=cam_kwargs((2.5, 2.0, 4.0), fovy=50, near=0.5, far=20, backface_culling=True),
               debug_camera=cam_kwargs((2.5, 2.0, 4.0), fovy=80, near=0.25, far=40, backface_culling=True),
               light=dict(position=[3, 4, 2], light_type='POINT_LIGHTNING', ambient_strength=0.2, linear=0.02,
                          quadratic=0.002),
               models=lambda: [ref.Model.load_model(cube_path(5)) @ T.scale(0.5) @ T.rotate_xyz((20, 30, 10)),
                               decal(ref, -1.5, 1.0, -0.5, 1.2, 0.8), floor(ref), decal(ref, -0.5, 1.5, -0.9, 0.6, -0.6)])
    # the shading functions whose calls are commented out next to general_shading (triangular.py:120-130, 174-263):
    # what the reference renders with that line swapped in
    for k, shading in enumerate(('flat', 'gouraud', 'pbr')):
        yield dict(name=f'g{14 + k}_shading_{shading}', resolution=(120, 160), system='LH', subsystem='OPENGL',
                   camera=cam_kwargs((0.0, 1.5, 2.9), fovy=60, near=0.1, far=10, backface_culling=True),
                   debug_camera=cam_kwargs((0.0, 1.5, 2.9), fovy=90, near=0.05, far=20, backface_culling=True),
                   light=LIGHT, skymap=[0.1, 0.2, 0.3], shading=shading,
                   models=lambda: [ref.Model(*torus_arrays(64, 32)), diablo(ref, False)])
    yield dict(name='g9_diablo_transformed', resolution=(180, 240), system='LH', subsystem='OPENGL', camera=CAM,
               debug_camera=DCAM, light=dict(LIGHT, position=[-1.5, 2.5, 2.0]),
               models=lambda: [diablo(ref, True, 8) @ T.scale(0.8) @ T.translation((0.1, 0.0, -0.2))
                               @ T.rotate_xyz((0, 25, 0)), floor(ref, 8)])


def decal(ref, x0, x1, y0, y1, z):
    """Double-sided quad in the plane z = const (one side survives culling) that is z-TESTED but never writes z:
    Model(depth_test=False), core.py:232-236 / triangular.py:117."""
    v = np.array([[x0, y0, z, 1], [x1, y0, z, 1], [x1, y1, z, 1], [x0, y1, z, 1]], np.float32)
    uv = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    n = np.array([[0, 0, 1]], np.float32)
    tris = [(0, 1, 2), (0, 2, 3), (0, 2, 1), (0, 3, 2)]
    f = np.zeros((4, 3, 4), np.int32)
    for i, t in enumerate(tris):
        f[i, :, 0] = f[i, :, 1] = t
    return ref.Model(v, uv, n, f, depth_test=False)


def texture_u8(arr):
    """Recover the uint8 texels + decode kind from a reference float32 texture array (exact round trip)."""
    arr = np.asarray(arr)
    for signed in (False, True):
        u8 = np.rint(((arr.astype(np.float64) + 1) / 2 if signed else arr.astype(np.float64)) * 255)
        if u8.min() < 0 or u8.max() > 255:
            continue
        u8 = u8.astype(np.uint8)
        back = u8 / 255
        back = np.array(back * 2 - 1 if signed else back, dtype=np.float32)
        if np.array_equal(back, arr):
            return u8, signed
    raise AssertionError("texture does not round-trip through uint8")


def dump_model(prefix, m, out):
    out[prefix + 'vertices'] = np.asarray(m.vertices)
    out[prefix + 'faces'] = np.asarray(m._faces)
    if m.uv is not None:
        out[prefix + 'uv'] = np.asarray(m.uv)
    if m.normals is not None:
        out[prefix + 'normals'] = np.asarray(m.normals)
    mats = []
    for si, name in enumerate(m.material_group):
        mat = m.materials.get(name, m.materials['default'])
        rec = dict(Kd=np.asarray(mat.Kd, dtype=np.float64).ravel().tolist(),
                   Ks=np.asarray(mat.Ks, dtype=np.float64).ravel().tolist(), Ns=float(mat.Ns))
        for attr in ('map_Kd', 'map_Ks', 'norm'):
            if hasattr(mat, attr):
                arr = getattr(mat, attr)
                u8, signed = texture_u8(arr)
                key = f"{prefix}s{si}_{attr}"
                out[key] = u8
                meta = arr.dtype.metadata or {}
                rec[attr] = dict(key=key, signed=bool(signed), tangent=bool(meta.get('tangent', False)))
        mats.append(rec)
    return dict(groups=list(m.material_group), mats=mats, clip=bool(m.clip), depth_test=bool(m.depth_test))


def render_with_shading(ref, scene, shading):
    """instrumented render; for shading != 'general' the call of general_shading in rasterize() (triangular.py:127) is
    replaced by the call the reference keeps commented out on the following lines (128-130)."""
    if shading == 'general':
        return refboot.instrumented_render(scene)
    tri = ref.triangular
    original = tri.general_shading

    def swapped(face, bar, light, camera, frame, x, y, first_pass):
        if shading == 'flat':
            return tri.flat_shading(face, light, frame, x, y)
        if shading == 'gouraud':
            return tri.gouraud(face, light, frame, bar, x, y)
        return tri.pbr(face, light, camera, frame, bar, x, y)
    tri.general_shading = swapped
    try:
        with np.errstate(all='ignore'):
            return refboot.instrumented_render(scene)
    finally:
        tri.general_shading = original


def generate(spec, ref):
    t0 = time.time()
    models = spec['models']()
    sky = spec['skybox']() if 'skybox' in spec else None
    cam, dcam = make_camera(ref, spec['camera']), make_camera(ref, spec['debug_camera'])
    light = make_light(ref, spec['light'])
    skymap = sky if sky is not None else spec.get('skymap')
    scene = ref.Scene(cam, light, debug_camera=dcam, resolution=tuple(spec['resolution']),
                      system=getattr(ref.SYSTEM, spec['system']), subsystem=getattr(ref.SUBSYSTEM, spec['subsystem']),
                      skymap=skymap)
    for m in models:
        scene.add_model(m)
    out = {}
    meta = dict(name=spec['name'], resolution=list(spec['resolution']), system=spec['system'],
                subsystem=spec['subsystem'], camera=spec['camera'], debug_camera=spec['debug_camera'],
                light=spec['light'], skymap=spec.get('skymap'), shading=spec.get('shading', 'general'), models=[])
    for mi, m in enumerate(models):
        meta['models'].append(dump_model(f"m{mi}_", m, out))
    if sky is not None:
        out['sky_texels'] = np.rint(sky.textures * 255).astype(np.uint8)
        assert np.array_equal(out['sky_texels'] / 255, sky.textures)
    # matrices BEFORE render (skybox fill mutates the cached lookat)
    out['ref_mvp'] = np.array(cam.MVP)
    out['ref_mvp_dbg'] = np.array(dcam.MVP)
    out['ref_viewport'] = np.array(cam.viewport)
    out['ref_planes'] = np.array(cam.frustum_planes)
    res = render_with_shading(ref, scene, spec.get('shading', 'general'))
    out['rgb'], out['z'], out['stencil'] = res['rgb'], res['z'], res['stencil']
    out['winner1'], out['winner3'] = res['winner1'], res['winner3']
    meta['n_silhouette'] = [len(m.silhouette) for m in models]
    meta['log'] = res['log']
    meta['numpy'] = np.__version__
    out['meta'] = np.array(json.dumps(meta))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, spec['name'] + '.npz')
    np.savez_compressed(path, **out)
    print(f"{spec['name']}: {time.time() - t0:.1f}s  {os.path.getsize(path) / 1024:.0f} KiB  "
          f"covered={(res['winner1'] >= 0).sum()} lit={(res['winner3'] >= 0).sum()} "
          f"stencil=[{res['stencil'].min()},{res['stencil'].max()}] sil={meta['n_silhouette']}")


if __name__ == '__main__':
    ref = refboot.boot()
    want = set(sys.argv[1:])
    for spec in fixtures(ref):
        if not want or spec['name'] in want or any(spec['name'].startswith(w) for w in want):
            generate(spec, ref)
