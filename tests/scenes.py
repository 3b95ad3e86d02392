"""Scene builders shared by tests and bench.py (product API only)."""
import os

import numpy as np

import py_numpy_renderer_b200 as b2r
from py_numpy_renderer_b200.materials import Material, Texture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def asset_root():
    """Directory holding diablo3_pose/, floor_diffuse.tga, skybox/ ... or None."""
    for cand in (os.environ.get("B2R_ASSETS"), os.path.join(ROOT, "baseline", "_ref", "assets"), "/root/reference/obj"):
        if cand and os.path.isfile(os.path.join(cand, "diablo3_pose", "diablo3_pose.obj")):
            return cand
    return None


FLOOR_V = np.array([[-2, -1, -2, 1], [2, -1, -2, 1], [2, -1, 2, 1], [-2, -1, 2, 1]], np.float32)
FLOOR_UV = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
FLOOR_N = np.array([[0, 1, 0]], np.float32)
FLOOR_F = np.array([[[0, 0, 0, 0], [3, 3, 0, 0], [2, 2, 0, 0]], [[0, 0, 0, 0], [2, 2, 0, 0], [1, 1, 0, 0]]], np.int32)


def floor_model(texels=None):
    """SURVEY.md Appendix C floor.obj (the reference's own is missing): y=-1 plane, x,z in [-2,2]."""
    m = b2r.Model(FLOOR_V.copy(), FLOOR_UV.copy(), FLOOR_N.copy(), FLOOR_F.copy())
    if texels is not None:
        m.materials['default'].map_Kd = Texture(texels, signed=False, tangent=False)
    return m


def std_cameras(position=(0.5, 1.5, 3), center=(0, 0, 0)):
    cam = b2r.Camera(position, center=np.array(center), fovy=60, near=0.1, far=10, backface_culling=True)
    dcam = b2r.Camera(position, center=np.array(center), fovy=90, near=0.05, far=20, backface_culling=True)
    return cam, dcam


def std_light(light_type=b2r.Lightning.POINT_LIGHTNING, **kw):
    args = dict(ambient_strength=0.1, linear=0.01, quadratic=0.001)
    args.update(kw)
    return b2r.Light((2, 3, 3), light_type=light_type, **args)


def kat1(assets, resolution=(800, 800)):
    d = b2r.Model.load_model(os.path.join(assets, 'diablo3_pose/diablo3_pose.obj'))
    cam, dcam = std_cameras()
    sc = b2r.Scene(cam, std_light(), debug_camera=dcam, resolution=resolution, system=b2r.SYSTEM.LH,
                   subsystem=b2r.SUBSYSTEM.OPENGL)
    sc.verbose = False
    sc.add_model(d)
    return sc


def kat2(assets, resolution=(1080, 1920), light_type=b2r.Lightning.POINT_LIGHTNING, center=(0, 0, 0)):
    d = b2r.Model.load_model(os.path.join(assets, 'diablo3_pose/diablo3_pose.obj'))
    d.textures.register('normals', os.path.join(assets, 'diablo3_pose/diablo3_pose_nm_tangent.tga'), tangent=True)
    d.textures.register('diffuse', os.path.join(assets, 'diablo3_pose/diablo3_pose_diffuse.tga'), normalize=False)
    fl = floor_model()
    fl.textures.register('diffuse', os.path.join(assets, 'floor_diffuse.tga'), normalize=False)
    cam, dcam = std_cameras()
    sc = b2r.Scene(cam, std_light(light_type, center=center), debug_camera=dcam, resolution=resolution,
                   system=b2r.SYSTEM.LH, subsystem=b2r.SUBSYSTEM.OPENGL)
    sc.verbose = False
    sc.add_model(d)
    sc.add_model(fl)
    return sc


# ---- synthetic stand-ins (no assets needed; BASELINE.json: "synthetic meshes of the named size") ---------------
def torus_arrays(nu, nv, R=1.0, r0=0.4, amp=0.05, fu=12, fv=9):
    """Closed displaced torus (SURVEY.md Appendix C generator): V = nu*nv vertices, F = 2*nu*nv triangles."""
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing='ij')
    u, v = 2 * np.pi * i / nu, 2 * np.pi * j / nv
    r = r0 + amp * np.sin(fu * u) * np.sin(fv * v)
    x, y, z = (R + r * np.cos(v)) * np.cos(u), r * np.sin(v), (R + r * np.cos(v)) * np.sin(u)
    verts = np.stack([x, y, z, np.ones_like(x)], -1).reshape(-1, 4).astype(np.float32)
    uv = np.stack([u / (2 * np.pi), v / (2 * np.pi), np.zeros_like(u)], -1).reshape(-1, 3).astype(np.float32)
    n = np.stack([x, y, z], -1) - np.stack([R * np.cos(u), np.zeros_like(u), R * np.sin(u)], -1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    normals = n.reshape(-1, 3).astype(np.float32)
    idx = lambda a, b: ((a % nu) * nv + (b % nv))
    a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
    tris = np.stack([np.stack([a, d, c], -1), np.stack([a, c, b], -1)], 2).reshape(-1, 3)
    faces = np.zeros((tris.shape[0], 3, 4), np.int32)
    faces[..., 0] = faces[..., 1] = faces[..., 2] = tris
    return verts, uv, normals, faces


def procedural_texture(size, seed, kind='color'):
    """Deterministic uint8 RGB texture (value noise + stripes); kind='normal' gives a tangent-space normal map."""
    rng = np.random.default_rng(seed)
    g = rng.random((size // 16 + 2, size // 16 + 2, 3))
    yy, xx = np.meshgrid(np.arange(size) / 16.0, np.arange(size) / 16.0, indexing='ij')
    y0, x0 = yy.astype(int), xx.astype(int)
    fy, fx = (yy - y0)[..., None], (xx - x0)[..., None]
    val = (g[y0, x0] * (1 - fy) * (1 - fx) + g[y0 + 1, x0] * fy * (1 - fx) + g[y0, x0 + 1] * (1 - fy) * fx
           + g[y0 + 1, x0 + 1] * fy * fx)
    if kind == 'normal':
        nx, ny = (val[..., 0] - 0.5) * 0.8, (val[..., 1] - 0.5) * 0.8
        nz = np.sqrt(np.clip(1 - nx * nx - ny * ny, 0, 1))
        val = np.stack([nx, ny, nz], -1) * 0.5 + 0.5
    else:
        val = 0.25 + 0.75 * val * (0.75 + 0.25 * np.sin(xx * 1.7)[..., None])
    return np.clip(np.rint(val * 255), 0, 255).astype(np.uint8)


def figure_model(nu=100, nv=25, tex=1024):
    """Diablo-class stand-in: 2*nu*nv = 5000 triangles, bumpy closed surface standing on the floor, with a
    tex x tex diffuse map and tangent-space normal map (sizes of diablo3_pose: 5022 tris, 1024^2 maps)."""
    v, uv, n, f = torus_arrays(nu, nv, R=0.55, r0=0.28, amp=0.07, fu=9, fv=5)
    # stand the ring up (swap y/z) so it faces the camera like a figure, bbox ~ +-0.9 x +-0.9 x +-0.35
    v = v[:, [0, 2, 1, 3]].copy()
    n = n[:, [0, 2, 1]].copy()
    f = f[:, ::-1, :].copy()  # the axis swap mirrors the surface: reverse the winding to keep outward normals
    m = b2r.Model(v, uv, n, f)
    m.materials['default'].norm = Texture(procedural_texture(tex, 1, 'normal'), signed=True, tangent=True)
    m.materials['default'].map_Kd = Texture(procedural_texture(tex, 2), signed=False, tangent=False)
    return m


def c3_synthetic(resolution=(1080, 1920), light_type=b2r.Lightning.POINT_LIGHTNING, tex=1024, nu=100, nv=25):
    """BASELINE.json config 3 with synthetic stand-ins: figure (5000 tris, diffuse + tangent normal map) + floor
    (2 tris, 600^2 diffuse), shadow volumes, one light, LH/OpenGL perspective camera."""
    cam, dcam = std_cameras()
    sc = b2r.Scene(cam, std_light(light_type), debug_camera=dcam, resolution=resolution, system=b2r.SYSTEM.LH,
                   subsystem=b2r.SUBSYSTEM.OPENGL)
    sc.verbose = False
    sc.add_model(figure_model(nu, nv, tex))
    sc.add_model(floor_model(procedural_texture(600 if tex >= 600 else tex, 3)))
    return sc


def orbit_cameras(n, radius=3.0, height=1.5, start=0.0, fovy=60, near=0.1, far=10):
    """n cameras on a circle looking at the origin (pos_k = (r sin t, h, r cos t))."""
    cams = []
    for k in range(n):
        t = start + 2 * np.pi * k / max(n, 1)
        cams.append(b2r.Camera((radius * np.sin(t), height, radius * np.cos(t)), center=np.array((0, 0, 0)),
                               fovy=fovy, near=near, far=far, backface_culling=True))
    return cams


# ---- seeded random scenes (fuzzing the parity of oracle / CUDA / reference) ------------------------------------
def random_scene_spec(seed):
    """A small random scene as plain data: meshes (closed blobs + loose triangles), textures, camera, light."""
    rng = np.random.default_rng(seed)
    spec = dict(seed=seed, models=[])
    n_models = int(rng.integers(1, 4))
    for mi in range(n_models):
        kind = rng.choice(['torus', 'soup', 'plane'])
        if kind == 'torus':
            nu, nv = int(rng.integers(5, 14)), int(rng.integers(4, 10))
            v, uv, n, f = torus_arrays(nu, nv, R=float(rng.uniform(0.4, 0.9)), r0=float(rng.uniform(0.15, 0.35)),
                                       amp=float(rng.uniform(0, 0.08)), fu=3, fv=2)
        elif kind == 'plane':
            s = float(rng.uniform(1, 4))
            v = (FLOOR_V * np.array([s / 2, 1, s / 2, 1], np.float32)).astype(np.float32)
            uv, n, f = FLOOR_UV.copy(), FLOOR_N.copy(), FLOOR_F.copy()
        else:
            nt = int(rng.integers(3, 40))
            v = np.concatenate([rng.uniform(-1.2, 1.2, (nt * 3, 3)), np.ones((nt * 3, 1))], 1).astype(np.float32)
            uv = np.concatenate([rng.uniform(0, 1, (nt * 3, 2)), np.zeros((nt * 3, 1))], 1).astype(np.float32)
            n = rng.standard_normal((nt * 3, 3)).astype(np.float32)
            n /= np.linalg.norm(n, axis=1, keepdims=True)
            f = np.zeros((nt, 3, 4), np.int32)
            f[..., 0] = f[..., 1] = f[..., 2] = np.arange(nt * 3).reshape(nt, 3)
            if nt > 4:
                f[1, 2, 0] = f[1, 1, 0]                      # a degenerate face (repeated vertex)
        offset = rng.uniform(-0.6, 0.6, 3) * (mi > 0)
        transform = None
        if rng.random() < 0.5:                               # float64 vertices through an `@` chain
            transform = dict(scale=float(rng.uniform(0.6, 1.3)), translation=offset.tolist(),
                             rotate=rng.uniform(-40, 40, 3).tolist())
        else:
            v = v.copy()
            v[:, :3] += offset.astype(np.float32)
        tex = {}
        if rng.random() < 0.6:
            tex['map_Kd'] = dict(texels=procedural_texture(32, int(rng.integers(1 << 30))), signed=False, tangent=False)
        r = rng.random()
        if r < 0.35:
            tex['norm'] = dict(texels=procedural_texture(32, int(rng.integers(1 << 30)), 'normal'), signed=True, tangent=True)
        elif r < 0.5:
            tex['norm'] = dict(texels=procedural_texture(16, int(rng.integers(1 << 30)), 'normal'), signed=True, tangent=False)
        if rng.random() < 0.3:
            tex['map_Ks'] = dict(texels=procedural_texture(16, int(rng.integers(1 << 30))), signed=False, tangent=False)
        spec['models'].append(dict(v=v, uv=uv, n=n if rng.random() < 0.85 or 'norm' in tex else None, f=f,
                                   transform=transform, tex=tex, clip=bool(rng.random() < 0.85),
                                   Ns=float(rng.choice([8, 32, 64, 12.5]))))
    ang = rng.uniform(0, 2 * np.pi)
    rad = rng.uniform(1.8, 4.0)
    pos = (float(rad * np.sin(ang)), float(rng.uniform(0.3, 2.5)), float(rad * np.cos(ang)))
    fovy = float(rng.uniform(40, 80))
    near, far = float(rng.uniform(0.05, 0.5)), float(rng.uniform(6, 15))
    spec['camera'] = dict(position=pos, center=rng.uniform(-0.3, 0.3, 3).tolist(), fovy=fovy, near=near, far=far,
                          backface_culling=bool(rng.random() < 0.7))
    # the debug frustum contains the camera frustum, so the reference's frustum overlay (core.py:638) draws nothing
    spec['debug_camera'] = dict(spec['camera'], fovy=fovy + 30, near=near / 2, far=far * 2)
    spec['light'] = dict(position=rng.uniform(-3, 3, 3).round(2).tolist(),
                         light_type=str(rng.choice(['POINT_LIGHTNING', 'SPOT_LIGHTNING', 'DIRECTIONAL_LIGHTNING'])),
                         center=rng.uniform(-0.5, 0.5, 3).round(2).tolist(), ambient_strength=float(rng.uniform(0, 0.3)),
                         linear=float(rng.uniform(0.001, 0.2)), quadratic=float(rng.uniform(0.0005, 0.1)),
                         specular_strength=float(rng.uniform(0.05, 0.6)), color=rng.uniform(0.5, 1, 3).round(2).tolist())
    spec['light']['position'][1] = abs(spec['light']['position'][1]) + 1.0
    combos = [('LH', 'OPENGL'), ('RH', 'OPENGL'), ('LH', 'DIRECTX'), ('RH', 'DIRECTX')]
    spec['system'], spec['subsystem'] = combos[int(rng.integers(4))]
    spec['resolution'] = (int(rng.integers(20, 120)), int(rng.integers(20, 160)))
    spec['skymap'] = rng.uniform(0, 1, 3).round(3).tolist() if rng.random() < 0.5 else None
    return spec


def build_random_scene(spec, api=None, texture_factory=None):
    """Instantiate `spec` through an API namespace: this package (default) or the booted reference."""
    api = api or b2r
    own = api is b2r
    models = []
    for m in spec['models']:
        model = api.Model(m['v'].copy(), m['uv'].copy(), None if m['n'] is None else m['n'].copy(), m['f'].copy(),
                          clip=m['clip'])
        mat = model.materials['default']
        object.__setattr__(mat, 'Ns', m['Ns'])
        for attr, t in m['tex'].items():
            if own:
                setattr(mat, attr, Texture(t['texels'], signed=t['signed'], tangent=t['tangent']))
            else:
                object.__setattr__(mat, attr, texture_factory(t))
        if m['transform'] is not None:
            tr = m['transform']
            model = model @ api.scale(tr['scale']) @ api.translation(tuple(tr['translation'])) @ api.rotate_xyz(tuple(tr['rotate']))
        models.append(model)

    def cam(kw):
        kw = dict(kw)
        return api.Camera(tuple(kw.pop('position')), center=np.array(kw.pop('center')), **kw)
    lk = dict(spec['light'])
    light = api.Light(tuple(lk.pop('position')), light_type=getattr(api.Lightning, lk.pop('light_type')),
                      center=tuple(lk.pop('center')), color=tuple(lk.pop('color')), **lk)
    scene = api.Scene(cam(spec['camera']), light, debug_camera=cam(spec['debug_camera']),
                      resolution=tuple(spec['resolution']), system=getattr(api.SYSTEM, spec['system']),
                      subsystem=getattr(api.SUBSYSTEM, spec['subsystem']), skymap=spec['skymap'])
    if own:
        scene.verbose = False
        scene.persist_silhouette = False
    for model in models:
        scene.add_model(model)
    return scene
