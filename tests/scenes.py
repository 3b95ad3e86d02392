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
