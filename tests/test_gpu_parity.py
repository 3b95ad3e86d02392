"""Parity tests proper: the CUDA path (through the Python API -> C ABI -> sm_100a kernels) against
(1) the committed outputs of the unmodified reference and (2) the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north star): z-buffer (float64), stencil counts, z-test winners and per-face status bit-exact;
uint8 RGB within 1 LSB on >= 99.9 % of the pixels.  All run on the B200 box:  pytest -m gpu
"""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import golden_util as gu
import scenes
import py_numpy_renderer_b200 as b2r

pytestmark = pytest.mark.gpu
NAMES = gu.fixture_names()


def gpu_render(scene):
    scene.persist_silhouette = False
    dbg = {}
    rgb = scene.render(debug=dbg)
    # the production path (no debug planes) takes the stencil shortcuts (background pixels skipped, depth-range
    # rejection of whole quad/tile pairs): it must produce the very same frame
    assert np.array_equal(scene.render(), rgb)
    return dict(rgb=rgb, z=dbg['z'], stencil=dbg['stencil'], winner=dbg['winner'], face_status=dbg['face_status'],
                n_silhouette=dbg['n_silhouette'])


def oracle_render(oracle, scene, cameras=None):
    out = oracle.render_scene(scene, cameras=cameras, threads=4)
    return out


def assert_parity(got, want, exact_rgb=False):
    rep = gu.compare_planes(got, want)
    assert rep['z_mismatch'] == 0, rep
    assert rep['stencil_mismatch'] == 0, rep
    assert rep['winner_mismatch'] == 0, rep
    assert rep['rgb_px_gt1'] == 0, rep                        # never off by more than 1 LSB
    assert rep['rgb_px_diff'] * 1000 <= rep['pixels'], rep    # >= 99.9 % of the pixels identical
    if exact_rgb:
        assert rep['rgb_px_diff'] == 0, rep
    return rep


@pytest.mark.parametrize("name", NAMES)
def test_fixture_vs_reference_and_oracle(name, oracle):
    scene, exp, meta = gu.load(name)
    got = gpu_render(scene)
    assert_parity(got, exp)                                   # vs the unmodified Python reference
    want = gu.oracle_frame(oracle, scene)                     # oracle (+ host overlay pass where it applies)
    assert_parity(got, want)                                  # vs the oracle
    assert np.array_equal(got['face_status'], want['face_status'])
    assert list(got['n_silhouette']) == meta['n_silhouette']


@pytest.mark.parametrize("light", ["POINT_LIGHTNING", "SPOT_LIGHTNING", "DIRECTIONAL_LIGHTNING"])
def test_c3_synthetic_1080p_vs_oracle(light, oracle):
    """BASELINE config 3 at its full size (1080x1920, 5000-triangle figure + floor, shadow volumes)."""
    scene = scenes.c3_synthetic((1080, 1920), light_type=getattr(b2r.Lightning, light))
    got = gpu_render(scene)
    want = {k: v[0] for k, v in oracle_render(oracle, scene).items()}
    assert_parity(got, want)
    assert np.array_equal(got['face_status'], want['face_status'])
    assert (got['stencil'] != 0).sum() > 10000 and (got['winner'] >= 0).sum() > 500000


@pytest.mark.parametrize("dbg_cam", ["narrow", "offset", "near_far"])
def test_clip_elision_keeps_the_cut_of_the_debug_frustum_1080p(dbg_cam, oracle):
    """k_clip_elide drops the per-pixel clip test of a (screen-filling face, tile) pair only where all four corners prove
    that every pixel passes in BOTH frusta.  Here the debug camera's frustum cuts through the visible floor and figure
    (narrower / looking from the side / short far plane), so tiles along the cut must keep the test: z, stencil and
    winner bit-exact against the oracle at the size where the floor's faces are 'huge' (more than 256 tiles)."""
    scene = scenes.c3_synthetic((1080, 1920))
    if dbg_cam == "narrow":
        scene.debug_camera = b2r.Camera((0.5, 1.5, 3), center=np.array((0, 0, 0)), fovy=28, near=0.1, far=10, backface_culling=True)
    elif dbg_cam == "offset":
        scene.debug_camera = b2r.Camera((2.5, 1.0, 2.0), center=np.array((0.3, -0.6, 0)), fovy=50, near=0.1, far=10, backface_culling=True)
    else:
        scene.debug_camera = b2r.Camera((0.5, 1.5, 3), center=np.array((0, 0, 0)), fovy=90, near=2.9, far=3.9, backface_culling=True)
    got = gpu_render(scene)
    want = gu.oracle_frame(oracle, scene)
    rep = gu.compare_planes(got, want)
    assert rep['z_mismatch'] == 0 and rep['stencil_mismatch'] == 0 and rep['winner_mismatch'] == 0, rep
    assert rep['rgb_px_gt1'] == 0 and rep['rgb_px_diff'] * 1000 <= rep['pixels'], rep
    covered = int((got['winner'] >= 0).sum())
    full = int((gpu_render(scenes.c3_synthetic((1080, 1920)))['winner'] >= 0).sum())
    assert 10000 < covered < full - 10000, (covered, full)    # the cut really removes a good part of the frame


@pytest.mark.skipif(scenes.asset_root() is None, reason="reference assets not staged")
def test_kat2_real_assets_1080p(oracle):
    """The headline scene itself: diablo3_pose (diffuse + tangent normal map) + floor, 1080p (SURVEY KAT-2)."""
    scene = scenes.kat2(scenes.asset_root())
    got = gpu_render(scene)
    want = {k: v[0] for k, v in oracle_render(oracle, scene).items()}
    assert_parity(got, want)
    assert list(got['n_silhouette']) == [1381, 4]
    assert int((got['winner'] >= 0).sum()) == 619692
    st = got['face_status'][:5022]
    assert [(st == v).sum() for v in (0, 1, 8, 16)] == [1409, 2424, 1128, 61]


def test_c5_style_many_small_triangles(oracle):
    """Config-5 regime at reduced count: 2*300*150 = 90k-triangle torus, most faces cover 0-2 pixel centres."""
    v, uv, n, f = scenes.torus_arrays(300, 150)
    scene = scenes.c3_synthetic((540, 960))
    scene.models.clear()
    scene._invalidate_device()
    scene.add_model(b2r.Model(v, uv, n, f))
    cams = scenes.orbit_cameras(2, radius=2.9)
    dcams = scenes.orbit_cameras(2, radius=2.9, fovy=90, near=0.05, far=20)
    dbg = {}
    rgb = scene.render_batch(cams, debug_cameras=dcams, debug=dbg)
    for k, (cam, dcam) in enumerate(zip(cams, dcams)):
        scene.camera, scene.debug_camera = cam, dcam
        want = {kk: vv[0] for kk, vv in oracle_render(oracle, scene, cameras=[cam]).items()}
        got = dict(rgb=rgb[k], z=dbg['z'][k], stencil=dbg['stencil'][k], winner=dbg['winner'][k])
        assert_parity(got, want)
        assert np.array_equal(dbg['face_status'][k], want['face_status'])


def test_batch_equals_single_and_is_deterministic():
    scene = scenes.c3_synthetic((270, 480), tex=128)
    cams = scenes.orbit_cameras(5)
    dcams = scenes.orbit_cameras(5, fovy=90, near=0.05, far=20)
    a = scene.render_batch(cams, debug_cameras=dcams)
    b = scene.render_batch(cams, debug_cameras=dcams)
    assert np.array_equal(a, b)
    assert len({a[k].tobytes() for k in range(5)}) == 5
    for k in (0, 3):
        scene.camera, scene.debug_camera = cams[k], dcams[k]
        scene.persist_silhouette = False
        assert np.array_equal(scene.render(), a[k])


def test_row_bands_compose_to_the_full_frame():
    """Screen-band sharding is exact: bands rendered separately are byte-identical to one frame."""
    scene = scenes.c3_synthetic((300, 500), tex=128)
    cams, dcams = [scene.camera], [scene.debug_camera]
    full = scene.render_batch(cams, debug_cameras=dcams)
    H = 300
    out = np.zeros_like(full)
    for r0, r1 in ((0, 77), (77, 200), (200, H)):             # deliberately not tile aligned
        part = scene.render_batch(cams, debug_cameras=dcams, band=(r0, r1))
        rows = slice(H - r1, H - r0)                           # output rows are flipped buffer rows
        assert not part[0, :H - r1].any() and not part[0, H - r0:].any()
        out[0, rows] = part[0, rows]
    assert np.array_equal(out, full)


def test_persistent_silhouette_quirk():
    """model.silhouette survives render(); the second call toggles it empty (no shadows),
    the third restores it."""
    scene, exp, meta = gu.load("g1_diablo_plain")
    scene.persist_silhouette = True
    sizes, frames, stencils = [], [], []
    for _ in range(3):
        dbg = {}
        frames.append(scene.render(debug=dbg))
        stencils.append(dbg['stencil'])
        sizes.append(len(scene.models[0].silhouette))
    assert sizes == [1381, 0, 1381]
    assert np.array_equal(frames[0], exp['rgb']) and np.array_equal(frames[2], exp['rgb'])
    assert not stencils[1].any() and stencils[0].any()
    assert not np.array_equal(frames[0], frames[1])


def test_device_resident_output_matches_host():
    torch = pytest.importorskip("torch")
    from py_numpy_renderer_b200 import _native
    scene = scenes.c3_synthetic((270, 480), tex=128)
    cams = scenes.orbit_cameras(3)
    dcams = scenes.orbit_cameras(3, fovy=90, near=0.05, far=20)
    host = scene.render_batch(cams, debug_cameras=dcams)
    dev = torch.zeros((3, 270, 480, 3), dtype=torch.uint8, device="cuda:0")
    torch.cuda.synchronize()
    scene.render_batch(cams, debug_cameras=dcams, out=dev)
    _native.sync()
    assert np.array_equal(dev.cpu().numpy(), host)
    pinned = torch.empty((3, 270, 480, 3), dtype=torch.uint8, pin_memory=True).numpy()
    scene.render_batch(cams, debug_cameras=dcams, out=pinned)
    assert np.array_equal(pinned, host)


@pytest.mark.parametrize("resolution", [(1, 1), (7, 5), (33, 31), (64, 96), (100, 333)])
def test_odd_resolutions(resolution, oracle):
    """Ragged sizes: not a multiple of the tile, row length not a multiple of 4 bytes, single pixel."""
    scene, _, _ = gu.load("g7_cube_mtl_rh_directx", resolution=resolution)
    got = gpu_render(scene)
    want = {k: v[0] for k, v in oracle_render(oracle, scene).items()}
    assert_parity(got, want)


def test_empty_and_offscreen_scenes(oracle):
    cam, dcam = scenes.std_cameras()
    scene = b2r.Scene(cam, scenes.std_light(), debug_camera=dcam, resolution=(48, 64), system=b2r.SYSTEM.LH,
                      subsystem=b2r.SUBSYSTEM.OPENGL, skymap=[0.2, 0.4, 0.6])
    scene.verbose = False
    scene.persist_silhouette = False
    empty = scene.render()                                     # no models at all: background only
    assert empty.shape == (48, 64, 3) and (empty == empty[0, 0]).all()
    far_away = scenes.floor_model() @ b2r.translation((500.0, 0, 0))       # entirely outside the frustum
    behind = scenes.floor_model() @ b2r.translation((0, 0, 40.0))          # behind the camera (w < 0)
    degenerate = b2r.Model(np.array([[0, 0, 0, 1], [0, 0, 0, 1], [1, 1, 0, 1]], np.float32), None, None,
                           np.zeros((1, 3, 4), np.int32) + np.array([0, 1, 2])[None, :, None] * np.array([1, 0, 0, 0]))
    for m in (far_away, behind, degenerate):
        scene.add_model(m)
    got = gpu_render(scene)
    want = {k: v[0] for k, v in oracle_render(oracle, scene).items()}
    assert_parity(got, want)
    assert np.array_equal(got['face_status'], want['face_status'])


def test_c5_one_million_triangles_one_frame(oracle):
    """BASELINE config 5 at its full mesh size: nu=1000 x nv=500 displaced torus = 1 000 000 triangles, 1080p."""
    v, uv, n, f = scenes.torus_arrays(1000, 500)
    cams = scenes.orbit_cameras(1, radius=2.9, start=0.3)
    dcams = scenes.orbit_cameras(1, radius=2.9, start=0.3, fovy=90, near=0.05, far=20)
    scene = b2r.Scene(cams[0], scenes.std_light(), debug_camera=dcams[0], resolution=(1080, 1920),
                      system=b2r.SYSTEM.LH, subsystem=b2r.SUBSYSTEM.OPENGL)
    scene.verbose = False
    scene.add_model(b2r.Model(v, uv, n, f))
    got = gpu_render(scene)
    want = {k: vv[0] for k, vv in oracle.render_scene(scene, threads=1).items()}
    assert_parity(got, want)
    assert np.array_equal(got['face_status'], want['face_status'])
    assert list(got['n_silhouette']) == list(want['n_silhouette']) and got['n_silhouette'][0] > 10000


@pytest.mark.parametrize("projection", ["PERSPECTIVE", "ORTHOGRAPHIC"])
def test_c4_skybox_4k(projection, oracle):
    """BASELINE config 4 at its full frame size: 2160x3840 with a cubemap skybox (procedural 256^2 faces), perspective
    and orthographic cameras.  At 4K the integer dot products of the skybox barycentrics exceed 2^24, so the float32
    rounding of `np.float32(int64)` (transformation.py:19-23 via cube_map.py:88) is exercised for real."""
    faces = np.stack([scenes.procedural_texture(256, 50 + k) for k in range(6)])
    v, uv, n, f = scenes.torus_arrays(60, 30, R=0.6, r0=0.25)
    kw = dict(fovy=60, near=0.1, far=10, backface_culling=True)
    dkw = dict(fovy=90, near=0.05, far=20, backface_culling=True)
    if projection == "ORTHOGRAPHIC":
        kw['projection_type'] = dkw['projection_type'] = b2r.PROJECTION_TYPE.ORTHOGRAPHIC
    cam = b2r.Camera((0.3, 0.5, 1.5), center=np.array((0, 0, -2)), **kw)
    dcam = b2r.Camera((0.3, 0.5, 1.5), center=np.array((0, 0, -2)), **dkw)
    scene = b2r.Scene(cam, scenes.std_light(), debug_camera=dcam, resolution=(2160, 3840), system=b2r.SYSTEM.LH,
                      subsystem=b2r.SUBSYSTEM.OPENGL, skymap=gu.SkyFromTexels(faces))
    scene.verbose = False
    scene.add_model(b2r.Model(v, uv, n, f) @ b2r.translation((0, 0, -2.)))
    got = gpu_render(scene)
    want = gu.oracle_frame(oracle, scene)
    assert_parity(got, want)
    assert len(np.unique(got['rgb'].reshape(-1, 3), axis=0)) > 1000     # the skybox really is sampled


@pytest.mark.parametrize("name", ["g2_diablo_floor_point", "g7_cube_mtl_rh_directx"])
def test_verbose_render_prints_the_reference_log(name, capsys):
    """core.py:624-636 prints three lines per model (total / rendered / discarded-by-reason); with `verbose` on, the
    same text comes out, produced from the per-face status bytes of the device."""
    scene, exp, meta = gu.load(name)
    scene.verbose = True
    scene.persist_silhouette = False
    capsys.readouterr()
    rgb = scene.render()
    out = capsys.readouterr().out
    assert out == meta['log']
    assert np.array_equal(rgb, exp['rgb'])


_OVERFLOW_SCRIPT = r"""
import sys, numpy as np
sys.path[:0] = [{root!r}, {root!r} + '/tests']
import golden_util as gu
import py_numpy_renderer_b200 as b2r
from py_numpy_renderer_b200 import _native
scene, exp, meta = gu.load('g2_diablo_floor_point')
scene.persist_silhouette = False
# synchronous call: the lists overflow, the call grows them and renders again before it returns
dbg = {{}}
rgb = scene.render(debug=dbg)
rep = gu.compare_planes(dict(rgb=rgb, z=dbg['z'], stencil=dbg['stencil'], winner=dbg['winner']), exp)
assert rep['z_mismatch'] == 0 and rep['stencil_mismatch'] == 0 and rep['winner_mismatch'] == 0 and rep['rgb_px_gt1'] == 0, rep
# asynchronous device-resident call on a fresh scene: the overflow is reported by the next sync, the retry is right
scene2, exp2, _ = gu.load('g2_diablo_floor_point')
scene2.persist_silhouette = False
import torch
out = torch.empty((1, *scene2.resolution, 3), dtype=torch.uint8, device='cuda:0')
scene2.render_batch([scene2.camera], debug_cameras=[scene2.debug_camera], out=out)
try:
    _native.sync()
    raise SystemExit('the overflow of the asynchronous render was not reported')
except RuntimeError as e:
    assert 'capacity' in str(e), e
scene2.render_batch([scene2.camera], debug_cameras=[scene2.debug_camera], out=out)
_native.sync()
got = out[0].cpu().numpy()
assert np.abs(got.astype(int) - exp2['rgb'].astype(int)).max() <= 1
print('OVERFLOW-OK')
"""


def test_tile_list_capacity_overflow_grows_and_retries():
    """Per-tile list capacities far too small (B2R_TRI_CAP / B2R_QUAD_CAP = 64 entries per view): the synchronous
    call must grow them and still return the exact frame; an asynchronous call must report the overflow at the next
    sync and be right when rendered again.  Separate process: the capacities are read at b2r_init."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B2R_TRI_CAP="64", B2R_QUAD_CAP="64")
    res = subprocess.run([sys.executable, "-c", _OVERFLOW_SCRIPT.format(root=root)], env=env, capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0 and "OVERFLOW-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_async_pipeline_many_views_matches_single_frames():
    """The throughput paths: 70 views in one call (two set-up batches of <= 64 views, several raster/shade
    sub-chunks on auxiliary streams, cost-ordered tile launch) through (a) the synchronous host path, (b) the
    device-resident path and (c) two host-asynchronous calls in flight -- all byte-identical, and identical to
    frames rendered one at a time."""
    torch = pytest.importorskip("torch")
    from py_numpy_renderer_b200 import _native
    scene = scenes.c3_synthetic((270, 480), tex=128)
    n = 70
    cams = scenes.orbit_cameras(n)
    dcams = scenes.orbit_cameras(n, fovy=90, near=0.05, far=20)
    host = scene.render_batch(cams, debug_cameras=dcams)
    dev = torch.zeros((n, 270, 480, 3), dtype=torch.uint8, device="cuda:0")
    torch.cuda.synchronize()
    scene.render_batch(cams, debug_cameras=dcams, out=dev)
    _native.sync()
    assert np.array_equal(dev.cpu().numpy(), host)
    pinned = [torch.empty((n, 270, 480, 3), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
    f0 = scene.render_batch_async(cams, debug_cameras=dcams, out=pinned[0])
    f1 = scene.render_batch_async(cams[::-1], debug_cameras=dcams[::-1], out=pinned[1])
    f0.result(); f1.result()
    assert np.array_equal(pinned[0], host)
    assert np.array_equal(pinned[1], host[::-1])
    for k in (0, 33, 64, 69):                                  # 64: first view of the second set-up batch
        scene.camera, scene.debug_camera = cams[k], dcams[k]
        scene.persist_silhouette = False
        assert np.array_equal(scene.render(), host[k])


@pytest.mark.gpu
def test_fused_tile_kernel_alternative_stays_bit_exact():
    """B2R_FUSED=1 (shading inside the tile kernel; measured slower, kept as the documented alternative) is selected when
    the library initialises, hence a child process: three fixtures, z / stencil / winner bit-exact, RGB identical to the
    production path's bytes."""
    import subprocess
    import sys
    code = r'''
import sys
sys.path[:0] = [%r, %r + "/tests", %r + "/oracle"]
import numpy as np
import golden_util as gu, oracle as orc
for name in ("g2_diablo_floor_point", "g6_skybox_orthographic", "g12_depth_test_false"):
    scene, exp, meta = gu.load(name)
    scene.persist_silhouette = False
    ref = gu.oracle_frame(orc, scene)
    dbg = {}
    rgb = scene.render(debug=dbg)
    rep = gu.compare_planes(dict(rgb=rgb, z=dbg["z"], stencil=dbg["stencil"], winner=dbg["winner"]), ref)
    assert rep["z_mismatch"] == 0 and rep["stencil_mismatch"] == 0 and rep["winner_mismatch"] == 0 and rep["rgb_px_gt1"] == 0, (name, rep)
    assert np.array_equal(scene.render(), rgb), name
    print("FUSED-OK", name, rep)
''' % (ROOT, ROOT, ROOT)
    res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B2R_FUSED="1"), capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0 and res.stdout.count("FUSED-OK") == 3, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.gpu
def test_production_stencil_path_counts_equal_the_oracle_under_faces():
    """The production stencil path (no debug planes: background pixels skipped, whole (quad, tile) pairs and whole rows
    classified by depth range, per-row difference arrays) normally only shows through the lit bit.  B2R_DEBUG_SKIP_BG=1
    hands out ITS stencil plane: on every covered pixel the count must equal the oracle's, and so must z and the winner.
    B2R_SHADE_F64=1 in the same child: the all-float64 shading kernel stays selectable and within the RGB bar."""
    import subprocess
    import sys
    code = r'''
import sys
sys.path[:0] = [%r, %r + "/tests", %r + "/oracle"]
import numpy as np
import golden_util as gu, oracle as orc
for name in ("g2_diablo_floor_point", "g3_diablo_floor_spot", "g4_diablo_floor_directional", "g8_torus_flat", "g9_diablo_transformed",
             "g7_cube_mtl_rh_directx", "g7_cube_mtl_lh_directx"):
    scene, exp, meta = gu.load(name)
    scene.persist_silhouette = False
    ref = gu.oracle_frame(orc, scene)
    dbg = {}
    rgb = scene.render(debug=dbg)
    cov = ref["winner"] >= 0
    assert np.array_equal(dbg["winner"], ref["winner"]), name
    assert np.array_equal(dbg["z"][cov], ref["z"][cov]), name
    assert np.array_equal(dbg["stencil"][cov], ref["stencil"][cov]), (name, int((dbg["stencil"][cov] != ref["stencil"][cov]).sum()))
    assert (ref["stencil"][cov] != 0).any() or name.startswith("g7"), name
    d = np.abs(rgb.astype(int) - ref["rgb"].astype(int)).max(-1)
    assert (d > 1).sum() == 0 and (d > 0).sum() * 1000 <= d.size, (name, int((d > 0).sum()))
    print("SKIPBG-OK", name, int(cov.sum()), int((ref["stencil"][cov] != 0).sum()))
''' % (ROOT, ROOT, ROOT)
    res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B2R_DEBUG_SKIP_BG="1", B2R_SHADE_F64="1"),
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and res.stdout.count("SKIPBG-OK") == 7, res.stdout[-2000:] + res.stderr[-2000:]
