"""BASELINE config 5 sanity: 1M-triangle displaced torus (nu=1000, nv=500), orbit cameras at 1080p.
    python tools/c5_check.py [views] [--oracle]   (--oracle: compare view 0 with the CPU oracle)"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/tests', R + '/oracle']
import numpy as np, torch, scenes
import py_numpy_renderer_b200 as b2r
from py_numpy_renderer_b200 import _native
views = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4
t0 = time.time()
v, uv, n, f = scenes.torus_arrays(1000, 500)
cams = scenes.orbit_cameras(views, radius=2.9)
dcams = scenes.orbit_cameras(views, radius=2.9, fovy=90, near=0.05, far=20)
scene = b2r.Scene(cams[0], scenes.std_light(), debug_camera=dcams[0], resolution=(1080, 1920), system=b2r.SYSTEM.LH,
                  subsystem=b2r.SUBSYSTEM.OPENGL)
scene.verbose = False
scene.add_model(b2r.Model(v, uv, n, f))
print("mesh built %.1fs, faces %d" % (time.time() - t0, len(f)))
_native.init(0)
t0 = time.time(); dev = scene._device_scene(); _native.sync(); print("scene upload %.2fs, %.0f MB" % (time.time() - t0, dev.device_bytes / 1e6))
out = torch.empty((views, 1080, 1920, 3), dtype=torch.uint8, device="cuda:0")
_native.set_stage_timing(True)
for it in range(3):
    t0 = time.time()
    scene.render_batch(cams, debug_cameras=dcams, out=out); _native.sync()
    print("iter", it, "%.2f ms/frame" % ((time.time() - t0) * 1e3 / views), {k: round(v, 3) for k, v in _native.last_stage_ms().items()})
if "--oracle" in sys.argv:
    import oracle as orc
    dbg = {}
    rgb = scene.render_batch(cams[:1], debug_cameras=dcams[:1], debug=dbg)
    t0 = time.time(); want = orc.render_scene(scene, cameras=cams[:1]); print("oracle %.1fs" % (time.time() - t0))
    import golden_util as gu
    got = dict(rgb=rgb[0], z=dbg['z'][0], stencil=dbg['stencil'][0], winner=dbg['winner'][0])
    print(gu.compare_planes(got, {k: v[0] for k, v in want.items()}), "status equal:", np.array_equal(dbg['face_status'][0], want['face_status'][0]),
          "n_sil", dbg['n_silhouette'][0], want['n_silhouette'][0])
