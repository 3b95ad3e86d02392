"""Work counters of the raster kernel (needs a -DB2R_STATS build: B2R_STATS=1 python -m py_numpy_renderer_b200.build --force)."""
import ctypes, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/tests']
import numpy as np, torch, scenes
from py_numpy_renderer_b200 import _native
views = int(sys.argv[1]) if len(sys.argv) > 1 else 4
workload = sys.argv[2] if len(sys.argv) > 2 else "synthetic"
lib = _native.init(0)
radius = 3.0
if workload == "diablo":
    scene = scenes.kat2(scenes.asset_root(), (1080, 1920))
elif workload == "torus1m":
    import py_numpy_renderer_b200 as b2r
    cam, dcam = scenes.std_cameras()
    scene = b2r.Scene(cam, scenes.std_light(), debug_camera=dcam, resolution=(1080, 1920), system=b2r.SYSTEM.LH,
                      subsystem=b2r.SUBSYSTEM.OPENGL)
    scene.verbose = False
    scene.add_model(b2r.Model(*scenes.torus_arrays(1000, 500)))
    radius = 2.9
else:
    scene = scenes.c3_synthetic((1080, 1920))
out = torch.empty((views, 1080, 1920, 3), dtype=torch.uint8, device="cuda:0")
buf = (ctypes.c_ulonglong * 16)()
names = ["quad_tile_pairs", "pairs_rejected_by_depth_range", "pairs_with_pixel_work", "stencil_pixel_items",
         "pairs_uniform_counter", "pairs_full_no_span_search", "pairs_all_pass", "tiles_full_winner_pass", "clip_elision_tests", "clip_elision_hits", "tri_pixel_tests_with_clip",
         "tiles", "tri_pixel_tests", "tri_pixel_covered", "quad_list_entries", "tri_list_entries"]
for it in range(2):
    cams = scenes.orbit_cameras(views, radius=radius, start=0.37 * it)
    dcams = scenes.orbit_cameras(views, radius=radius, start=0.37 * it, fovy=90, near=0.05, far=20)
    lib.b2r_debug_stats(buf, 1)
    scene.render_batch(cams, debug_cameras=dcams, out=out)
    _native.sync()
    lib.b2r_debug_stats(buf, 0)
    print("iter", it, "per view:")
    for n, v in zip(names, buf):
        print(f"   {n:28s} {v / views:14.0f}")
