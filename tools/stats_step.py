"""Work counters of the raster kernel (needs a -DB2R_STATS build: B2R_STATS=1 python -m py_numpy_renderer_b200.build --force)."""
import ctypes, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/tests']
import numpy as np, torch, scenes
from py_numpy_renderer_b200 import _native
views = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lib = _native.init(0)
scene = scenes.c3_synthetic((1080, 1920))
out = torch.empty((views, 1080, 1920, 3), dtype=torch.uint8, device="cuda:0")
buf = (ctypes.c_ulonglong * 16)()
names = ["pairs", "pairs_rejected_tile_range", "short_rows", "long_rows", "long_rows_bg_only", "long_rows_all_fail",
         "long_rows_all_pass", "long_rows_mixed", "pixel_evals", "pixel_bg_skips", "stencil_updates", "tiles_active",
         "tri_pixel_tests", "tri_pixel_covered", "quad_list_entries", "tri_list_entries"]
for it in range(2):
    cams = scenes.orbit_cameras(views, start=0.37 * it)
    dcams = scenes.orbit_cameras(views, start=0.37 * it, fovy=90, near=0.05, far=20)
    lib.b2r_debug_stats(buf, 1)
    scene.render_batch(cams, debug_cameras=dcams, out=out)
    _native.sync()
    lib.b2r_debug_stats(buf, 0)
    print("iter", it, "per view:")
    for n, v in zip(names, buf):
        print(f"   {n:28s} {v / views:14.0f}")
