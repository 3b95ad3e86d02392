"""Small fixed workload for ncu: a few batches of the bench scene (config 3, 1080p) through the public API.
    python tools/profile_step.py [views] [iters] [synthetic|diablo|torus1m]"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/tests']
import numpy as np  # noqa: E402
import torch  # noqa: E402
import scenes  # noqa: E402
from py_numpy_renderer_b200 import _native  # noqa: E402

views = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
workload = sys.argv[3] if len(sys.argv) > 3 else "synthetic"
_native.init(0)
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
_native.set_stage_timing(True)
for it in range(iters):
    start = 0.37 * it
    cams = scenes.orbit_cameras(views, radius=radius, start=start)
    dcams = scenes.orbit_cameras(views, radius=radius, start=start, fovy=90, near=0.05, far=20)
    scene.render_batch(cams, debug_cameras=dcams, out=out)
    _native.sync()
    print(it, {k: round(v, 3) for k, v in _native.last_stage_ms().items()})
print("checksum", int(out.sum().item()))
