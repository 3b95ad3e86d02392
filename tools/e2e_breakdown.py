import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/tests']
import numpy as np, torch, scenes
from py_numpy_renderer_b200 import _native
sys.path.insert(0, R)
import bench
_native.init(0)
scene = scenes.c3_synthetic((1080, 1920))
dev = scene._device_scene()
B = 16
pinned = torch.empty((B, 1080, 1920, 3), dtype=torch.uint8, pin_memory=True).numpy()
devout = torch.empty((B, 1080, 1920, 3), dtype=torch.uint8, device="cuda:0")
def T(f, n=20):
    f(); f()
    t = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t) / n * 1e3
cams, dcams = bench.step_cameras(3, 0, 1, B)
for c in cams + dcams: c.scene = scene
bg = scene._background()
print("camera construction   %.2f ms" % T(lambda: bench.step_cameras(3, 0, 1, B)))
def pack():
    for c in cams + dcams:
        c.__dict__.pop('MVP', None); c.__dict__.pop('lookat', None)
    return dev.pack(cams, dcams, scene.light, scene.resolution, scene.system, bg)
print("pack 16 views         %.2f ms" % T(pack))
fp, views = pack()
print("render_packed -> host %.2f ms" % T(lambda: dev.render_packed(fp, views, out=pinned)))
def devr():
    dev.render_packed(fp, views, out=devout); _native.sync()
print("render_packed -> dev  %.2f ms" % T(devr))
h = torch.from_numpy(pinned)
def cp():
    h.copy_(devout, non_blocking=True); torch.cuda.synchronize()
print("D2H 16 frames         %.2f ms" % T(cp))
