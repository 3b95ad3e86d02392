import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/tests']
import numpy as np, torch, scenes
from py_numpy_renderer_b200 import _native
import bench
_native.init(0)
scene = scenes.c3_synthetic((1080, 1920))
dev = scene._device_scene()
B = 16
cams, dcams = bench.step_cameras(3, 0, 1, B)
for c in cams + dcams: c.scene = scene
fp, views = dev.pack(cams, dcams, scene.light, scene.resolution, scene.system, scene._background())
devout = torch.empty((B, 1080, 1920, 3), dtype=torch.uint8, device="cuda:0")
other = torch.empty((B, 1080, 1920, 3), dtype=torch.uint8, device="cuda:0")
h = torch.empty((B, 1080, 1920, 3), dtype=torch.uint8, pin_memory=True)
cs = torch.cuda.Stream()
def both(n=20):
    torch.cuda.synchronize(); _native.sync()
    t = time.perf_counter()
    for _ in range(n):
        dev.render_packed(fp, views, out=devout)
        with torch.cuda.stream(cs):
            h.copy_(other, non_blocking=True)
    _native.sync(); torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3
def only_render(n=20):
    _native.sync(); t = time.perf_counter()
    for _ in range(n): dev.render_packed(fp, views, out=devout)
    _native.sync(); return (time.perf_counter() - t) / n * 1e3
def only_copy(n=20):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n):
        with torch.cuda.stream(cs): h.copy_(other, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
only_render(3); only_copy(3); both(3)
print("render only %.2f ms | copy only %.2f ms | both concurrently %.2f ms per step" % (only_render(), only_copy(), both()))
