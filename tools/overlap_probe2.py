"""Which side slows down when renders and D2H copies run concurrently?  Per-stream CUDA-event timing."""
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
lib = torch.cuda.ExternalStream(_native.stream_ptr())
N = 20
def run(render, copy):
    torch.cuda.synchronize(); _native.sync()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(lib); c0.record(cs)
    for _ in range(N):
        if render: dev.render_packed(fp, views, out=devout)
        if copy:
            with torch.cuda.stream(cs): h.copy_(other, non_blocking=True)
    r1.record(lib); c1.record(cs)
    torch.cuda.synchronize(); _native.sync()
    return r0.elapsed_time(r1) / N, c0.elapsed_time(c1) / N
run(True, True)
print("alone      : render %.2f ms, copy %.2f ms" % (run(True, False)[0], run(False, True)[1]))
print("concurrent : render %.2f ms, copy %.2f ms" % run(True, True))
# copy in 8 pieces (like the library's sub-chunk copies)
def run_pieces():
    torch.cuda.synchronize(); _native.sync()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(lib); c0.record(cs)
    for _ in range(N):
        dev.render_packed(fp, views, out=devout)
        with torch.cuda.stream(cs):
            for i in range(0, B, 2): h[i:i + 2].copy_(other[i:i + 2], non_blocking=True)
    r1.record(lib); c1.record(cs)
    torch.cuda.synchronize(); _native.sync()
    return r0.elapsed_time(r1) / N, c0.elapsed_time(c1) / N
print("concurrent, copy in 8 pieces: render %.2f ms, copy %.2f ms" % run_pieces())
