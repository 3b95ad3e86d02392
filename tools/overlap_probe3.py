"""Is the slow-down of kernels under a concurrent D2H copy specific to this pipeline?  Other kernels / other copies."""
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
other2 = torch.empty_like(other)
h = torch.empty((B, 1080, 1920, 3), dtype=torch.uint8, pin_memory=True)
cs = torch.cuda.Stream()
lib = torch.cuda.ExternalStream(_native.stream_ptr())
ws = torch.cuda.Stream()
a = torch.randn(4096, 4096, device="cuda:0"); b = torch.randn(4096, 4096, device="cuda:0")
x = torch.randn(64 << 20, device="cuda:0")
N = 20
def work_render(): dev.render_packed(fp, views, out=devout)
def work_mm():
    with torch.cuda.stream(ws): torch.mm(a, b)
def work_elem():
    with torch.cuda.stream(ws): x.mul_(1.0001)
def copy_d2h():
    with torch.cuda.stream(cs): h.copy_(other, non_blocking=True)
def copy_h2d():
    with torch.cuda.stream(cs): other.copy_(h, non_blocking=True)
def copy_d2d():
    with torch.cuda.stream(cs): other2.copy_(other, non_blocking=True)
def run(work, wstream, copy):
    torch.cuda.synchronize(); _native.sync()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(wstream); c0.record(cs)
    for _ in range(N):
        if work: work()
        if copy: copy()
    r1.record(wstream); c1.record(cs)
    torch.cuda.synchronize(); _native.sync()
    return r0.elapsed_time(r1) / N, c0.elapsed_time(c1) / N
for wn, w, wsx in (("render", work_render, lib), ("mm fp32 4096", work_mm, ws), ("elementwise 256MB", work_elem, ws)):
    run(w, wsx, copy_d2h)
    alone = run(w, wsx, None)[0]
    for cn, c in (("D2H", copy_d2h), ("H2D", copy_h2d), ("D2D", copy_d2d)):
        ca = run(None, wsx, c)[1]
        wr, cr = run(w, wsx, c)
        print(f"{wn:18s} alone {alone:6.2f} ms | {cn} alone {ca:6.2f} ms | concurrent: work {wr:6.2f} ms, copy {cr:6.2f} ms")
