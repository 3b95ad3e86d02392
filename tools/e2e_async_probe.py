import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/tests']
import numpy as np, torch, scenes
from py_numpy_renderer_b200 import _native
import bench
_native.init(0)
scene = scenes.c3_synthetic((1080, 1920))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pinned = [torch.empty((B, 1080, 1920, 3), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
t_cam = t_enq = t_wait = 0.0
pend = None
N = 30
for s in range(N + 4):
    if s == 4:
        t_cam = t_enq = t_wait = 0.0; t0 = time.perf_counter()
    a = time.perf_counter()
    cams, dcams = bench.step_cameras(100 + s, 0, 1, B)
    b = time.perf_counter()
    f = scene.render_batch_async(cams, debug_cameras=dcams, out=pinned[s % 2])
    c = time.perf_counter()
    if pend is not None: pend.result()
    d = time.perf_counter()
    pend = f
    t_cam += b - a; t_enq += c - b; t_wait += d - c
pend.result()
tot = time.perf_counter() - t0
print("per step: cameras %.2f ms, pack+enqueue %.2f ms, wait %.2f ms, total %.2f ms -> %.0f fps" % (t_cam / N * 1e3, t_enq / N * 1e3, t_wait / N * 1e3, tot / N * 1e3, N * B / tot))
