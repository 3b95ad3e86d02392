"""Where one plain scene.render() call spends its time (1080p headline scene, one view per call)."""
import contextlib, io, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + '/tests']
import numpy as np, torch, scenes
from py_numpy_renderer_b200 import _native
import bench
_native.init(0)
root = scenes.asset_root()
scene = scenes.kat2(root, (1080, 1920)) if root else scenes.c3_synthetic((1080, 1920))
scene.persist_silhouette = False
dev = scene._device_scene()
def T(f, n=60):
    for _ in range(3): f()
    t = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t) / n * 1e3
cams, dcams = bench.step_cameras(3, 0, 1, 64)
for c in cams + dcams: c.scene = scene
bg = scene._background()
k = [0]
def pack():
    k[0] = (k[0] + 1) % 64
    c, d = cams[k[0]], dcams[k[0]]
    for x in (c, d):
        x.__dict__.pop('MVP', None); x.__dict__.pop('lookat', None)
    return dev.pack([c], [d], scene.light, scene.resolution, scene.system, bg)
print("pack one view             %.3f ms" % T(pack))
fp, views = pack()
devout = torch.empty((1, 1080, 1920, 3), dtype=torch.uint8, device="cuda:0")
pinned = torch.empty((1, 1080, 1920, 3), dtype=torch.uint8, pin_memory=True).numpy()
def devr():
    dev.render_packed(fp, views, out=devout); _native.sync()
print("render_packed -> device   %.3f ms (kernels of one view + sync)" % T(devr))
print("render_packed -> pinned   %.3f ms (+ D2H 6.2 MB)" % T(lambda: dev.render_packed(fp, views, out=pinned)))
_native.set_stage_timing(True); devr(); print("   stage ms", {a: round(b, 3) for a, b in _native.last_stage_ms().items()}); _native.set_stage_timing(False)
def call(verbose):
    scene.verbose = verbose
    def f():
        k[0] = (k[0] + 1) % 64
        scene.camera, scene.debug_camera = cams[k[0]], dcams[k[0]]
        with contextlib.redirect_stdout(io.StringIO()):
            scene.render()
    return f
print("scene.render() verbose    %.3f ms" % T(call(True)))
print("scene.render() quiet      %.3f ms" % T(call(False)))
