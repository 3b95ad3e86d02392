#!/bin/bash
# round 2, GPU call 8: pinned output pool + overlay fast exit: parity subset, e2e_single, cProfile of scene.render()
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "parity or host or window or integration" 2>&1 | tail -6 > $O/r2c8_pytest.txt
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs > $O/r2c8_bench.json 2> $O/r2c8_bench.err
timeout 300 python - > $O/r2c8_profile_render.txt 2>&1 <<'PY'
import sys, io, contextlib, cProfile, pstats, time
sys.path[:0] = ['.', 'tests']
import numpy as np, scenes
from py_numpy_renderer_b200 import _native
_native.init(0)
sc = scenes.kat2(scenes.asset_root(), (1080, 1920)) if scenes.asset_root() else scenes.c3_synthetic((1080, 1920))
sc.persist_silhouette = False
cams = scenes.orbit_cameras(100); dcams = scenes.orbit_cameras(100, fovy=90, near=0.05, far=20)
for verbose in (True, False):
    sc.verbose = verbose
    with contextlib.redirect_stdout(io.StringIO()):
        for k in range(5):
            sc.camera, sc.debug_camera = cams[k], dcams[k]; sc.render()
        pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
        for k in range(5, 100):
            sc.camera, sc.debug_camera = cams[k], dcams[k]; f = sc.render()
        pr.disable(); dt = time.perf_counter() - t0
    print("verbose", verbose, "ms per render()", 1e3 * dt / 95)
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])
PY
cat $O/r2c8_pytest.txt; head -60 $O/r2c8_profile_render.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c8_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), d.get("e2e_single"), d["e2e"].get("pcie_gbs"))
PY
