#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: frames/s (and Mpix/s) at 1080p for diablo3_pose + floor with shadow
volumes, on N B200s, next to the reference timed on the box's host cores.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 2 --warmup 1        # CPU arm: the NumPy reference itself when staged
                                                                 # (baseline/_ref/ref_src.zip), else the C port

A STEP = one camera-orbit batch of `--views` frames (BASELINE config 3: diablo3_pose, 5022 triangles, 1024^2 diffuse +
tangent-space normal maps; floor with a 600^2 diffuse map; point light; shadow volumes; 1080x1920; LH/OpenGL).
`--workload auto` renders the real assets when they are staged (baseline/_ref/assets, git-ignored, travels with the
snapshot) and a synthetic stand-in of the same sizes otherwise; the line says which.

value : device-resident throughput (scene + pre-evaluated view constants in HBM, frames left in HBM): one CUDA-event
        pair around the K steps, L2 flushed before every step (inside the region), MAX over ranks.
e2e   : through the public API `Scene.render_batch_async(cameras, out=pinned)`: host camera maths + H2D of the view
        constants + render + D2H of the frames inside the timed region (wall clock, MAX over ranks).
e2e_single : a loop of plain `scene.render()` calls (the reference's own entry point), one frame per call.
configs    : (N = 1) the other BASELINE configurations, each with its own frames/s, e2e, algorithmic bytes and
             roofline fraction: config 1 (800x800 diablo), config 4 (4K skybox, perspective + orthographic), config 5
             (1M-triangle torus).
Multi-GPU: `--scaling weak` (default) every rank renders its own `--views` frames per step; `--scaling strong
--frames F` splits a fixed batch of F frames; `--split bands` splits every frame into screen row bands.  Frames are
independent: no data-path collective, one gather (peer window / NCCL) assembles the uint8 frames on rank 0.
"""
from __future__ import annotations

import argparse
import contextlib
import ctypes
import io
import json
import os
import subprocess
import sys
import threading
import time

# --impl reference with several worker processes: one BLAS thread per worker (the thread count is fixed when NumPy loads
# its BLAS, i.e. before the arguments are parsed).  A single worker (--ref-procs 1) keeps the library's own default, which
# is how the reference runs when a user starts it.
if "reference" in sys.argv and not ("--ref-procs" in sys.argv and sys.argv[sys.argv.index("--ref-procs") + 1:][:1] == ["1"]):
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(_v, "1")

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "frames_per_s_1080p_shadow_volumes"
UNIT = "frames/s"
HBM_FALLBACK_GBS = 6650.0
REF_ZIP = os.path.join(ROOT, "baseline", "_ref", "ref_src.zip")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--views", type=int, default=64, help="frames per step per GPU (weak scaling)")
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--workload", default="auto", choices=["auto", "synthetic", "diablo", "torus1m"],
                    help="auto = diablo (the real assets, BASELINE config 3) when staged, else the synthetic stand-in; "
                         "torus1m = BASELINE config 5 (1M-triangle displaced torus, camera orbit)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--frames", type=int, default=256, help="--scaling strong: frames per step over ALL GPUs")
    ap.add_argument("--split", default="frames", choices=["frames", "bands"],
                    help="bands: every rank renders its screen row band of every frame (implies strong scaling)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU sample (default: by core count)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0,
                    help="--impl reference: wall-clock bound of the whole run (steps are seconds each on the CPU)")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "numpy", "port"],
                    help="--impl reference: numpy = the unmodified Python reference (needs baseline/_ref/ref_src.zip or "
                         "/root/reference), one process per core; port = the C restatement, one thread per core")
    ap.add_argument("--ref-procs", type=int, default=0, help="--impl reference: worker processes / threads (0 = all cores)")
    ap.add_argument("--gather", default="auto", choices=["auto", "window", "window-copy", "nccl"],
                    help="N>1: how the frames reach rank 0 -- 'window': every rank's tile kernel stores into rank 0's "
                         "buffer over NVLink (CUDA IPC, parallel.FrameWindow), the only collective is a one-element "
                         "all-reduce per step; 'window-copy': frames are rendered locally and pushed into the window by "
                         "the copy engines while the next step renders; 'nccl': one NCCL gather per step; 'auto': window "
                         "for N <= 2, window-copy beyond")
    ap.add_argument("--push", default="copy", choices=["copy", "sparse"],
                    help="--gather window-copy: 'copy' = whole frames through the copy engines (default), 'sparse' = "
                         "b2r_window_push (tile kernel with peer stores, tiles of constant colour are sent once; measured on "
                         "8 B200: 128.2 k frames/s against 137.4 k with the copy engines, profiles/r02_multi_gpu.md)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the config 1 / 4 / 5 measurements (N = 1 only)")
    ap.add_argument("--no-numpy-ref", action="store_true", help="skip the one-frame NumPy reference timing")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------------
def resolve_workload(name):
    import scenes
    if name == "auto":
        return "diablo" if scenes.asset_root() is not None else "synthetic"
    return name


def build_scene(workload, res):
    """-> (scene, description, orbit radius)"""
    import scenes
    if workload == "diablo":
        assets = scenes.asset_root()
        if assets is None:
            raise SystemExit("--workload diablo needs the reference assets (baseline/_ref/assets)")
        return (scenes.kat2(assets, res), "diablo3_pose.obj (5022 tris, 1024^2 diffuse + tangent normal maps) + floor "
                "(600^2 diffuse), point light, shadow volumes [reference assets]", 3.0)
    if workload == "torus1m":
        import py_numpy_renderer_b200 as b2r
        v, uv, n, f = scenes.torus_arrays(1000, 500)
        cam, dcam = scenes.std_cameras()
        sc = b2r.Scene(cam, scenes.std_light(), debug_camera=dcam, resolution=res, system=b2r.SYSTEM.LH,
                       subsystem=b2r.SUBSYSTEM.OPENGL)
        sc.verbose = False
        sc.add_model(b2r.Model(v, uv, n, f))
        return sc, "synthetic 1M-triangle displaced torus (nu=1000, nv=500), camera orbit radius 2.9, shadow volumes", 2.9
    return (scenes.c3_synthetic(res), "synthetic diablo3-class figure (5000 tris, diffuse + tangent normal map 1024^2) "
            "+ floor (2 tris, 600^2 diffuse), point light, shadow volumes [assets not staged]", 3.0)


def step_cameras(step, rank, world, views, radius=3.0):
    """Orbit cameras of one step: every (step, rank, view) gets its own angle; the debug frustum contains the
    camera frustum."""
    import scenes
    start = 2 * np.pi * ((step * world + rank) * 0.6180339887498949 % 1.0)
    return (scenes.orbit_cameras(views, radius=radius, start=start),
            scenes.orbit_cameras(views, radius=radius, start=start, fovy=90, near=0.05, far=20))


def algorithmic_bytes(scene, n_shaded, n_background=0):
    """B_alg = B_geom + B_tex + B_fb per frame, in the reference's own storage; every model's maps
    counted for that model's shaded pixels only is not knowable without a per-model count, so the bound used is
    min(sum of the map sizes, shaded pixels x 3 bytes x maps of the model with the most maps)."""
    from py_numpy_renderer_b200.cube_map import CubeMap
    b_geom, map_bytes, maps_per_model = 0, 0, 0
    for m in scene.models:
        v = np.asarray(m.vertices)
        b_geom += v.shape[0] * v.dtype.itemsize * 4
        b_geom += (0 if m.uv is None else len(m.uv) * 12) + (0 if m.normals is None else len(m.normals) * 12)
        b_geom += len(m._faces) * 48
        n_maps = 0
        for mat in m.materials.values():
            for attr in ("map_Kd", "map_Ks", "norm"):
                t = getattr(mat, attr, None)
                if t is not None:
                    map_bytes += t.texels.shape[0] * t.texels.shape[1] * 3
                    n_maps += 1
        maps_per_model = max(maps_per_model, n_maps)
    b_tex = min(map_bytes, n_shaded * 3 * maps_per_model)
    if isinstance(scene.skybox, CubeMap):
        S = scene.skybox.texels.shape[1]
        b_tex += min(6 * S * S * 3, n_background * 3)
    b_fb = scene.resolution[0] * scene.resolution[1] * 3
    return dict(geom=int(b_geom), tex=int(b_tex), fb=int(b_fb), total=int(b_geom + b_tex + b_fb))


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (NVML, 50 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arms (test infrastructure under oracle/ is only ever executed here)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_port_sample(scene, cams, dcams, threads):
    """The C restatement of the reference (oracle port) on the host cores: one frame per thread.  -> (frames/s, s)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.build()
    from py_numpy_renderer_b200 import _abi
    from py_numpy_renderer_b200.cube_map import CubeMap
    sky = scene.skybox if isinstance(scene.skybox, CubeMap) else None
    packed = _abi.PackedScene(scene.models, sky)
    fp = _abi.pack_frame_params(scene.light, scene.resolution, scene._background(), False)
    for c in list(cams) + list(dcams):
        c.scene = scene
    n = len(cams)
    views = (_abi.View * n)(*[_abi.pack_view(c, d, scene.system, sky is not None) for c, d in zip(cams, dcams)])
    t0 = time.perf_counter()
    orc.render_packed(packed, fp, views, threads=threads, planes=False)
    dt = time.perf_counter() - t0
    return n / dt, dt


def numpy_reference_available():
    return os.path.isfile(REF_ZIP) or os.path.isfile("/root/reference/obj/core.py")


def _numpy_worker(workload, res, frame_ids, n_total, radius, barrier, queue):
    """One worker process of the NumPy-reference arm: boots the unmodified reference, converts the product-side scene
    description into the reference's own objects (fresh Models per frame: model.silhouette persists, SURVEY B-3) and
    times scene.render() per frame."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import refboot
        refboot.boot()
        scene, _, _ = build_scene(workload, res)
        import scenes
        cams = scenes.orbit_cameras(n_total, radius=radius, start=0.123)
        dcams = scenes.orbit_cameras(n_total, radius=radius, start=0.123, fovy=90, near=0.05, far=20)
        jobs = [refboot.to_reference_scene(scene, cams[i], dcams[i]) for i in frame_ids]
        if barrier is not None:
            barrier.wait()
        times = []
        for rs in jobs:
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                rs.render()
            times.append(time.perf_counter() - t0)
        queue.put(times)
    except BaseException as exc:  # noqa: BLE001 -- the parent must not wait for ever
        if barrier is not None:
            barrier.abort()
        queue.put(f"{type(exc).__name__}: {exc}")


def numpy_reference_sample(workload, res, n_frames, procs, radius):
    """The UNMODIFIED NumPy reference: `n_frames` orbit frames over `procs` processes (the reference is single
    threaded: P processes render P different frames).  -> (frames/s, wall s, per-frame seconds)"""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    procs = max(1, min(procs, n_frames))
    shards = [list(range(p, n_frames, procs)) for p in range(procs)]
    barrier, queue = ctx.Barrier(procs + 1), ctx.Queue()
    workers = [ctx.Process(target=_numpy_worker, args=(workload, res, s, n_frames, radius, barrier, queue)) for s in shards]
    for w in workers:
        w.start()
    try:
        barrier.wait(timeout=600)          # every worker has booted the reference and built its scenes
    except Exception:  # noqa: BLE001
        pass
    t0 = time.perf_counter()
    out = [queue.get(timeout=3600) for _ in workers]
    wall = time.perf_counter() - t0
    for w in workers:
        w.join()
    bad = [o for o in out if isinstance(o, str)]
    if bad:
        raise RuntimeError("NumPy reference worker failed: " + bad[0])
    times = [t for part in out for t in part]
    return n_frames / wall, wall, times


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    workload = resolve_workload(args.workload)
    res = (args.height, args.width)
    scene, desc, radius = build_scene(workload, res)
    cores = os.cpu_count() or 1
    procs = args.ref_procs or cores
    kind = args.ref_kind
    if kind == "auto":
        kind = "numpy" if numpy_reference_available() else "port"
    if kind == "numpy" and not numpy_reference_available():
        raise SystemExit("--ref-kind numpy needs baseline/_ref/ref_src.zip (python -c 'import __graft_entry__ as g; g.build()')")
    n_frames = args.cpu_frames or max(4, min(procs, 128))   # one frame per worker: every core works
    t_start = time.perf_counter()
    total_t, total_f, done = 0.0, 0, 0
    per_frame = []
    import scenes
    for s in range(args.steps + max(0, min(args.warmup, 1 if kind == "port" else 0))):
        if kind == "numpy":
            _, dt, times = numpy_reference_sample(workload, res, n_frames, procs, radius)
            per_frame += times
        else:
            cams = scenes.orbit_cameras(n_frames, radius=radius, start=0.123 + s)
            dcams = scenes.orbit_cameras(n_frames, radius=radius, start=0.123 + s, fovy=90, near=0.05, far=20)
            _, dt = cpu_port_sample(scene, cams, dcams, procs)
            if kind == "port" and s == 0 and args.warmup > 0:
                continue                                    # first pass pages the library and the textures in
        total_t += dt
        total_f += n_frames
        done += 1
        if time.perf_counter() - t_start + dt > args.cpu_budget_s:
            break
    fps = total_f / total_t
    H, W = res
    base = {"value": fps, "unit": UNIT, "cores": cores,
            "kind": "reference" if kind == "numpy" else "port"}
    if kind == "numpy":
        try:   # the C restatement on the same cores, for comparison (one short pass)
            cams = scenes.orbit_cameras(n_frames, radius=radius, start=0.123)
            dcams = scenes.orbit_cameras(n_frames, radius=radius, start=0.123, fovy=90, near=0.05, far=20)
            cpu_port_sample(scene, cams[:1], dcams[:1], 1)
            base["port_fps_same_cores"] = cpu_port_sample(scene, cams, dcams, procs)[0]
        except Exception as exc:  # noqa: BLE001
            base["port_fps_same_cores"] = f"unavailable: {exc}"[:120]
        base["blas_threads_per_process"] = os.environ.get("OPENBLAS_NUM_THREADS", "library default")
        base["sample"] = (f"{n_frames} orbit frames per step over {min(procs, n_frames)} processes (one frame each, "
                          f"fresh Models per frame), the UNMODIFIED NumPy reference (Scene.render, obj/core.py:587-640) "
                          f"imported from baseline/_ref/ref_src.zip; median single-frame time "
                          f"{float(np.median(per_frame)):.2f} s")
        base["single_process_fps"] = 1.0 / float(np.median(per_frame))
    else:
        base["sample"] = (f"{n_frames} orbit frames per step, {min(procs, n_frames)} host threads (one frame each), "
                          f"C restatement of Scene.render() pinned to the Python reference")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / done,
            "higher_is_better": True, "scaling": args.scaling if args.split == "frames" else "strong",
            "vs_baseline": None, "dtype": "f64",
            "data": "reference assets" if workload == "diablo" else "synthetic",
            "mpix_per_s": fps * H * W / 1e6,
            "config": {"workload": desc, "resolution": [H, W], "frames_per_step_per_gpu": args.views,
                       "split": args.split, "cpu_frames_per_timed_step": n_frames,
                       "note": "same scene, resolution and orbit as the B200 arm; a CPU step is a bounded sample of "
                               "cpu_frames_per_timed_step frames (one per worker), rates are per frame"},
            "cpu_baseline": base,
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def numpy_reference_one_frame(workload, res):
    """cpu_baseline.numpy_reference of the B200 arm: ONE frame of the unmodified NumPy reference, single process, in a
    child interpreter (this process holds a CUDA context)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--ref-kind", "numpy", "--ref-procs", "1",
           "--cpu-frames", "1", "--steps", "1", "--warmup", "0", "--workload", workload,
           "--height", str(res[0]), "--width", str(res[1])]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
        return {"value": d["value"], "unit": UNIT, "processes": 1, "host_cores": os.cpu_count(),
                "seconds_per_frame": 1.0 / d["value"],
                "what": "one frame of the unmodified NumPy reference on this box (scene.render(), fresh Models)"}
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}


# ---------------------------------------------------------------------------------------------------------------------
# single-GPU measurement of one scene (used for the other BASELINE configurations)
# ---------------------------------------------------------------------------------------------------------------------
def measure_scene(scene, cam_fn, B, K, Wm, torch, _native, lib_stream, flush, peak, e2e_steps=4, n_shaded_cb=None):
    """Device-resident frames/s (CUDA events, L2 flushed between steps), stage times, end-to-end frames/s through
    render_batch_async, algorithmic bytes and the HBM roofline fraction of the dominant stage."""
    H, W = scene.resolution
    device = flush.device
    dev = scene._device_scene()
    bg = scene._background()
    cur = torch.cuda.current_stream()
    dbg = {}
    cams0, dcams0 = cam_fn(0, 1)
    scene.render_batch(cams0, debug_cameras=dcams0, debug=dbg)
    n_shaded = int((dbg["winner"][0] >= 0).sum())
    balg = algorithmic_bytes(scene, n_shaded, H * W - n_shaded)
    packed = []
    for s in range(Wm + K):
        cams, dcams = cam_fn(s, B)
        for c in cams + dcams:
            c.scene = scene
        packed.append(dev.pack(cams, dcams, scene.light, scene.resolution, scene.system, bg))
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=device)

    def run(first, count, stage_acc=None):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(cur)
        for s in range(first, first + count):
            flush.zero_()
            ready = torch.cuda.Event()
            ready.record(cur)
            lib_stream.wait_event(ready)
            dev.render_packed(*packed[s], out=out)
            done = torch.cuda.Event()
            done.record(lib_stream)
            cur.wait_event(done)
            if stage_acc is not None:
                torch.cuda.synchronize()
                _native.sync()
                for k, v in _native.last_stage_ms().items():
                    stage_acc[k] = stage_acc.get(k, 0.0) + v
        b.record(cur)
        torch.cuda.synchronize()
        _native.sync()
        return a.elapsed_time(b)

    run(0, Wm)
    l0 = _native.launch_count()
    ms = run(Wm, K)
    launches = _native.launch_count() - l0
    fps = K * B / (ms / 1e3)
    stage = {}
    n_st = min(K, 3)
    _native.set_stage_timing(True)
    run(Wm + K - n_st, n_st, stage)
    _native.set_stage_timing(False)
    stage = {k: v / n_st for k, v in stage.items()}
    # end to end
    pinned = [torch.empty((B, H, W, 3), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
    for w in range(2):   # warm-up of BOTH staging slots of the asynchronous path (their first use allocates device memory)
        cams, dcams = cam_fn(1000 + w, B)
        scene.render_batch_async(cams, debug_cameras=dcams, out=pinned[w]).result()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fl = None
    for s in range(e2e_steps):
        cams, dcams = cam_fn(2000 + s, B)
        fut = scene.render_batch_async(cams, debug_cameras=dcams, out=pinned[s % 2])
        if fl is not None:
            fl.result()
        fl = fut
    fl.result()
    e2e = e2e_steps * B / (time.perf_counter() - t0)
    top = max(stage, key=stage.get)
    achieved = balg["total"] * B / (stage[top] / 1e3) / 1e9
    del pinned, out
    return {"value": fps, "unit": UNIT, "mpix_per_s": fps * H * W / 1e6, "ms_per_step": ms / K, "frames_per_step": B,
            "steps": K, "gpu_launches": int(launches), "resolution": [H, W],
            "e2e": {"value": e2e, "unit": UNIT, "d2h_bytes_per_step": B * H * W * 3, "steps": e2e_steps},
            "stage_ms_per_step": stage, "shaded_pixels_per_frame": n_shaded,
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "algorithmic_bytes_per_frame": balg,
                         "whole_frame_frac": balg["total"] * fps / 1e9 / peak}}


def config_scenes():
    """BASELINE configs 1, 4 (perspective / orthographic) and 5 as (key, builder) pairs.  Builders return
    (scene, description, camera function (step, n) -> (cameras, debug cameras), frames per step, steps)."""
    import py_numpy_renderer_b200 as b2r
    import scenes
    assets = scenes.asset_root()

    def orbit(radius, **kw):
        def fn(step, n):
            start = 2 * np.pi * (step * 0.6180339887498949 % 1.0)
            return (scenes.orbit_cameras(n, radius=radius, start=start, **kw),
                    scenes.orbit_cameras(n, radius=radius, start=start, **dict(kw, fovy=90, near=0.05, far=20)))
        return fn

    def c1():
        if assets:
            sc, what = scenes.kat1(assets, (800, 800)), "diablo3_pose.obj alone, untextured [reference assets]"
        else:
            sc = scenes.c3_synthetic((800, 800))
            sc.models[:] = sc.models[:1]
            sc._invalidate_device()
            what = "synthetic 5000-triangle figure alone"
        return sc, what + ", point light, 800x800, shadow pass on (the reference cannot switch it off)", orbit(3.0), 64, 20

    def c4(ortho):
        def build():
            res = (2160, 3840)
            proj = dict(projection_type=b2r.PROJECTION_TYPE.ORTHOGRAPHIC) if ortho else {}
            if assets and os.path.isdir(os.path.join(assets, "skybox")):
                sky = b2r.CubeMap(**{k: os.path.join(assets, "skybox", k + ".jpg")
                                     for k in ("left", "right", "top", "bottom", "front", "back")})
                what = "diablo3_pose + 2048^2 skybox [reference assets]"
            else:
                import golden_util as gu
                tex = np.stack([scenes.procedural_texture(2048, 40 + i) for i in range(6)])
                sky = gu.SkyFromTexels(tex)
                what = "figure + procedural 2048^2 cubemap"
            model = (b2r.Model.load_model(os.path.join(assets, "diablo3_pose/diablo3_pose.obj")) if assets
                     else scenes.figure_model())
            if assets:
                model.textures.register('diffuse', os.path.join(assets, 'diablo3_pose/diablo3_pose_diffuse.tga'),
                                        normalize=False)
            model = model @ b2r.translation((0, 0, -2.))
            center = np.array((0, 0, -2.))

            def fn(step, n):
                cams, dcams = [], []
                for k in range(n):
                    pos = (0.3 + 0.02 * ((step * n + k) % 16), 0.5, 1.5)
                    cams.append(b2r.Camera(pos, center=center, fovy=60, near=0.1, far=10, backface_culling=True, **proj))
                    dcams.append(b2r.Camera(pos, center=center, fovy=90, near=0.05, far=20, backface_culling=True, **proj))
                return cams, dcams
            cam, dcam = fn(0, 1)
            sc = b2r.Scene(cam[0], scenes.std_light(), debug_camera=dcam[0], resolution=res, system=b2r.SYSTEM.LH,
                           subsystem=b2r.SUBSYSTEM.OPENGL, skymap=sky)
            sc.verbose = False
            sc.add_model(model)
            return (sc, what + (", ORTHOGRAPHIC" if ortho else ", PERSPECTIVE") + " camera, 2160x3840 (reflective "
                    "materials do not exist in the reference: not rendered)", fn, 8, 10)
        return build

    def c5():
        sc, what, radius = build_scene("torus1m", (1080, 1920))
        return sc, what, orbit(radius), 64, 4   # a quarter of BASELINE's 256-frame orbit per step

    return [("config1_800x800", c1), ("config4_4k_skybox_perspective", c4(False)),
            ("config4_4k_skybox_orthographic", c4(True)), ("config5_torus_1m_triangles", c5)]


# ---------------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    from py_numpy_renderer_b200 import _native, parallel
    cpus_before = len(os.sched_getaffinity(0))
    numa_cpus = _native.bind_host_to_gpu(local)             # pinned buffers on the GPU's NUMA node (before torch threads)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    _native.init(local)
    parallel.init_process_group("nccl" if world > 1 else None)
    device = torch.device("cuda", local)

    workload = resolve_workload(args.workload)
    scene, desc, radius = build_scene(workload, (args.height, args.width))
    H, W = scene.resolution
    K, Wm = args.steps, args.warmup
    bands = None
    if args.split == "bands":
        B = args.views
        bands = parallel.shard_bands(H, world)
        scaling = "strong"
    elif args.scaling == "strong":
        if args.frames % world:
            raise SystemExit("--frames must be divisible by the number of GPUs")
        B = args.frames // world
        scaling = "strong"
    else:
        B = args.views
        scaling = "weak"
    frames_per_step_total = B if bands is not None else B * world
    dev = scene._device_scene()
    lib_stream = torch.cuda.ExternalStream(_native.stream_ptr(), device=device)
    bg = scene._background()

    # one debug render: shaded-pixel count for the algorithmic-bytes model, silhouette size of the workload
    dbg = {}
    cams0, dcams0 = step_cameras(0, rank, world, 1, radius)
    scene.render_batch(cams0, debug_cameras=dcams0, debug=dbg)
    n_shaded = int((dbg["winner"][0] >= 0).sum())
    n_sil = [int(x) for x in dbg["n_silhouette"][0]]
    n_lit = int(((dbg["winner"][0] >= 0) & (dbg["stencil"][0] == 0)).sum())
    balg = algorithmic_bytes(scene, n_shaded, H * W - n_shaded)

    def cams_of(step, r):
        """cameras of (step, rank r); band split: every rank renders the same frames"""
        return step_cameras(step, 0 if bands is not None else r, 1 if bands is not None else world, B, radius)

    # pre-evaluate the view constants of every step (host maths is part of e2e, not of `value`)
    packed_steps = []
    for s in range(Wm + K):
        cams, dcams = cams_of(s, rank)
        for c in cams + dcams:
            c.scene = scene
        packed_steps.append(dev.pack(cams, dcams, scene.light, scene.resolution, scene.system, bg,
                                     band=bands[rank] if bands is not None else None))
    frames_dev = [torch.zeros((B, H, W, 3), dtype=torch.uint8, device=device) for _ in range(2)]
    if args.gather == "auto":
        args.gather = "window" if world <= 2 else "window-copy"
    if bands is not None:
        args.gather = "nccl"
    win, gather_mode = None, ("none" if world == 1 else args.gather)
    if world > 1 and args.gather.startswith("window"):
        try:
            win = parallel.FrameWindow(B, H, W, slots=2, dst=0)
        except Exception as exc:                              # no IPC between these processes: fall back, and say so
            print(f"[bench] rank {rank}: output window unavailable ({exc}); using the NCCL gather", file=sys.stderr)
        ok = torch.tensor([1 if win is not None else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            win, gather_mode = None, "nccl (window unavailable)"
    push = win is not None and args.gather == "window-copy"
    push_stream = torch.cuda.Stream(priority=0) if push else None   # default (lowest) priority: the renders come first
    win_blocks = [torch.as_tensor(win.block(s), device=device) for s in range(2)] if push else None
    gathered = [[torch.empty((B, H, W, 3), dtype=torch.uint8, device=device) for _ in range(world)]
                for _ in range(2)] if (world > 1 and rank == 0 and win is None and bands is None) else [None, None]
    band_full = [None, None]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    cur = torch.cuda.current_stream()

    def run_steps(first, count, per_step_sync):
        """`count` steps as a stream pipeline: [L2 flush] -> render (library stream) -> the step's frames reach rank 0
        (peer window / NCCL), that transfer overlapping the render of step s+1 (two frame buffers).  Returns elapsed
        device milliseconds between the first flush and the last transfer."""
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        works = [None, None]
        ev_a.record(cur)
        for s in range(first, first + count):
            fp, views = packed_steps[s]
            slot = s % 2
            flush.zero_()                                   # evict L2 BETWEEN steps: 256 MiB > 126 MB, timed (~45 us)
            if works[slot] is not None:
                works[slot].wait()                          # frames_dev[slot] was handed to NCCL two steps ago
                works[slot] = None
            ready = torch.cuda.Event()
            ready.record(cur)
            lib_stream.wait_event(ready)
            # asynchronous, frames stay in HBM: this rank's buffer, or its block of rank 0's window (peer stores)
            dev.render_packed(fp, views, out=frames_dev[slot] if (win is None or push) else win.block_ptr(slot))
            done = torch.cuda.Event()
            done.record(lib_stream)
            cur.wait_event(done)                            # the next flush must not start before this render ended
            if push:                                        # copy engines move the frames while the next step renders
                push_stream.wait_event(done)
                with torch.cuda.stream(push_stream):
                    if args.push == "sparse":
                        win.push(slot, frames_dev[slot], push_stream)   # tiles of constant colour cross NVLink once
                    else:
                        win_blocks[slot].copy_(frames_dev[slot], non_blocking=True)
                    works[slot] = win.fence(async_op=True)
            elif win is not None:
                works[slot] = win.fence(async_op=True)      # one-element all-reduce: the step's frames are on rank 0
            elif bands is not None and world > 1:
                band_full[slot] = parallel.gather_bands(frames_dev[slot], bands, H, dst=0)
            elif world > 1:
                works[slot] = dist.gather(frames_dev[slot], gathered[slot], dst=0, async_op=True)
            if per_step_sync:
                torch.cuda.synchronize()
                _native.sync()
                for k, v in _native.last_stage_ms().items():
                    stage_ms[k] = stage_ms.get(k, 0.0) + v
        for w in works:
            if w is not None:
                w.wait()
        fin = torch.cuda.Event()
        fin.record(lib_stream)
        cur.wait_event(fin)
        ev_b.record(cur)
        torch.cuda.synchronize()
        _native.sync()
        return ev_a.elapsed_time(ev_b)

    stage_ms = {}
    run_steps(0, Wm, False)                                 # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _native.launch_count()
    wall0 = time.perf_counter()
    dev_ms = run_steps(Wm, K, False)                        # the timed region: exactly K steps
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = _native.launch_count() - launches0
    t = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = K * frames_per_step_total / (dev_ms / 1e3)

    # ---- exactness of the assembled frames on real GPUs (outside the timed region): rank 0 re-renders a sample of
    # EVERY rank's views of the last step itself and compares bytes with what arrived ----
    byte_check = None
    if world > 1 and rank == 0:
        last, slot = Wm + K - 1, (Wm + K - 1) % 2
        if win is not None:
            arrived = torch.as_tensor(win.frames(slot), device=device).view(world, B, H, W, 3)
        elif bands is not None:
            arrived = band_full[slot].view(1, B, H, W, 3)
        else:
            arrived = torch.stack(gathered[slot])
        sample = sorted({0, B // 2, B - 1})
        tmp = torch.empty((len(sample), H, W, 3), dtype=torch.uint8, device=device)
        bad, checked = 0, 0
        for r in range(arrived.shape[0]):
            cams, dcams = cams_of(last, r)
            scene.render_batch([cams[i] for i in sample], debug_cameras=[dcams[i] for i in sample], out=tmp)
            _native.sync()
            torch.cuda.synchronize()
            for j, i in enumerate(sample):
                bad += int((arrived[r, i] != tmp[j]).any())
                checked += 1
                assert int(tmp[j].max()) > int(tmp[j].min()), "re-rendered frame is constant"
        byte_check = {"frames_checked": checked, "frames_differing": bad,
                      "what": f"rank 0 re-rendered views {sample} of every rank's last step and compared bytes"}
    if world > 1:
        dist.barrier()                   # nobody overwrites a window slot (stage pass below) while rank 0 still compares
        assert byte_check is None or byte_check["frames_differing"] == 0, byte_check

    # per-stage CUDA-event timing (library stream) on a few more steps of the same workload, synchronised per step
    n_stage_steps = min(K, 5)
    _native.set_stage_timing(True)
    run_steps(Wm + K - n_stage_steps, n_stage_steps, True)
    _native.set_stage_timing(False)

    # ---- PCIe ceiling of this box: all ranks copy a step's frames device -> pinned host at the same time ----
    pinned = [torch.empty((B, H, W, 3), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    for _ in range(2):
        pinned[0].copy_(frames_dev[0], non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n_copies = 6
    for i in range(n_copies):
        pinned[i % 2].copy_(frames_dev[i % 2], non_blocking=True)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pcie_gbs = n_copies * B * H * W * 3 * world / float(t.item()) / 1e9       # aggregate over the ranks
    pinned = [p.numpy() for p in pinned]

    # ---- end to end through the public API: host camera maths + H2D + render + D2H into pinned memory ----
    e2e_steps = max(3, min(K, 20))
    band = bands[rank] if bands is not None else None
    for s in range(2):
        cams, dcams = cams_of(1000 + s, rank)
        scene.render_batch(cams, debug_cameras=dcams, out=pinned[s], band=band)
    # three repetitions of e2e_steps steps, the median is reported (wall clock on a shared host: a single short run
    # swings by tens of percent with whatever else uses the host's memory / PCIe fabric at that moment)
    e2e_runs = []
    for rep in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        in_flight = None
        for s in range(e2e_steps):
            # camera construction + matrix evaluation of step s happen here while step s-1 renders / copies
            cams, dcams = cams_of(2000 + 100 * rep + s, rank)
            fut = scene.render_batch_async(cams, debug_cameras=dcams, out=pinned[s % 2], band=band)
            if in_flight is not None:
                in_flight.result()                                      # frames of step s-1 are in host memory
            in_flight = fut
        in_flight.result()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_runs.append(e2e_steps * frames_per_step_total / float(t.item()))
    e2e_fps = sorted(e2e_runs)[1]
    d2h_step = B * H * W * 3 if bands is None else B * (bands[rank][1] - bands[rank][0]) * W * 3

    # ---- plain scene.render() calls, one frame per call (the reference's own entry point) ----
    e2e_single = None
    if world == 1:
        scene.persist_silhouette = False    # every call does the full shadow work (the reference's persistent set makes
        single = {}                         # every second call shadow-free, SURVEY B-3)
        n_single = 60
        cams, dcams = step_cameras(3000, 0, 1, n_single, radius)
        for verbose in (True, False):
            scene.verbose = verbose
            with contextlib.redirect_stdout(io.StringIO()):
                for k in range(3):
                    scene.camera, scene.debug_camera = cams[k], dcams[k]
                    scene.render()
                rates = []
                for rep in range(3):   # 60 calls are ~50 ms of wall clock on a shared host: the median of three passes
                    t0 = time.perf_counter()
                    for k in range(n_single):
                        scene.camera, scene.debug_camera = cams[k], dcams[k]
                        frame = scene.render()
                    rates.append(n_single / (time.perf_counter() - t0))
            single["verbose_default" if verbose else "verbose_off"] = sorted(rates)[1]
        scene.verbose = False
        e2e_single = {"value": single["verbose_default"], "unit": UNIT, "verbose_off": single["verbose_off"],
                      "frames": n_single, "reported": "median of three passes of 60 calls", "d2h_bytes_per_frame": int(frame.nbytes),
                      "api": "scene.render() -> new uint8 (H, W, 3) ndarray per call (core.py:587-640 semantics incl. the "
                             "three status lines per model when verbose, its default), one camera per call, blocking"}
    sampler.stop_flag = True
    sampler.join(timeout=1)

    # ---- the other BASELINE configurations (single GPU) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
    configs = None
    if world == 1 and not args.no_configs:
        configs = {}
        scene._invalidate_device()
        del frames_dev, pinned
        torch.cuda.empty_cache()
        for key, build in config_scenes():
            try:
                sc, what, cam_fn, Bc, Kc = build()
                res = measure_scene(sc, cam_fn, Bc, Kc, 3, torch, _native, lib_stream, flush, peak)
                res["workload"] = what
                if not args.no_cpu_baseline:                # one frame of the C port (seconds)
                    c, d = cam_fn(0, 1)
                    fps, dt = cpu_port_sample(sc, c, d, 1)
                    res["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": 1, "kind": "port",
                                           "sample": f"one frame, one thread, {dt:.1f} s"}
                sc._invalidate_device()
                configs[key] = res
            except Exception as exc:  # noqa: BLE001 -- a failing side measurement must not lose the headline line
                configs[key] = {"failed": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    from py_numpy_renderer_b200 import _abi
    top = max(stage_ms, key=stage_ms.get) if stage_ms else None
    stage_avg = {k: v / n_stage_steps for k, v in stage_ms.items()}
    roofline = None
    if top:
        achieved = balg["total"] * B / (stage_avg[top] / 1e3) / 1e9
        traffic, ceilings = None, {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))   # ncu dram bytes of one launch of `frames_per_launch` frames: scale to B
            if tj.get(top) is not None:
                traffic = int(tj[top] * B / tj.get("frames_per_launch", B))
            ceilings = tj.get("stage_ceilings", {})
        # every stage against ITS OWN algorithmic bytes (what the stage must read and write at the least) and, where ncu
        # measured one, against the ceiling that actually binds it (issue slots / launch latency), profiles/traffic.json
        F = sum(len(m._faces) for m in scene.models)
        S = sum(n_sil)
        own = {"silhouette": balg["geom"] // max(B, 1) + S * 128, "tri_setup": balg["geom"] + F * 128,
               "quad_setup": S * (128 + 256), "bin": F * 16 + S * 256, "tile": F * 128 + S * 256 + balg["tex"] + balg["fb"],
               "raster": F * 128 + S * 256 + H * W * 6, "shade": H * W * 6 + balg["tex"] + balg["fb"]}
        stages = {}
        for k, v in stage_avg.items():
            if v <= 0:
                continue
            gbs = own.get(k, balg["total"]) * B / (v / 1e3) / 1e9
            stages[k] = {"ms_per_step": v, "algorithmic_bytes_per_frame": int(own.get(k, balg["total"])),
                         "hbm_achieved_gbs": gbs, "hbm_frac": gbs / peak}
            if k in ceilings:
                stages[k].update(ceilings[k])
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_frame": balg, "frames_per_launch": B,
                    "kernel_ms_per_launch": stage_avg[top], "stage_ms_per_step": stage_avg,
                    "whole_frame_frac": balg["total"] * value / world / 1e9 / peak,
                    "stages": stages,
                    "note": "HBM is the roofline the contract names; ncu shows the tile kernel bound by issue slots "
                            "(float64 edge / depth arithmetic the reference's semantics prescribe), see "
                            "profiles/README.md and the per-stage `ceiling` entries"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "reference assets" if workload == "diablo" else "synthetic",
            "mpix_per_s": value * H * W / 1e6,
            "config": {"workload": desc, "resolution": [H, W], "frames_per_step_per_gpu": B,
                       "frames_per_step": frames_per_step_total, "split": args.split,
                       "workload_stats": {"faces": sum(len(m._faces) for m in scene.models), "silhouette_edges": n_sil,
                                          "shaded_pixels": n_shaded, "lit_pixels": n_lit},
                       "parallelism": (f"{'row bands' if bands is not None else 'frames'} x{world} ({scaling}); frames reach "
                                       f"rank 0 by: " +
                                       ("peer stores of the tile kernel into rank 0's window over NVLink + a "
                                        "one-element all-reduce per step" if gather_mode == "window" else
                                        ("b2r_window_push (a tile kernel with peer stores over NVLink: tiles of constant "
                                         "colour are sent once)" if args.push == "sparse" else "copy-engine pushes") +
                                        " into rank 0's window over NVLink, overlapped with the next "
                                        "render, + a one-element all-reduce per step" if gather_mode == "window-copy" else
                                        f"{gather_mode} gather per step" + ("" if bands is not None else
                                                                            ", overlapped with the next render")))
                       if world > 1 else "single GPU",
                       "arithmetic": "float64 coverage / depth / stencil / perspective weights / texel addressing (bit-exact "
                                     "against the reference), float32 lighting sum (B2R_SHADE_F64=1 selects the all-float64 "
                                     "shading kernel)" if not os.environ.get("B2R_SHADE_F64") else "float64 throughout",
                       "l2": "flushed before every step (256 MiB memset, inside the timed region)",
                       "timing": "one CUDA-event pair around the K steps (first flush .. last transfer), MAX over ranks; "
                                 "stage times from a separate per-step-synchronised pass",
                       "wall_ms_per_step": 1e3 * wall / K},
            "e2e": {"value": e2e_fps, "unit": UNIT,
                    "h2d_bytes_per_step": B * ctypes.sizeof(_abi.View) + ctypes.sizeof(_abi.FrameParams),
                    "d2h_bytes_per_step": d2h_step, "steps": e2e_steps,
                    "runs": [round(v, 1) for v in e2e_runs], "reported": "median of the three runs",
                    "api": "Scene.render_batch_async(cameras, out=pinned ndarray), two batches in flight",
                    "pcie_gbs": pcie_gbs, "pcie_frac": e2e_fps * H * W * 3 / 1e9 / pcie_gbs if bands is None else None,
                    "pcie_what": f"measured here: {world} rank(s) copying a step's frames device -> pinned host "
                                 f"simultaneously (aggregate GB/s); e2e moves H*W*3 bytes per frame over the same link",
                    "host_affinity": (f"{numa_cpus} of {cpus_before} CPUs (GPU-local, NVML)" if numa_cpus else "unchanged")},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline}
    if e2e_single is not None:
        line["e2e_single"] = e2e_single
    if byte_check is not None:
        line["byte_check"] = byte_check
    if configs is not None:
        line["configs"] = configs
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_frames = args.cpu_frames or max(4, min(cores, 24))
        cams, dcams = step_cameras(10_000, 0, 1, n_frames, radius)
        cpu_port_sample(scene, cams[:1], dcams[:1], 1)                       # warm-up (page in)
        fps, dt = cpu_port_sample(scene, cams, dcams, cores)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n_frames} orbit frames of the same scene at {H}x{W}, {cores} threads "
                                          f"(one frame each), {dt:.1f} s wall; C restatement of the reference"}
        if not args.no_numpy_ref and numpy_reference_available():
            line["cpu_baseline"]["numpy_reference"] = numpy_reference_one_frame(workload, (H, W))
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
