#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: frames/s (and Mpix/s) at 1080p for the diablo3-class mesh + floor
scene with shadow volumes, on N B200s, next to the reference algorithm timed on the box's host cores.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 2 --warmup 1        # CPU arm (oracle port of the reference)

A STEP = one camera-orbit batch of `--views` frames of the scene (BASELINE config 3: 5000-triangle figure with
1024^2 diffuse + tangent-space normal maps, floor with a 600^2 diffuse map, point light, shadow volumes,
1080x1920, LH/OpenGL).  Data is synthetic (no assets travel to the GPU box): same sizes as diablo3_pose.
Multi-GPU: weak scaling, every rank renders its own `--views` frames per step (frames are independent, no
data-path collective) and ONE NCCL gather per step assembles the uint8 frames on rank 0.

value : device-resident throughput (scene + pre-evaluated view constants in HBM, frames left in HBM): one CUDA-event
        pair around the K steps, L2 flushed before every step (inside the region), MAX over ranks.
e2e   : through the public API `Scene.render_batch_async(cameras, out=pinned)`: host camera maths + H2D of the view
        constants + render + D2H of the frames inside the timed region (wall clock, MAX over ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "frames_per_s_1080p_shadow_volumes"
UNIT = "frames/s"
HBM_FALLBACK_GBS = 6650.0
ORBIT_RADIUS = 3.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--views", type=int, default=64, help="frames per step per GPU")
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--workload", default="synthetic", choices=["synthetic", "diablo", "torus1m"],
                    help="synthetic = BASELINE config 3 stand-in (headline); diablo = the real assets if staged; "
                         "torus1m = BASELINE config 5 (1M-triangle displaced torus, camera orbit)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU sample (default: by core count)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0,
                    help="--impl reference: wall-clock bound of the whole run (steps are seconds each on the CPU)")
    ap.add_argument("--gather", default="auto", choices=["auto", "window", "window-copy", "nccl"],
                    help="N>1: how the frames reach rank 0 -- 'window': every rank's shading kernel stores into rank "
                         "0's buffer over NVLink (CUDA IPC, parallel.FrameWindow), the only collective is a one-element "
                         "all-reduce per step; 'window-copy': frames are rendered locally and pushed into the window by the "
                         "copy engines while the next step renders; 'nccl': one NCCL gather per step; 'auto': window for "
                         "N <= 2, window-copy beyond (seven ranks bursting 1.5 TB/s of stores into one GPU stall on NVLink)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def build_scene(args):
    import scenes
    res = (args.height, args.width)
    if args.workload == "diablo":
        assets = scenes.asset_root()
        if assets is None:
            raise SystemExit("--workload diablo needs the reference assets (baseline/_ref/assets)")
        return scenes.kat2(assets, res), "diablo3_pose(5022 tris, diffuse+tangent nm 1024^2) + floor(600^2), point light, shadow volumes"
    if args.workload == "torus1m":
        import py_numpy_renderer_b200 as b2r
        v, uv, n, f = scenes.torus_arrays(1000, 500)
        cam, dcam = scenes.std_cameras()
        sc = b2r.Scene(cam, scenes.std_light(), debug_camera=dcam, resolution=res, system=b2r.SYSTEM.LH,
                       subsystem=b2r.SUBSYSTEM.OPENGL)
        sc.verbose = False
        sc.add_model(b2r.Model(v, uv, n, f))
        return sc, "synthetic 1M-triangle displaced torus (nu=1000, nv=500), camera orbit radius 2.9, shadow volumes"
    return scenes.c3_synthetic(res), ("synthetic diablo3-class figure (5000 tris, diffuse+tangent normal map 1024^2) "
                                      "+ floor (2 tris, 600^2 diffuse), point light, shadow volumes")


def step_cameras(step, rank, world, views):
    """Orbit cameras of one step: every (step, rank, view) gets its own angle; the debug frustum contains the
    camera frustum."""
    import scenes
    start = 2 * np.pi * ((step * world + rank) * 0.6180339887498949 % 1.0)
    return (scenes.orbit_cameras(views, radius=ORBIT_RADIUS, start=start),
            scenes.orbit_cameras(views, radius=ORBIT_RADIUS, start=start, fovy=90, near=0.05, far=20))


def algorithmic_bytes(scene, n_shaded):
    """B_alg = B_geom + B_tex + B_fb per frame, in the reference's own storage."""
    b_geom = 0
    maps = []
    for m in scene.models:
        v = np.asarray(m.vertices)
        b_geom += v.shape[0] * v.dtype.itemsize * 4
        b_geom += (0 if m.uv is None else len(m.uv) * 12) + (0 if m.normals is None else len(m.normals) * 12)
        b_geom += len(m._faces) * 48
        for mat in m.materials.values():
            for attr in ("map_Kd", "map_Ks", "norm"):
                t = getattr(mat, attr, None)
                if t is not None:
                    maps.append(t.texels.shape[0] * t.texels.shape[1] * 3)
    b_tex = min(sum(maps), n_shaded * 3 * max(1, len(maps)))
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


def cpu_sample(scene, n_frames, threads):
    """The reference algorithm (oracle port, test infrastructure) on the host cores: n_frames orbit frames of the
    same workload, `threads` frames in flight.  Returns frames/s."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.build()
    cams, dcams = step_cameras(10_000, 0, 1, n_frames)
    # the oracle wrapper takes one debug camera per call: render per-view debug frusta by patching the scene
    from py_numpy_renderer_b200 import _abi
    packed = _abi.PackedScene(scene.models, None)
    fp = _abi.pack_frame_params(scene.light, scene.resolution, scene._background(), False)
    for c in cams + dcams:
        c.scene = scene
    views = (_abi.View * n_frames)(*[_abi.pack_view(c, d, scene.system, False) for c, d in zip(cams, dcams)])
    orc.render_packed(packed, fp, (_abi.View * 1)(views[0]), threads=1, planes=False)  # warm-up (page in)
    t0 = time.perf_counter()
    orc.render_packed(packed, fp, views, threads=threads, planes=False)
    dt = time.perf_counter() - t0
    return n_frames / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's own algorithm on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    scene, workload = build_scene(args)
    cores = os.cpu_count() or 1
    n_frames = args.cpu_frames or max(4, min(cores, 128))   # one frame per host thread: every core works
    # One step = n_frames frames, one per host thread (the port has no parallelism inside a frame: the reference's
    # per-face N-dependent evaluation order needs whole-box counts), i.e. seconds per step whatever the sample.  The
    # run is bounded in wall time: warm-up samples and timed steps stop when --cpu-budget-s is used up, and the line
    # reports how many steps were timed next to how many were asked for.
    t_start = time.perf_counter()
    for _ in range(max(0, min(args.warmup - 1, 1))):
        cpu_sample(scene, min(n_frames, cores), cores)
    total_t, total_f, done = 0.0, 0, 0
    for _ in range(args.steps):
        _, dt = cpu_sample(scene, n_frames, cores)
        total_t += dt
        total_f += n_frames
        done += 1
        if time.perf_counter() - t_start + dt > args.cpu_budget_s:
            break
    fps = total_f / total_t
    H, W = scene.resolution
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / done, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "mpix_per_s": fps * H * W / 1e6,
            "config": {"workload": workload, "resolution": [H, W], "frames_per_step": n_frames},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_frames} orbit frames per step, {cores} host threads (one frame each), "
                                       f"C restatement of Scene.render() pinned to the Python reference"},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


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

    scene, workload = build_scene(args)
    if args.workload == "torus1m":
        global ORBIT_RADIUS
        ORBIT_RADIUS = 2.9
    H, W = scene.resolution
    B, K, Wm = args.views, args.steps, args.warmup
    dev = scene._device_scene()
    lib_stream = torch.cuda.ExternalStream(_native.stream_ptr(), device=device)
    bg = scene._background()

    # one debug render: shaded-pixel count for the algorithmic-bytes model
    dbg = {}
    cams0, dcams0 = step_cameras(0, rank, world, 1)
    scene.render_batch(cams0, debug_cameras=dcams0, debug=dbg)
    n_shaded = int((dbg["winner"][0] >= 0).sum())
    balg = algorithmic_bytes(scene, n_shaded)

    # pre-evaluate the view constants of every step (host maths is part of e2e, not of `value`)
    packed_steps = []
    for s in range(Wm + K):
        cams, dcams = step_cameras(s, rank, world, B)
        for c in cams + dcams:
            c.scene = scene
        packed_steps.append(dev.pack(cams, dcams, scene.light, scene.resolution, scene.system, bg))
    frames_dev = [torch.empty((B, H, W, 3), dtype=torch.uint8, device=device) for _ in range(2)]
    if args.gather == "auto":
        args.gather = "window" if world <= 2 else "window-copy"
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
    push_stream = torch.cuda.Stream() if push else None
    win_blocks = [torch.as_tensor(win.block(s), device=device) for s in range(2)] if push else None
    gathered = [[torch.empty((B, H, W, 3), dtype=torch.uint8, device=device) for _ in range(world)]
                for _ in range(2)] if (world > 1 and rank == 0 and win is None) else [None, None]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    cur = torch.cuda.current_stream()

    def run_steps(first, count, per_step_sync):
        """`count` steps as a stream pipeline: [L2 flush] -> render (library stream) -> NCCL gather of the step's
        frames to rank 0 (NCCL stream), the gather of step s overlapping the render of step s+1 (two frame
        buffers).  Returns elapsed device milliseconds between the first flush and the last gather."""
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
                    win_blocks[slot].copy_(frames_dev[slot], non_blocking=True)
                    works[slot] = win.fence(async_op=True)
            elif win is not None:
                works[slot] = win.fence(async_op=True)      # one-element all-reduce: the step's frames are on rank 0
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
    if win is not None and rank == 0:   # every rank's block of the last step really holds a rendered frame
        last = torch.as_tensor(win.frames((Wm + K - 1) % 2), device=device)
        for r in range(world):
            blk = last[r * B]
            assert int(blk.max()) > int(blk.min()), f"window block of rank {r} is empty"
    value = K * B * world / (dev_ms / 1e3)
    # per-stage CUDA-event timing (library stream) on a few more steps of the same workload, synchronised per step
    n_stage_steps = min(K, 5)
    _native.set_stage_timing(True)
    run_steps(Wm + K - n_stage_steps, n_stage_steps, True)
    _native.set_stage_timing(False)

    # ---- end to end through the public API: host camera maths + H2D + render + D2H into pinned memory ----
    pinned = [torch.empty((B, H, W, 3), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
    e2e_steps = max(3, min(K, 20))
    for s in range(2):
        cams, dcams = step_cameras(1000 + s, rank, world, B)
        scene.render_batch(cams, debug_cameras=dcams, out=pinned[s])
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
            cams, dcams = step_cameras(2000 + 100 * rep + s, rank, world, B)
            fut = scene.render_batch_async(cams, debug_cameras=dcams, out=pinned[s % 2])
            if in_flight is not None:
                in_flight.result()                                      # frames of step s-1 are in host memory
            in_flight = fut
        in_flight.result()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_runs.append(e2e_steps * B * world / float(t.item()))
    e2e_fps = sorted(e2e_runs)[1]
    sampler.stop_flag = True
    sampler.join(timeout=1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    from py_numpy_renderer_b200 import _abi
    import ctypes
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
    top = max(stage_ms, key=stage_ms.get) if stage_ms else None
    stage_avg = {k: v / n_stage_steps for k, v in stage_ms.items()}
    roofline = None
    if top:
        achieved = balg["total"] * B / (stage_avg[top] / 1e3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))   # ncu dram bytes of one launch of `frames_per_launch` frames: scale to B
            traffic = tj.get(top)
            if traffic is not None:
                traffic = int(traffic * B / tj.get("frames_per_launch", B))
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_frame": balg, "frames_per_launch": B,
                    "kernel_ms_per_launch": stage_avg[top], "stage_ms_per_step": stage_avg,
                    "whole_frame_frac": balg["total"] * value / world / 1e9 / peak,
                    # every stage against the same ceiling: the frame's algorithmic bytes over the stage's own time
                    "stage_frac": {k: (balg["total"] * B / (v / 1e3) / 1e9 / peak) if v > 0 else None
                                   for k, v in stage_avg.items()},
                    "secondary_ceilings": "profiles/README.md (FP64 issue, shared / L2 atomics, PCIe: measured on the box)",
                    "note": "HBM is the roofline the contract names, not what binds this path: the float64 semantics "
                            "of the reference make raster / shade issue- and latency-bound (ncu, 16-view launch: issue "
                            "slots 72 % / 64 % busy, DRAM 3 % / 7 % of peak)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic" if args.workload == "synthetic" else "reference assets",
            "mpix_per_s": value * H * W / 1e6,
            "config": {"workload": workload, "resolution": [H, W], "frames_per_step_per_gpu": B,
                       "parallelism": (f"frames x{world} (weak); frames reach rank 0 by: " +
                                       ("peer stores of the shading kernel into rank 0's window over NVLink + a "
                                        "one-element all-reduce per step" if gather_mode == "window" else
                                        "copy-engine pushes into rank 0's window over NVLink, overlapped with the next "
                                        "render, + a one-element all-reduce per step" if gather_mode == "window-copy" else
                                        f"{gather_mode} gather per step, overlapped with the next render"))
                       if world > 1 else "single GPU",
                       "l2": "flushed before every step (256 MiB memset, inside the timed region)",
                       "timing": "one CUDA-event pair around the K steps (first flush .. last gather), MAX over ranks; "
                                 "stage times from a separate per-step-synchronised pass",
                       "wall_ms_per_step": 1e3 * wall / K},
            "e2e": {"value": e2e_fps, "unit": UNIT,
                    "h2d_bytes_per_step": B * ctypes.sizeof(_abi.View) + ctypes.sizeof(_abi.FrameParams),
                    "d2h_bytes_per_step": B * H * W * 3, "steps": e2e_steps,
                    "runs": [round(v, 1) for v in e2e_runs], "reported": "median of the three runs",
                    "api": "Scene.render_batch_async(cameras, out=pinned ndarray), two batches in flight",
                    "host_affinity": (f"{numa_cpus} of {cpus_before} CPUs (GPU-local, NVML)" if numa_cpus else "unchanged")},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_frames = args.cpu_frames or max(4, min(cores, 24))
        fps, dt = cpu_sample(scene, n_frames, cores)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n_frames} orbit frames of the same scene at {H}x{W}, {cores} threads "
                                          f"(one frame each), {dt:.1f} s wall; C restatement of the reference "
                                          f"(the NumPy reference itself: 0.118 frames/s, SURVEY.md section 6)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
