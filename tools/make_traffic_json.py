"""profiles/traffic.json from an `ncu --set full` report of ONE batch (tools/profile_step.py <views> ...): per stage the
DRAM bytes of a launch (roofline.traffic of bench.py) and the ceiling ncu shows the kernel running against (issue-slot
utilisation; for the small set-up kernels the launch is shorter than a wave, so it is launch latency).
    python tools/make_traffic_json.py gpurun_out/prof.ncu-rep <frames_per_launch> > profiles/traffic.json"""
import csv
import io
import json
import subprocess
import sys

rep, frames = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
idx = {h: i for i, h in enumerate(rows[0])}
units = rows[1]
STAGE = {"k_tile": "raster", "k_shade_packed": "shade", "k_bin": "bin", "k_scan": "bin", "k_order": "bin", "k_clip_elide": "bin",
         "k_tri_setup": "tri_setup", "k_vertex": "tri_setup", "k_tri_count": "tri_setup", "k_quad_setup": "quad_setup",
         "k_facing": "silhouette", "k_silhouette": "silhouette", "k_frame_consts": "silhouette"}


def val(r, name, scale=1.0):
    try:
        v = float(r[idx[name]])
    except (KeyError, ValueError):
        return None
    u = units[idx[name]].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0) * scale


out = {"frames_per_launch": frames, "source": f"{rep} (ncu --set full --clock-control none, one {frames}-view batch)",
       "stage_ceilings": {}}
acc = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    key = next((s for k, s in STAGE.items() if name.startswith(k) or ("::" + k) in name or name.startswith("void " + k)), None)
    if key is None:
        key = next((s for k, s in STAGE.items() if k in name), None)
    if key is None:
        continue
    a = acc.setdefault(key, dict(bytes=0.0, ms=0.0, issue=0.0, fp64=0.0, warps=0.0, kernels=[]))
    t = val(r, "gpu__time_duration.sum") or 0.0
    tu = units[idx["gpu__time_duration.sum"]]
    t_ms = t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(tu, 1.0)
    a["bytes"] += (val(r, "dram__bytes_read.sum") or 0) + (val(r, "dram__bytes_write.sum") or 0)
    a["ms"] += t_ms
    a["issue"] += t_ms * (val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0)
    a["fp64"] += t_ms * (val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") or 0)
    a["warps"] += t_ms * (val(r, "sm__warps_active.avg.pct_of_peak_sustained_active") or 0)
    a["kernels"].append(name.split("(")[0])
for key, a in acc.items():
    out[key] = int(a["bytes"])
    ms = max(a["ms"], 1e-9)
    issue, fp64, warps = a["issue"] / ms / 100, a["fp64"] / ms / 100, a["warps"] / ms / 100
    small = key in ("silhouette", "tri_setup", "quad_setup") and a["ms"] < 0.2
    out["stage_ceilings"][key] = {
        "ceiling_name": "launch latency (sub-wave launches)" if small else "issue slots (ncu smsp__issue_active)",
        "ceiling_frac": round(issue, 3), "fp64_pipe_frac": round(fp64, 3), "warps_active_frac": round(warps, 3),
        "ncu_ms": round(a["ms"], 4), "ncu_dram_bytes_per_frame": int(a["bytes"] / frames),
        "kernels": sorted(set(a["kernels"]))}
print(json.dumps(out, indent=1))
