"""The one-line JSON contract of bench.py (both arms): keys the driver reads, types, and internal consistency."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(*flags):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line_and_wall_budget():
    """--impl reference: the oracle port on the host cores, bounded in wall time whatever --steps says."""
    d = run_bench("--impl", "reference", "--ref-kind", "port", "--steps", "50", "--warmup", "1", "--height", "135",
                  "--width", "240", "--cpu-frames", "2", "--cpu-budget-s", "3")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert 1 <= d["steps"] <= 50 and d["steps_requested"] == 50
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_runs_the_unmodified_numpy_reference_when_staged():
    """--ref-kind numpy: the reference's own Scene.render() in worker processes (needs /root/reference or the zip that
    __graft_entry__.build() stages under baseline/_ref/)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refboot
    if not refboot.importable():
        pytest.skip("reference sources neither present nor staged")
    d = run_bench("--impl", "reference", "--ref-kind", "numpy", "--steps", "1", "--warmup", "0", "--height", "90",
                  "--width", "160", "--cpu-frames", "2", "--ref-procs", "2", "--workload", "synthetic")
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["value"] > 0
    assert d["cpu_baseline"]["single_process_fps"] > 0 and d["steps"] == 1


@pytest.mark.gpu
def test_b200_arm_line():
    d = run_bench("--steps", "4", "--warmup", "3", "--views", "3", "--height", "270", "--width", "480",
                  "--cpu-frames", "2", "--no-configs", "--no-numpy-ref")
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline", "cpu_baseline", "mpix_per_s"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] == 3 and d["scaling"] == "weak"
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and d["value"] > 0 and d["gpu_launches"] > 0
    assert abs(d["value"] - 3 * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]          # frames per step / time per step
    e = d["e2e"]
    assert e["value"] > 0 and e["d2h_bytes_per_step"] == 3 * 270 * 480 * 3 and e["h2d_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and set(r["stages"]) <= set(r["stage_ms_per_step"])
    for st in r["stages"].values():                       # every stage against its OWN bytes: nothing reads far above 1
        assert 0 < st["hbm_frac"] < 1.2 and st["algorithmic_bytes_per_frame"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert d["e2e_single"]["value"] > 0 and d["e2e"]["pcie_gbs"] > 0
    assert set(d["config"]["workload_stats"]) >= {"faces", "silhouette_edges", "shaded_pixels"}
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
