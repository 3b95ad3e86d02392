#!/bin/bash
# round 2, GPU call 5: pipelined set-up / raster, TriBox binning, silhouette capacity: parity + pipe sweep + torus1m
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > $O/r2c5_pytest.txt
for p in 64 32 16 8; do
  B2R_PIPE=$p timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs --workload diablo > $O/r2c5_bench_diablo_pipe$p.json 2> $O/r2c5_bench_diablo_pipe$p.err
done
for p in 64 16; do
B2R_PIPE=$p timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-configs --workload torus1m > $O/r2c5_bench_torus1m_pipe$p.json 2> $O/r2c5_bench_torus1m_pipe$p.err
done
timeout 600 python tools/profile_step.py 16 2 torus1m > $O/r2c5_plain_torus.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_tile|k_bin|k_tri_setup|k_shade' -s 5 -c 5 \
    -o $O/r2c5_prof_torus -f python tools/profile_step.py 16 2 torus1m > $O/r2c5_ncu_torus.log 2>&1
cat $O/r2c5_pytest.txt
cat $O/r2c5_plain_torus.log
for f in gpurun_out/r2c5_bench_*.json; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["e2e"]["value"]), round(d.get("e2e_single",{}).get("value",0)), {k: round(v,4) for k,v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e: print(f, "failed", e)
PY
done
