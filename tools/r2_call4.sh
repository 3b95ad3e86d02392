#!/bin/bash
# round 2, GPU call 4: k_tile with the dense loops: parity in mode 2 (full) and mode 1 (subset), A/B bench of the modes
cd "$(dirname "$0")/.."
O=gpurun_out
B2R_FUSED=2 timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > $O/r2c4_pytest_m2.txt
B2R_FUSED=1 timeout 900 python -m pytest tests -q -m gpu -x -k "parity or random or reference_boundary" 2>&1 | tail -8 > $O/r2c4_pytest_m1.txt
for m in 0 1 2; do
  B2R_FUSED=$m timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs --workload diablo > $O/r2c4_bench_diablo_m$m.json 2> $O/r2c4_bench_diablo_m$m.err
done
B2R_FUSED=2 timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-configs --workload torus1m > $O/r2c4_bench_torus1m_m2.json 2> $O/r2c4_bench_torus1m_m2.err
B2R_FUSED=2 timeout 300 python tools/profile_step.py 64 3 diablo > $O/r2c4_plain_diablo.log 2>&1 &&
B2R_FUSED=2 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_tile|k_shade_packed' -s 2 -c 2 \
    -o $O/r2c4_prof_diablo -f python tools/profile_step.py 64 3 diablo > $O/r2c4_ncu_diablo.log 2>&1
cat $O/r2c4_pytest_m2.txt $O/r2c4_pytest_m1.txt
for f in gpurun_out/r2c4_bench_*.json; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["e2e"]["value"]), round(d.get("e2e_single",{}).get("value",0)), {k: round(v,4) for k,v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e: print(f, "failed", e)
PY
done
