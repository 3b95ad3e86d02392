#!/bin/bash
# round 2, GPU call 1: sanity + work counters + baselines of the three workloads + ncu captures (diablo 64 views, torus1m)
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2c1_smi.txt
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > $O/r2c1_pytest.txt
for w in diablo synthetic; do
  B2R_LIB=$PWD/tools/variant_stats.so timeout 300 python tools/stats_step.py 8 $w > $O/r2c1_stats_$w.txt 2>&1
done
for w in diablo synthetic; do
  timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --workload $w > $O/r2c1_bench_$w.json 2> $O/r2c1_bench_$w.err
done
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --workload torus1m > $O/r2c1_bench_torus1m.json 2> $O/r2c1_bench_torus1m.err
timeout 300 python tools/profile_step.py 64 3 diablo > $O/r2c1_plain_diablo.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_raster|k_shade' -s 2 -c 2 \
    -o $O/r2c1_prof_diablo -f python tools/profile_step.py 64 3 diablo > $O/r2c1_ncu_diablo.log 2>&1
timeout 600 python tools/profile_step.py 8 2 torus1m > $O/r2c1_plain_torus.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_raster|k_shade|k_bin|k_tri_setup' -s 5 -c 5 \
    -o $O/r2c1_prof_torus -f python tools/profile_step.py 8 2 torus1m > $O/r2c1_ncu_torus.log 2>&1
cat $O/r2c1_pytest.txt
tail -3 $O/r2c1_plain_diablo.log $O/r2c1_plain_torus.log
for w in diablo synthetic torus1m; do python - $w <<'PY'
import json,sys
w=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2c1_bench_{w}.json").read().strip().splitlines()[-1])
    print(w, round(d["value"]), round(d["e2e"]["value"]), {k: round(v,4) for k,v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e: print(w, "failed", e)
PY
done
