#!/bin/bash
# round 2, GPU call 2: parity of the fused tile kernel, work counters, bench of the three workloads, ncu captures
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > $O/r2c2_pytest_fused.txt
for w in diablo synthetic; do
  B2R_LIB=$PWD/tools/variant_stats.so timeout 300 python tools/stats_step.py 8 $w > $O/r2c2_stats_${w}.txt 2>&1
done
for w in diablo synthetic; do
  timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --workload $w > $O/r2c2_bench_${w}.json 2> $O/r2c2_bench_${w}.err
done
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --workload torus1m > $O/r2c2_bench_torus1m.json 2> $O/r2c2_bench_torus1m.err
for lib in tools/variant_*.so; do
  [ "$lib" = tools/variant_stats.so ] && continue
  B2R_LIB=$PWD/$lib timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --workload diablo > $O/r2c2_bench_diablo_$(basename $lib .so).json 2>&1
done
timeout 300 python tools/profile_step.py 64 3 diablo > $O/r2c2_plain_diablo.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_tile' -s 1 -c 1 \
    -o $O/r2c2_prof_diablo -f python tools/profile_step.py 64 3 diablo > $O/r2c2_ncu_diablo.log 2>&1
timeout 600 python tools/profile_step.py 8 2 torus1m > $O/r2c2_plain_torus.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_tile' -s 1 -c 1 \
    -o $O/r2c2_prof_torus -f python tools/profile_step.py 8 2 torus1m > $O/r2c2_ncu_torus.log 2>&1
cat $O/r2c2_pytest_fused.txt
tail -n 3 $O/r2c2_plain_diablo.log $O/r2c2_plain_torus.log
for f in gpurun_out/r2c2_bench_*.json; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["e2e"]["value"]), {k: round(v,4) for k,v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e: print(f, "failed", e)
PY
done
