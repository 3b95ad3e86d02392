#!/bin/bash
# round 2, GPU call 1: parity of the fused tile kernel and of the refactored API (both kernel paths), work counters,
# A/B bench of the three workloads, ncu captures (diablo 64 views, torus1m)
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2c1_smi.txt
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > $O/r2c1_pytest_fused.txt
B2R_FUSED=0 timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > $O/r2c1_pytest_unfused.txt
for f in 1 0; do for w in diablo synthetic; do
  B2R_FUSED=$f B2R_LIB=$PWD/tools/variant_stats.so timeout 300 python tools/stats_step.py 8 $w > $O/r2c1_stats_${w}_f$f.txt 2>&1
done; done
for f in 1 0; do for w in diablo synthetic; do
  B2R_FUSED=$f timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --workload $w > $O/r2c1_bench_${w}_f$f.json 2> $O/r2c1_bench_${w}_f$f.err
done
B2R_FUSED=$f timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --workload torus1m > $O/r2c1_bench_torus1m_f$f.json 2> $O/r2c1_bench_torus1m_f$f.err
done
timeout 300 python tools/profile_step.py 64 3 diablo > $O/r2c1_plain_diablo.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_tile' -s 1 -c 1 \
    -o $O/r2c1_prof_diablo -f python tools/profile_step.py 64 3 diablo > $O/r2c1_ncu_diablo.log 2>&1
timeout 600 python tools/profile_step.py 8 2 torus1m > $O/r2c1_plain_torus.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_tile|k_bin|k_tri_setup' -s 4 -c 4 \
    -o $O/r2c1_prof_torus -f python tools/profile_step.py 8 2 torus1m > $O/r2c1_ncu_torus.log 2>&1
cat $O/r2c1_pytest_fused.txt $O/r2c1_pytest_unfused.txt
tail -3 $O/r2c1_plain_diablo.log $O/r2c1_plain_torus.log
for f in 1 0; do for w in diablo synthetic torus1m; do python - $w $f <<'PY'
import json,sys
w,f=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/r2c1_bench_{w}_f{f}.json").read().strip().splitlines()[-1])
    print(w, "fused" if f=="1" else "unfused", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,4) for k,v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e: print(w, f, "failed", e)
PY
done; done
