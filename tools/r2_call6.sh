#!/bin/bash
# round 2, GPU call 6: lean tri_setup + k_tri_count, 62-triangle rounds: parity, diablo / synthetic / torus1m, torus profile
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > $O/r2c6_pytest.txt
for w in diablo synthetic; do
  timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs --workload $w > $O/r2c6_bench_$w.json 2> $O/r2c6_bench_$w.err
done
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-configs --workload torus1m > $O/r2c6_bench_torus1m.json 2> $O/r2c6_bench_torus1m.err
timeout 600 python tools/c5_check.py 1 --oracle > $O/r2c6_c5check.txt 2>&1
timeout 600 python tools/profile_step.py 16 2 torus1m > $O/r2c6_plain_torus.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_tile|k_bin|k_tri_setup|k_tri_count|k_shade' -s 6 -c 6 \
    -o $O/r2c6_prof_torus -f python tools/profile_step.py 16 2 torus1m > $O/r2c6_ncu_torus.log 2>&1
cat $O/r2c6_pytest.txt $O/r2c6_c5check.txt
cat $O/r2c6_plain_torus.log
for f in gpurun_out/r2c6_bench_*.json; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["e2e"]["value"]), round(d.get("e2e_single",{}).get("value",0)), {k: round(v,4) for k,v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e: print(f, "failed", e)
PY
done
