#!/bin/bash
# round 2, GPU call 7: per-vertex kernel, staged quad edges, 32-rounds: parity (incl. integration stub + fused test), benches,
# one full default bench run (configs, numpy reference, e2e_single)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > $O/r2c7_pytest.txt
for w in diablo synthetic; do
  timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs --workload $w > $O/r2c7_bench_$w.json 2> $O/r2c7_bench_$w.err
done
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-configs --workload torus1m > $O/r2c7_bench_torus1m.json 2> $O/r2c7_bench_torus1m.err
( time timeout 1200 python bench.py --steps 100 --warmup 5 > $O/r2c7_bench_full.json 2> $O/r2c7_bench_full.err ) 2> $O/r2c7_bench_full.time
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c7_bench_ref.json 2> $O/r2c7_bench_ref.err ) 2> $O/r2c7_bench_ref.time
cat $O/r2c7_pytest.txt
cat $O/r2c7_bench_full.time $O/r2c7_bench_ref.time
tail -c 600 $O/r2c7_bench_ref.json
for f in gpurun_out/r2c7_bench_diablo.json gpurun_out/r2c7_bench_synthetic.json gpurun_out/r2c7_bench_torus1m.json gpurun_out/r2c7_bench_full.json; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["e2e"]["value"]), round(d.get("e2e_single",{}).get("value",0)), {k: round(v,4) for k,v in d["roofline"]["stage_ms_per_step"].items()})
    for k,v in (d.get("configs") or {}).items():
        print("   ", k, {kk: (round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("value","failed")}, round(v.get("e2e",{}).get("value",0)), v.get("cpu_baseline",{}).get("value"))
    print("   cpu", d.get("cpu_baseline"))
except Exception as e: print(f, "failed", e)
PY
done
