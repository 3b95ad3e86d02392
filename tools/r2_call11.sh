#!/bin/bash
# round 2, GPU call 11: occupancy variants of the tile kernel (diablo + synthetic)
cd "$(dirname "$0")/.."
O=gpurun_out
for lib in default tools/variant_tile10.so tools/variant_tile8.so; do
  if [ "$lib" = default ]; then unset B2R_LIB; tag=default; else export B2R_LIB=$PWD/$lib; tag=$(basename $lib .so); fi
  for w in diablo synthetic; do
    timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --no-configs --workload $w > $O/r2c11_${tag}_$w.json 2> $O/r2c11_${tag}_$w.err
  done
done
unset B2R_LIB
for f in gpurun_out/r2c11_*.json; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["e2e"]["value"]), {k: round(v,4) for k,v in d["roofline"]["stage_ms_per_step"].items()})
except Exception as e: print(f, "failed", e)
PY
done
