#!/bin/bash
# usage: tools/variant_sweep.sh  -- runs the bench with every tools/variant_*.so (tuning builds) and the default library
cd "$(dirname "$0")/.."
for lib in default tools/variant_*.so; do
  if [ "$lib" = default ]; then unset B2R_LIB; else export B2R_LIB=$PWD/$lib; fi
  python bench.py --steps 40 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms_per_step']; print('$lib', round(d['value']), 'raster', round(s['raster'],3), 'shade', round(s['shade'],3))"
done
