#!/bin/bash
cd "$(dirname "$0")/.."
for c in 4 8 16; do for a in 1 2 3; do
  B2R_ASYNC_CHUNK=$c B2R_AUX_HOST=$a python bench.py --steps 30 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('async_chunk', $c, 'aux_host', $a, round(d['e2e']['value']))"
done; done
