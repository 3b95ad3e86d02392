#!/bin/bash
# device-resident throughput for a few (B2R_DEV_CHUNK, B2R_AUX_DEV) settings
cd "$(dirname "$0")/.."
for c in 8 16 32 64; do for a in 2 3; do
  B2R_DEV_CHUNK=$c B2R_AUX_DEV=$a python bench.py --steps 60 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dev_chunk', $c, 'aux', $a, round(d['value']), round(d['e2e']['value']))"
done; done
