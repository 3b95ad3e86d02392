#!/bin/bash
cd "$(dirname "$0")/.."
X="--no-cpu-baseline --no-configs --no-numpy-ref"
for gb in 6 30; do echo "== B2R_SCRATCH_GB=$gb"; B2R_SCRATCH_GB=$gb timeout 400 python bench.py --workload torus1m --views 64 --steps 6 --warmup 3 $X 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print(round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['roofline']['stage_ms_per_step'])"; done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "c5 or kat2 or fixture" 2>&1 | tail -3
