#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs --no-numpy-ref 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print(round(d['value']), round(d['e2e']['value']), d['e2e_single']['value'], d['e2e_single']['verbose_off'], d['gpu_launches'], d['clocks'])"
