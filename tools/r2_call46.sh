#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python bench.py --no-cpu-baseline --no-numpy-ref --steps 60 2>gpurun_out/r2c46.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print(round(d['value']), round(d['e2e']['value']), d['e2e_single']['value'])
for k,v in d.get('configs',{}).items(): print(k, round(v['value']), round(v['e2e']['value']), v['frames_per_step'], round(v['roofline']['frac'],4))"
tail -c 300 gpurun_out/r2c46.err
