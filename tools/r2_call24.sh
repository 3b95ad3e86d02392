#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "diablo or window or async or device_resident" 2>&1 | tail -4
tools/ab_step.sh 64 3 diablo 2>&1 | grep -v "^1 " | tee $O/r2c24_skip.txt
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs --no-numpy-ref > $O/r2c24_bench.json 2> $O/r2c24_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c24_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), d.get("e2e_single",{}).get("value"), d["roofline"]["stage_ms_per_step"])
PY
