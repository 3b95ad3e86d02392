#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -8 > $O/r2c22_pytest.txt
cat $O/r2c22_pytest.txt
echo "== default"; timeout 300 python tools/profile_step.py 64 4 diablo 2>&1 | tail -3
echo "== torus default"; timeout 300 python tools/profile_step.py 16 3 torus1m 2>&1 | tail -3
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs --no-numpy-ref > $O/r2c22_bench.json 2> $O/r2c22_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c22_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), d.get("e2e_single",{}).get("value"), d["roofline"]["stage_ms_per_step"])
PY
