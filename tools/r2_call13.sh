#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15 > $O/r2c13_pytest.txt
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs > $O/r2c13_bench.json 2> $O/r2c13_bench.err
cat $O/r2c13_pytest.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c13_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), d.get("e2e_single",{}).get("value"), d["roofline"]["stage_ms_per_step"])
PY
