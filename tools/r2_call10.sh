#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > $O/r2c10_pytest.txt
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs > $O/r2c10_bench.json 2> $O/r2c10_bench.err
cat $O/r2c10_pytest.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c10_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), d.get("e2e_single"), d["e2e"].get("pcie_gbs"), d["roofline"]["stage_ms_per_step"])
PY
