#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > $O/r2c37_pytest.txt
cat $O/r2c37_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-numpy-ref > $O/r2c37_bench.json 2> $O/r2c37_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2c37_bench.json").read().splitlines() if l.startswith("{")][-1])
print(round(d["value"]), round(d["e2e"]["value"]), d.get("e2e_single",{}).get("value"), d["roofline"]["stage_ms_per_step"])
for k,v in d.get("configs",{}).items(): print(k, round(v["value"]), round(v["e2e"]["value"]))
PY
tail -c 400 $O/r2c37_bench.err
