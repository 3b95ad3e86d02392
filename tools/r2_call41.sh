#!/bin/bash
# round 2, last single-GPU verification: full -m gpu suite, smoke, driver-style default bench
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > $O/r2c41_pytest.txt
cat $O/r2c41_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 900 python bench.py > $O/r2c41_bench.json 2> $O/r2c41_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2c41_bench.json").read().splitlines() if l.startswith("{")][-1])
print(round(d["value"]), round(d["e2e"]["value"]), d.get("e2e_single"), d["roofline"]["stage_ms_per_step"])
for k,v in d.get("configs",{}).items(): print(k, round(v["value"]), round(v["e2e"]["value"]))
print(d["cpu_baseline"])
PY
tail -c 300 $O/r2c41_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
