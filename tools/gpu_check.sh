#!/bin/bash
# GPU round trip used during development: parity suite, then a short bench; prints the one-line summary.
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest_tail.txt
python bench.py --steps ${1:-100} --no-cpu-baseline > gpurun_out/b.json 2>gpurun_out/b.err
cat gpurun_out/pytest_tail.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/b.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 4) for k, v in d["roofline"]["stage_ms_per_step"].items()})
PY
