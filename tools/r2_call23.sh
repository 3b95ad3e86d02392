#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -8 > $O/r2c23_pytest.txt
cat $O/r2c23_pytest.txt
echo "== default"; timeout 300 python tools/profile_step.py 64 4 diablo 2>&1 | tail -3
echo "== torus default"; timeout 300 python tools/profile_step.py 16 3 torus1m 2>&1 | tail -3
