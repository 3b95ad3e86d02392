#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > $O/r2c40_pytest.txt
cat $O/r2c40_pytest.txt
timeout 300 python tools/single_breakdown.py 2>&1 | tail -7
timeout 300 python tools/profile_step.py 64 3 diablo 2>&1 | tail -2
