#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > $O/r2c39_pytest.txt
cat $O/r2c39_pytest.txt
timeout 300 python tools/single_breakdown.py 2>&1 | tail -8
for p in 1 2 8; do echo "== B2R_SPLIT=$p"; B2R_SPLIT=$p timeout 300 python tools/single_breakdown.py 2>&1 | grep -E "device|stage|quiet"; done
timeout 300 python tools/profile_step.py 64 3 diablo 2>&1 | tail -2
