#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -8 > $O/r2c25_pytest.txt
cat $O/r2c25_pytest.txt
echo "== default"; timeout 300 python tools/profile_step.py 64 4 diablo 2>&1 | tail -3
echo "== torus default"; timeout 300 python tools/profile_step.py 16 3 torus1m 2>&1 | tail -3
B2R_LIB=$PWD/tools/stats_libs/variant_stats.so timeout 300 python tools/stats_step.py 8 diablo 2>&1 | tail -17 | tee $O/r2c25_stats.txt
