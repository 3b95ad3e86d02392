#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -30 > $O/r2c17_pytest.txt
cat $O/r2c17_pytest.txt
B2R_LIB=$PWD/tools/stats_libs/variant_stats.so timeout 300 python tools/stats_step.py 8 diablo 2>&1 | tail -17 | tee $O/r2c17_stats.txt
tools/ab_step.sh 64 4 diablo 2>&1 | tee $O/r2c17_ab_diablo.txt
tools/ab_step.sh 16 3 torus1m 2>&1 | tee $O/r2c17_ab_torus.txt
