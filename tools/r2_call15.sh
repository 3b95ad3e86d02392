#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -5 > $O/r2c15_pytest.txt
cat $O/r2c15_pytest.txt
tools/ab_step.sh 64 4 diablo 2>&1 | tee $O/r2c15_ab_diablo.txt
tools/ab_step.sh 16 3 torus1m 2>&1 | tee $O/r2c15_ab_torus.txt
