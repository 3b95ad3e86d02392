#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -4
tools/ab_step.sh 64 3 diablo 2>&1 | grep -v "^1 "
tools/ab_step.sh 16 2 torus1m 2>&1 | grep -v "^0 "
tools/ab_step.sh 1 3 diablo 2>&1 | grep -v "^1 "
