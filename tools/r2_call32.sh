#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "clip_elision or fixture or kat2" 2>&1 | tail -8 > $O/r2c32_pytest.txt
cat $O/r2c32_pytest.txt
tools/ab_step.sh 64 3 diablo 2>&1 | grep -v "^1 " | tee $O/r2c32_vshade.txt
