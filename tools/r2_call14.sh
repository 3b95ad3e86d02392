#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -5 > $O/r2c14_pytest.txt
cat $O/r2c14_pytest.txt
tools/ab_step.sh 64 4 diablo 2>&1 | tee $O/r2c14_ab_diablo.txt
tools/ab_step.sh 16 3 torus1m 2>&1 | tee $O/r2c14_ab_torus.txt
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_tile' -s 2 -c 1 \
    -o $O/r2c14_prof_tile -f python tools/profile_step.py 64 3 diablo > $O/r2c14_ncu.log 2>&1
ls -la $O/*.ncu-rep
