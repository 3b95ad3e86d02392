#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -30 > $O/r2c16_pytest.txt
cat $O/r2c16_pytest.txt
for v in stats stats_noelide; do echo "== $v"; B2R_LIB=$PWD/tools/variant_$v.so timeout 300 python tools/stats_step.py 8 diablo 2>&1 | tail -17; done | tee $O/r2c16_stats.txt
echo "== default (f32 lighting)"; timeout 300 python tools/profile_step.py 64 4 diablo 2>&1 | tail -3
echo "== B2R_SHADE_F64=1"; B2R_SHADE_F64=1 timeout 300 python tools/profile_step.py 64 4 diablo 2>&1 | tail -3
timeout 600 python tools/parity_sweep.py > $O/r2c16_parity_sweep.txt 2>&1; tail -25 $O/r2c16_parity_sweep.txt
