#!/bin/bash
# round 2, GPU call 12: final evidence: parity suite, parity sweep, work counters, launch list, full ncu captures
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 > $O/r2c12_pytest.txt
timeout 600 python tools/parity_sweep.py > $O/r2c12_parity_sweep.txt 2>&1
B2R_LIB=$PWD/tools/variant_stats.so timeout 300 python tools/stats_step.py 8 diablo > $O/r2c12_stats_diablo.txt 2>&1
B2R_FUSED=1 timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-configs > $O/r2c12_bench_fused.json 2> $O/r2c12_bench_fused.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-numpy-ref > $O/r2c12_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2c12_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-numpy-ref > $O/r2c12_ncu_launches.log 2>&1
timeout 300 python tools/profile_step.py 64 3 diablo > $O/r2c12_plain_diablo.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_' -s 15 -c 15 \
    -o $O/r2c12_prof_diablo -f python tools/profile_step.py 64 3 diablo > $O/r2c12_ncu_diablo.log 2>&1
timeout 600 python tools/profile_step.py 16 2 torus1m > $O/r2c12_plain_torus.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_' -s 15 -c 15 \
    -o $O/r2c12_prof_torus -f python tools/profile_step.py 16 2 torus1m > $O/r2c12_ncu_torus.log 2>&1
B2R_FUSED=1 timeout 300 python tools/profile_step.py 64 3 diablo > $O/r2c12_plain_fused.log 2>&1 &&
B2R_FUSED=1 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_tile' -s 1 -c 1 \
    -o $O/r2c12_prof_fused -f python tools/profile_step.py 64 3 diablo > $O/r2c12_ncu_fused.log 2>&1
cat $O/r2c12_pytest.txt; tail -4 $O/r2c12_parity_sweep.txt; tail -3 $O/r2c12_plain_diablo.log $O/r2c12_plain_torus.log $O/r2c12_plain_fused.log
