#!/bin/bash
# round 2, final evidence on one GPU: full -m gpu suite, parity sweep, driver-style bench, launch list, full ncu captures
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > $O/r2c30_pytest.txt
timeout 600 python tools/parity_sweep.py > $O/r2c30_parity_sweep.txt 2>&1
B2R_LIB=$PWD/tools/stats_libs/variant_stats.so timeout 300 python tools/stats_step.py 8 diablo > $O/r2c30_stats_diablo.txt 2>&1
B2R_LIB=$PWD/tools/stats_libs/variant_stats.so timeout 300 python tools/stats_step.py 8 torus1m > $O/r2c30_stats_torus.txt 2>&1
timeout 900 python bench.py > $O/r2c30_bench.json 2> $O/r2c30_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-numpy-ref > $O/r2c30_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2c30_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-numpy-ref > $O/r2c30_ncu_launches.log 2>&1
timeout 300 python tools/profile_step.py 64 3 diablo > $O/r2c30_plain_diablo.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_' -s 16 -c 16 \
    -o $O/r2c30_prof_diablo -f python tools/profile_step.py 64 3 diablo > $O/r2c30_ncu_diablo.log 2>&1
timeout 600 python tools/profile_step.py 16 2 torus1m > $O/r2c30_plain_torus.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_' -s 16 -c 16 \
    -o $O/r2c30_prof_torus -f python tools/profile_step.py 16 2 torus1m > $O/r2c30_ncu_torus.log 2>&1
cat $O/r2c30_pytest.txt; tail -3 $O/r2c30_parity_sweep.txt; tail -3 $O/r2c30_plain_diablo.log $O/r2c30_plain_torus.log; tail -c 600 $O/r2c30_bench.err
ls -la $O/*.ncu-rep
