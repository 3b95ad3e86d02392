#!/bin/bash
# refresh of the launch list with the last kernels of the round (split single-view launches) + one full capture of a split launch
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-numpy-ref > $O/r2c42_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2c42_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-numpy-ref > $O/r2c42_ncu_launches.log 2>&1
timeout 300 python tools/profile_step.py 1 3 diablo > $O/r2c42_plain_single.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_tile' -s 2 -c 1 \
    -o $O/r2c42_prof_single -f python tools/profile_step.py 1 3 diablo > $O/r2c42_ncu_single.log 2>&1
tail -2 $O/r2c42_plain_single.log; ls -la $O/r2c42*
