#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_tile|k_clip_elide|k_shade_packed' -s 3 -c 3 \
    -o $O/r2c21_prof -f python tools/profile_step.py 64 3 diablo > $O/r2c21_ncu.log 2>&1
ls -la $O/*.ncu-rep
