#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python tools/single_breakdown.py 2>&1 | tail -12 | tee gpurun_out/r2c38_single.txt
