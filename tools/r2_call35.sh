#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_window.py -q -m gpu 2>&1 | tail -8
N=2 bash tools/r2_call34.sh
