#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_window.py tests/test_cabi.py -q -m gpu 2>&1 | tail -25 > $O/r2c33_pytest.txt
cat $O/r2c33_pytest.txt
