#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_incremental.py -q -s -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/r2_incr.log 2>&1
echo "rc=$?" >> gpurun_out/r2_incr.log
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider --deselect tests/test_gpu_incremental.py --deselect tests/test_gpu_live_strategies.py > gpurun_out/r2_pytest_rest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_pytest_rest.log
tail -40 gpurun_out/r2_incr.log | cut -c1-600; tail -12 gpurun_out/r2_pytest_rest.log | cut -c1-600
