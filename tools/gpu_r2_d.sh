#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_multi.py -q -x -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/r2_d_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2_d_tests.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench.err
echo "rc=$?" >> gpurun_out/r2_bench.err
tail -5 gpurun_out/r2_d_tests.log | cut -c1-600; tail -5 gpurun_out/r2_bench.err; cut -c1-6000 gpurun_out/r2_bench.log
