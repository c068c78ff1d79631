#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== k2x tune" > gpurun_out/tune_mma.log
for f in build/tune/k2x_mb2_b4 build/tune/k2x_mb2_b3 build/tune/k2x_mb1_b4 build/tune/k2x_mb2_b3_d10 build/tune/k2x_mb2_b5_d10 build/tune/k2x_mb2_b3_d30; do timeout 120 $f 262144 20000 3 >> gpurun_out/tune_mma.log 2>&1; done
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
cat gpurun_out/tune_mma.log | cut -c1-150; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300; tail -1 gpurun_out/bench.log | cut -c1-200
