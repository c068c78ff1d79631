#!/bin/bash
# first GPU session: tiling sweep, fp64 peaks, parity suite, smoke, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== tune" > gpurun_out/tune.log
for f in build/tune/k2_*; do timeout 120 $f 262144 20000 3 >> gpurun_out/tune.log 2>&1 || echo "FAILED $f rc=$?" >> gpurun_out/tune.log; done
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench.log
tail -5 gpurun_out/tune.log; tail -30 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -3 gpurun_out/bench.log
