#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/b_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_mma -s 1 -c 1 -f -o gpurun_out/prof_score_mma3 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu_f.log 2>&1
tail -12 gpurun_out/pytest_gpu.log | cut -c1-400; tail -1 gpurun_out/b_plain.log | cut -c1-300; tail -2 gpurun_out/ncu_f.log | cut -c1-200
