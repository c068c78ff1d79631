#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/b_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(score_|merit_|pick_|stats_|update_kernel|take_|negate_)' -c 120 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu_l.log 2>&1
tail -12 gpurun_out/pytest_gpu.log | cut -c1-400; tail -1 gpurun_out/bench.log | cut -c1-300; tail -2 gpurun_out/ncu_l.log | cut -c1-200
