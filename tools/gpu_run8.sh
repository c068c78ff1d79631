#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/b_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_mma -s 1 -c 1 -o gpurun_out/prof_score_mma python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
