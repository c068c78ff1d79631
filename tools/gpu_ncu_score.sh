#!/bin/bash
# ncu --set full of the bench's dominant kernel (after the same command exited 0 without ncu), plus the strategy replay timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/b_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_mma -s 1 -c 1 -f -o gpurun_out/prof_score_mma4 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu_f.log 2>&1
timeout 900 python tools/replay_timing.py > gpurun_out/replay_timing3.log 2>&1
tail -1 gpurun_out/b_plain.log | cut -c1-200; tail -2 gpurun_out/ncu_f.log | cut -c1-200; tail -3 gpurun_out/replay_timing3.log | cut -c1-420
