#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench.log
timeout 300 build/tune/dmma_probe > gpurun_out/dmma_probe.log 2>&1
timeout 300 python tools/gemm_profile.py 8192 128 > gpurun_out/gemm_plain.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/b_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:score|merit|pick|stats|update|take|negate' -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:score_kernel<50, 0, 0, 1>' -s 1 -c 1 -o gpurun_out/prof_score python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu_f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_sub -s 1 -c 1 -o gpurun_out/prof_gemm python tools/gemm_profile.py 8192 128 > gpurun_out/ncu_g.log 2>&1
tail -8 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench.log | cut -c1-600; cat gpurun_out/gemm_plain.log; tail -3 gpurun_out/ncu_f.log; tail -3 gpurun_out/ncu_g.log
