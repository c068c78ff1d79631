#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench.log
# ncu: launch list of the bench step, full capture of the score kernel, launch list of the N=20k fit
timeout 600 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/b_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 1 -c 1 -o gpurun_out/prof_score python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu_f.log 2>&1
timeout 300 python tools/fit_profile.py 20000 2 > gpurun_out/fit_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_fit.csv python tools/fit_profile.py 20000 1 > gpurun_out/ncu_fit.log 2>&1
tail -25 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench.log; cat gpurun_out/fit_plain.log; tail -3 gpurun_out/ncu_f.log
