#!/bin/bash
# full single-GPU check under gpurun: parity suite, smoke, bench, ncu launch lists of the bench and of one N = 20000 fit
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/b_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(score_|merit_|pick_|stats_|update_kernel|take_|negate_)' -c 120 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu_l.log 2>&1
timeout 300 python tools/fit_profile.py 20000 1 0 0 0 > gpurun_out/fit_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_fit_chol3.csv python tools/fit_profile.py 20000 1 0 0 0 > gpurun_out/ncu_fit.log 2>&1
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300; tail -2 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log | cut -c1-2500
