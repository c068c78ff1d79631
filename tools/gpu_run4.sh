#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "gemm or lu or fit or rbf_scenario" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/gemm_profile.py 8192 128 > gpurun_out/gemm_plain.log 2>&1
timeout 300 python tools/fit_profile.py 20000 2 > gpurun_out/fit_plain.log 2>&1
timeout 300 python tools/fit_profile.py 5000 2 >> gpurun_out/fit_plain.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/b_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_kernel -s 2 -c 1 -o gpurun_out/prof_score python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu_f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_sub -s 1 -c 1 -o gpurun_out/prof_gemm python tools/gemm_profile.py 8192 128 > gpurun_out/ncu_g.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:score|merit|pick|stats|update|take|negate' -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu_l.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/gemm_plain.log gpurun_out/fit_plain.log; tail -2 gpurun_out/ncu_f.log
