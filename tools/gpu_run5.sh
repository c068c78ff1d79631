#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "gemm or lu or fit or rbf_scenario" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/fit_profile.py 20000 3 > gpurun_out/fit_plain.log 2>&1
timeout 300 python tools/fit_profile.py 5000 3 >> gpurun_out/fit_plain.log 2>&1
timeout 300 python tools/fit_profile.py 1000 3 >> gpurun_out/fit_plain.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_all.log
tail -5 gpurun_out/pytest_gpu.log; grep -v tflops gpurun_out/fit_plain.log; tail -15 gpurun_out/pytest_all.log
