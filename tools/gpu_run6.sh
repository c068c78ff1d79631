#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/fit_plain.log
for cfg in "20000 2 0 3" "20000 2 0 2" "20000 3 1 0" "20000 2 1 3" "5000 3 1 0" "5000 2 0 3" "1000 3 1 0" "1000 2 0 3"; do
  timeout 120 python tools/fit_profile.py $cfg >> gpurun_out/fit_plain.log 2>&1 || echo "FAILED/TIMEOUT: $cfg rc=$?" >> gpurun_out/fit_plain.log
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "gemm or lu or fit or rbf_scenario" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python -m pytest tests/test_gpu_strategy_traces.py -m gpu -q --timeout 500 -p no:cacheprovider > gpurun_out/pytest_traces.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_traces.log
cat gpurun_out/fit_plain.log; tail -5 gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_traces.log
