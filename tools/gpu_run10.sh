#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "cholesky or fit_2000 or fit_predict_deriv_golden or rbf_scenario" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/fit_plain.log
for cfg in "20000 3 1 0 2" "20000 2 1 0 1" "5000 3 1 0 2" "5000 2 1 0 1" "1000 3 1 0 2" "1000 2 1 0 1"; do
  timeout 200 python tools/fit_profile.py $cfg >> gpurun_out/fit_plain.log 2>&1 || echo "FAILED/TIMEOUT: $cfg rc=$?" >> gpurun_out/fit_plain.log
done
tail -30 gpurun_out/pytest_gpu.log | cut -c1-250; cat gpurun_out/fit_plain.log
