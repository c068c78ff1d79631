#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "cholesky or fit or golden or scale or capped or interpolates" > gpurun_out/pytest_fit.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_fit.log
for n in 20000 5000 1000 300; do
timeout 300 python tools/fit_profile.py $n 3 0 0 0 2>&1 | tail -3 >> gpurun_out/fit_plain.log
B2_OLD_POTRF=1 timeout 300 python tools/fit_profile.py $n 3 0 0 0 2>&1 | tail -2 | sed 's/^/OLD /' >> gpurun_out/fit_plain.log
done
tail -5 gpurun_out/pytest_fit.log | cut -c1-300; grep -v "^$" gpurun_out/fit_plain.log | cut -c1-200
