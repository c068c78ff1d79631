#!/bin/bash
# fit-focused check: the fit / Cholesky parity tests, N = 20000 timing, ncu launch list of an N = 5000 fit
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "cholesky or fit or golden or scale or capped or interpolates" > gpurun_out/pytest_fit.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_fit.log
timeout 300 python tools/fit_profile.py 20000 3 0 0 0 2>&1 | tail -3 > gpurun_out/fit_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_fit5k.csv python tools/fit_profile.py 5000 1 0 0 0 > gpurun_out/ncu_fit.log 2>&1
tail -3 gpurun_out/pytest_fit.log | cut -c1-300; cat gpurun_out/fit_plain.log | cut -c1-200
