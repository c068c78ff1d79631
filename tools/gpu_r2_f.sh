#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large_shapes.py tests/test_gpu_incremental.py -q -x -m gpu --timeout 600 -p no:cacheprovider -k "fit or cholesky or config3 or golden or extension or interpolates" > gpurun_out/r2_f_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2_f_tests.log
for n in 1000 2000 5000 10000 20000; do timeout 300 python tools/fit_profile.py $n 3 0 0 0 2>&1 | grep -E "^fit|repeat" >> gpurun_out/r2_fit_profile.log; done
tail -4 gpurun_out/r2_f_tests.log | cut -c1-400; cat gpurun_out/r2_fit_profile.log
