#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -s -m gpu --timeout 600 -p no:cacheprovider -k "gradient or deriv or dimension_classes" > gpurun_out/r2_e_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2_e_tests.log
timeout 600 python tools/deriv_bench.py > gpurun_out/r2_deriv_bench.log 2>&1
grep -E "tensor path|passed|failed|rc=|Error|error" gpurun_out/r2_e_tests.log | cut -c1-300 | tail -30; cat gpurun_out/r2_deriv_bench.log
