#!/bin/bash
# round 2, first GPU call: the new parity tests at the benchmarked shapes, the live strategy runs, then the whole suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_large_shapes.py -q -s -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2_large.log 2>&1
echo "rc=$?" >> gpurun_out/r2_large.log
timeout 1500 python -m pytest tests/test_gpu_live_strategies.py -q -s -m gpu --timeout 1400 -p no:cacheprovider -k "not config5_sop_async8 and not config2_srbf_async8_history" > gpurun_out/r2_live.log 2>&1
echo "rc=$?" >> gpurun_out/r2_live.log
for c in 1 2; do timeout 600 python tools/live_strategy_run.py --config $c >> gpurun_out/r2_live_json.log 2>&1; done
timeout 600 python tools/live_strategy_run.py --config 2 --controller thread >> gpurun_out/r2_live_json.log 2>&1
timeout 900 python tools/live_strategy_run.py --config 5 --evals 600 >> gpurun_out/r2_live_json.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider --deselect tests/test_gpu_large_shapes.py --deselect tests/test_gpu_live_strategies.py > gpurun_out/r2_pytest_rest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_pytest_rest.log
tail -25 gpurun_out/r2_large.log | cut -c1-400; tail -15 gpurun_out/r2_live.log | cut -c1-600; cat gpurun_out/r2_live_json.log | cut -c1-900; tail -8 gpurun_out/r2_pytest_rest.log | cut -c1-400
