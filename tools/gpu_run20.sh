#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in base eb_s2 eb_s3 eb_s4 base eb_s3; do
  echo "== $v" >> gpurun_out/k2x_eb.txt
  timeout 120 build/tune/k2x_$v 1048576 20000 5 >> gpurun_out/k2x_eb.txt 2>&1
done
timeout 900 python -m pytest tests/test_gpu_strategy_traces.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_traces.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_traces.log
timeout 1200 python tools/replay_timing.py > gpurun_out/replay_timing.log 2>&1
cat gpurun_out/k2x_eb.txt; tail -5 gpurun_out/pytest_traces.log | cut -c1-300; tail -8 gpurun_out/replay_timing.log | cut -c1-400
