#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_ab2.log
: > $L
timeout 300 python tools/i8_one.py 20 0 >> $L 2>&1
timeout 300 python tools/call_latency2.py >> $L 2>&1
timeout 600 python tools/replay_probe.py >> $L 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large_shapes.py tests/test_gpu_strategy_traces.py tests/test_gpu_int8.py tests/test_gpu_sharded.py -q -x -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/r2_ab2_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2_ab2_tests.log
timeout 900 python tools/replay_timing.py --no-oracle > gpurun_out/r2_replay_timing.log 2>&1
cat $L | cut -c1-300; tail -5 gpurun_out/r2_ab2_tests.log | cut -c1-600; tail -6 gpurun_out/r2_replay_timing.log
