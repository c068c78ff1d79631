#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_incremental.py tests/test_gpu_parity.py tests/test_gpu_large_shapes.py -q -x -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/r2_c_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2_c_tests.log
for a in "10 400 100" "30 900 100" "20 2500 100" "30 5000 60"; do timeout 300 python tools/ext_profile.py $a >> gpurun_out/r2_ext_profile.log 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_ext.csv python tools/ext_profile.py 30 900 20 > gpurun_out/ncu_ext.log 2>&1
timeout 900 python tools/replay_timing.py --no-oracle > gpurun_out/r2_replay_timing.log 2>&1
tail -5 gpurun_out/r2_c_tests.log | cut -c1-600; cat gpurun_out/r2_ext_profile.log; tail -4 gpurun_out/r2_replay_timing.log
