#!/bin/bash
# run with gpurun --gpus N
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -q -x -m gpu --timeout 800 -p no:cacheprovider > gpurun_out/r2_multi_tests_$N.log 2>&1
echo "rc=$?" >> gpurun_out/r2_multi_tests_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_$N.log 2> gpurun_out/r2_bench_$N.err
echo "rc=$?" >> gpurun_out/r2_bench_$N.err
tail -8 gpurun_out/r2_multi_tests_$N.log | cut -c1-1200; tail -5 gpurun_out/r2_bench_$N.err | cut -c1-500; cut -c1-2500 gpurun_out/r2_bench_$N.log
