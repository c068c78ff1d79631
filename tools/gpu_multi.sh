#!/bin/bash
# multi-GPU check: N = $1 ranks
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/check_sharded.py > gpurun_out/check_sharded_$N.log 2>&1
echo "check_sharded rc=$?" >> gpurun_out/check_sharded_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_$N.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench_$N.log
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 500 -p no:cacheprovider > gpurun_out/pytest_sharded_$N.log 2>&1
tail -4 gpurun_out/check_sharded_$N.log | cut -c1-250; tail -2 gpurun_out/bench_$N.log | cut -c1-600; tail -2 gpurun_out/pytest_sharded_$N.log
