#!/bin/bash
# 2-GPU call: sharded correctness, reference arm, then 1-GPU ncu captures of the final kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_sharded.py > gpurun_out/check_sharded.log 2>&1
echo "check_sharded rc=$?" >> gpurun_out/check_sharded.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
echo "ref rc=$?" >> gpurun_out/bench_ref.log
export CUDA_VISIBLE_DEVICES=0
timeout 600 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/b_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_mma -s 1 -c 1 -o gpurun_out/prof_score_mma python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu_f.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:score|merit|pick|stats|update|take|negate' -c 120 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/ncu_l.log 2>&1
tail -8 gpurun_out/check_sharded.log | cut -c1-250; tail -2 gpurun_out/bench_ref.log | cut -c1-700; tail -2 gpurun_out/ncu_f.log
