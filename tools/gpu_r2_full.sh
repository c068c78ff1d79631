#!/bin/bash
# the round-end sequence: whole GPU suite, smoke, bench (both arms)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q --timeout 1500 -p no:cacheprovider --durations=8 > gpurun_out/r2_full_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2_full_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/r2_smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.log 2>&1
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench.err
echo "rc=$?" >> gpurun_out/r2_bench.err
tail -16 gpurun_out/r2_full_pytest.log | cut -c1-300; tail -3 gpurun_out/r2_smoke.log; cut -c1-400 gpurun_out/r2_bench_ref.log; tail -2 gpurun_out/r2_bench.err; cut -c1-1200 gpurun_out/r2_bench.log
