#!/bin/bash
# same-box A/B of two builds of the library (build/ab/head.so = last commit, build/ab/new.so = working tree)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_ab.log
: > $L
for rep in 1; do
for so in head new; do
  echo "== $so" >> $L
  B200RBF_LIB=$PWD/build/ab/$so.so timeout 300 python tools/i8_one.py 20 0 >> $L 2>&1
  B200RBF_LIB=$PWD/build/ab/$so.so timeout 300 python tools/i8_one.py 20 1 >> $L 2>&1
  B200RBF_LIB=$PWD/build/ab/$so.so timeout 300 python tools/deriv_one.py tps 10 20000 20 1 >> $L 2>&1
done
done
for so in head new; do
  echo "== $so replay" >> $L
  B200RBF_LIB=$PWD/build/ab/$so.so timeout 600 python tools/replay_probe.py >> $L 2>&1
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large_shapes.py tests/test_gpu_strategy_traces.py tests/test_gpu_int8.py -q -x -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/r2_ab_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2_ab_tests.log
cat $L | cut -c1-300; tail -5 gpurun_out/r2_ab_tests.log | cut -c1-600
