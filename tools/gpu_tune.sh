#!/bin/bash
# runs the named build/tune/k2x_<variant> binaries (tools/k2x_tune.cu builds of score_mma.cuh) on the bench shape
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/k2x_ep.txt
for v in "$@"; do
  echo "== $v" >> gpurun_out/k2x_ep.txt
  timeout 120 build/tune/k2x_$v 1048576 20000 5 >> gpurun_out/k2x_ep.txt 2>&1
done
cat gpurun_out/k2x_ep.txt
