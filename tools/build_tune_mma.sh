#!/bin/bash
set -e
cd "$(dirname "$0")/.."
mkdir -p build/tune
rm -f build/tune/k2_* build/tune/k2x_*
b() { nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DB2_MMA_MB=$2 -DB2_MMA_MINB=$3 $4 tools/k2x_tune.cu -o build/tune/k2x_$1 & }
b mb2_b3 2 3; b mb2_b4 2 4; b mb2_b2 2 2; b mb4_b2 4 2; b mb1_b4 1 4; b mb1_b6 1 6; b mb3_b2 3 2; b mb3_b3 3 3
wait
b mb2_b3_d10 2 3 -DTUNE_D=10; b mb2_b3_d30 2 3 -DTUNE_D=30; b mb4_b2_d10 4 2 -DTUNE_D=10; b mb2_b5_d10 2 5 -DTUNE_D=10
wait
ls build/tune
