#!/bin/bash
set -e
cd "$(dirname "$0")/.."
mkdir -p build/tune
rm -f build/tune/k2_* build/tune/k2x_*
b() { nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DB2_MMA_MB=$2 -DB2_MMA_MINB=$3 -DB2_MMA_JB=$4 -DB2_MMA_STAGES=$5 $6 tools/k2x_tune.cu -o build/tune/k2x_$1 & }
b mb2_b4_j4_s3 2 4 4 3; b mb2_b4_j2_s3 2 4 2 3; b mb2_b4_j8_s2 2 4 8 2; b mb2_b4_j4_s2 2 4 4 2; b mb2_b5_j4_s3 2 5 4 3; b mb1_b4_j8_s2 1 4 8 2; b mb1_b5_j4_s3 1 5 4 3; b mb1_b6_j4_s3 1 6 4 3
wait
b mb2_b4_j4_s3_d10 2 4 4 3 -DTUNE_D=10; b mb2_b4_j4_s3_d30 2 4 4 3 -DTUNE_D=30; b mb2_b4_j4_s3_d20 2 4 4 3 -DTUNE_D=20; b mb2_b4_j4_s3_d64 2 4 4 3 -DTUNE_D=64
wait
ls build/tune | wc -l
