#!/bin/bash
# Builds the k2_tune variants into build/tune/ (git-ignored; travels to the GPU box with gpurun).
set -e
cd "$(dirname "$0")/.."
mkdir -p build/tune
build() {  # name threads minb jt tile stages [extra]
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
       -DB2_SCORE_THREADS=$2 -DB2_SCORE_MINB=$3 -DB2_JT=$4 -DB2_TILE_ROWS=$5 -DB2_STAGES=$6 $7 \
       tools/k2_tune.cu -o build/tune/k2_$1 &
}
build t256_b1_j4   256 1 4 32 3
build t128_b3_j4   128 3 4 32 3
build t128_b4_j4   128 4 4 32 3
build t256_b2_j4   256 2 4 32 3
build t256_b1_j8   256 1 8 32 3
build t128_b3_j8   128 3 8 32 3
build t128_b2_j8   128 2 8 32 3
build t192_b2_j4   192 2 4 32 3
wait
build t256_b1_j2   256 1 2 32 3
build t128_b3_j2   128 3 2 32 3
build t512_b1_j4   512 1 4 32 3
build t64_b6_j4    64  6 4 32 3
build t128_b3_j4_exact 128 3 4 32 3 -DTUNE_EXACT=true
build t128_b3_j4_tps   128 3 4 32 3 -DTUNE_KERN=1
build t128_b3_j4_d10   128 3 4 32 3 -DTUNE_DPAD=10
build t256_b1_j4_t64   256 1 4 64 2
wait
ls build/tune
