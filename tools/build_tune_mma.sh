#!/bin/bash
# builds the tools/k2x_tune.cu variants of the DMMA score kernel under build/tune/ (run them with tools/gpu_tune.sh):
#   k2x_<name>: -DB2_MMA_MINB (CTAs/SM the registers are capped for) -DB2_MMA_JB (n8 blocks per step; 8 needs
#   -DB2_PAD_ROWS=64) -DB2_MMA_STAGES -DTUNE_D -DTUNE_KERN (0 cubic, 1 TPS, 2 linear) -DTUNE_AUG (norms inside the MMA)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/tune
b() { n=$1; shift; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo "$@" tools/k2x_tune.cu -o build/tune/k2x_$n & }
b cur -DTUNE_AUG=true
b b4 -DTUNE_AUG=true -DB2_MMA_MINB=5          # 4 CTAs/SM at d = 50 (the kernel takes MINB - 1 when d >= 45)
b s2 -DTUNE_AUG=true -DB2_MMA_STAGES=2
b j8 -DTUNE_AUG=true -DB2_MMA_JB=8 -DB2_PAD_ROWS=64
b noaug -DTUNE_AUG=false
b tps50 -DTUNE_AUG=true -DTUNE_KERN=1
b tps10 -DTUNE_AUG=true -DTUNE_KERN=1 -DTUNE_D=10
b d10 -DTUNE_AUG=true -DTUNE_D=10
wait
b d20 -DTUNE_AUG=false -DTUNE_D=20
b d30 -DTUNE_AUG=true -DTUNE_D=30
b d64 -DTUNE_AUG=false -DTUNE_D=64
wait
ls build/tune | grep -c k2x_
