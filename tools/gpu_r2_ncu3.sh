#!/bin/bash
# round 2, final profiling call (after the accuracy-guard rewrite and the center split): launch list of the bench step and
# full captures of the two score kernels at the bench shape, the DMMA kernel at a strategy-sized call, the gradient kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU_L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
NCU_F="ncu --set full --clock-control none --import-source on -f"
timeout 600 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/n3_bench_plain.log 2>&1 &&
timeout 900 $NCU_L -k 'regex:^(score_|merit_|pick_|best_|decide_|stats_|update_|take_|negate_|i8_)' -c 200 --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/n3_bench_ncu.log 2>&1
timeout 900 $NCU_F -k regex:score_i8 -s 1 -c 1 -o gpurun_out/r02_prof_score_i8_final2 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/n3_score_full.log 2>&1
timeout 300 python tools/i8_one.py 20 0 > gpurun_out/n3_k2x_plain.log 2>&1 &&
timeout 900 $NCU_F -k regex:score_mma -s 2 -c 1 -o gpurun_out/r02_prof_score_mma_final2 python tools/i8_one.py 20 0 > gpurun_out/n3_k2x_full.log 2>&1
timeout 300 python tools/small_call_one.py > gpurun_out/n3_small_plain.log 2>&1 &&
timeout 900 $NCU_F -k regex:score_mma -s 4 -c 2 -o gpurun_out/r02_prof_score_mma_small python tools/small_call_one.py > gpurun_out/n3_small_full.log 2>&1
timeout 300 python tools/deriv_one.py tps 10 20000 20 1 > gpurun_out/n3_deriv_plain.log 2>&1 &&
timeout 900 $NCU_F -k regex:deriv_mma -s 2 -c 1 -o gpurun_out/r02_prof_deriv_mma_final2 python tools/deriv_one.py tps 10 20000 20 1 > gpurun_out/n3_deriv_full.log 2>&1
ls -la gpurun_out/*final2*.ncu-rep gpurun_out/r02_prof_score_mma_small.ncu-rep; tail -1 gpurun_out/n3_bench_plain.log | cut -c1-200; cat gpurun_out/n3_k2x_plain.log gpurun_out/n3_small_plain.log gpurun_out/n3_deriv_plain.log
