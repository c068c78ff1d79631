#!/bin/bash
# round 2 profiling call: launch lists + full captures, each after the same command exited 0 without ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU_L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
NCU_F="ncu --set full --clock-control none --import-source on -f"
# 1. bench launch list (share of the step per kernel)
timeout 600 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/n_bench_plain.log 2>&1 &&
timeout 900 $NCU_L -k 'regex:^(score_|merit_|pick_|best_|decide_|stats_|update_kernel|take_|negate_)' -c 200 --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/n_bench_ncu.log 2>&1
# 2. the headline kernel, full capture (roofline.traffic)
timeout 900 $NCU_F -k regex:score_mma -s 1 -c 1 -o gpurun_out/r02_prof_score_mma python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/n_score_full.log 2>&1
# 3. the gradient kernel on the tensor path
timeout 300 python tools/deriv_one.py tps 10 20000 20 1 > gpurun_out/n_deriv_plain.log 2>&1 &&
timeout 900 $NCU_F -k regex:deriv_mma -s 2 -c 1 -o gpurun_out/r02_prof_deriv_mma python tools/deriv_one.py tps 10 20000 20 1 > gpurun_out/n_deriv_full.log 2>&1
timeout 300 python tools/deriv_one.py cubic 50 20000 19 1 > gpurun_out/n_deriv50_plain.log 2>&1 &&
timeout 900 $NCU_F -k regex:deriv_mma -s 2 -c 1 -o gpurun_out/r02_prof_deriv_mma_d50 python tools/deriv_one.py cubic 50 20000 19 1 > gpurun_out/n_deriv50_full.log 2>&1
# 4. the shipped DMMA GEMM variant <2,32> at K = 512 and K = 128
timeout 300 python tools/gemm_profile.py 12288 512 4 > gpurun_out/n_gemm512_plain.log 2>&1 &&
timeout 900 $NCU_F -k regex:gemm_sub -s 2 -c 1 -o gpurun_out/r02_prof_gemm_k512 python tools/gemm_profile.py 12288 512 4 > gpurun_out/n_gemm512_full.log 2>&1
timeout 300 python tools/gemm_profile.py 12288 128 4 > gpurun_out/n_gemm128_plain.log 2>&1 &&
timeout 900 $NCU_F -k regex:gemm_sub -s 2 -c 1 -o gpurun_out/r02_prof_gemm_k128 python tools/gemm_profile.py 12288 128 4 > gpurun_out/n_gemm128_full.log 2>&1
# 5. fit launch lists: N = 20000 and N = 1000 full fits, and the extension at n ~ 1000
timeout 300 python tools/fit_profile.py 20000 1 0 0 0 > gpurun_out/n_fit20k_plain.log 2>&1 &&
timeout 900 $NCU_L -c 6000 --log-file gpurun_out/r02_launches_fit20k.csv python tools/fit_profile.py 20000 1 0 0 0 > gpurun_out/n_fit20k_ncu.log 2>&1
timeout 300 python tools/fit_profile.py 1000 2 0 0 0 > gpurun_out/n_fit1k_plain.log 2>&1 &&
timeout 900 $NCU_L -c 6000 --log-file gpurun_out/r02_launches_fit1k.csv python tools/fit_profile.py 1000 2 0 0 0 > gpurun_out/n_fit1k_ncu.log 2>&1
timeout 300 python tools/ext_profile.py 30 900 20 > gpurun_out/n_ext_plain.log 2>&1 &&
timeout 900 $NCU_L -c 3000 --log-file gpurun_out/r02_launches_ext.csv python tools/ext_profile.py 30 900 20 > gpurun_out/n_ext_ncu.log 2>&1
# 6. the pipelined triangular solve at N = 20000 (inside the fit), full capture
timeout 900 $NCU_F -k regex:trsv_pipe -c 2 -o gpurun_out/r02_prof_trsv python tools/fit_profile.py 20000 1 0 0 0 > gpurun_out/n_trsv_full.log 2>&1
# live SOP run, JSON on record
timeout 600 python tools/live_strategy_run.py --config 5 > gpurun_out/r2_live_sop3000.json 2> gpurun_out/r2_live_sop3000.err
ls -la gpurun_out/*.ncu-rep | tail; tail -1 gpurun_out/n_deriv_plain.log gpurun_out/n_deriv50_plain.log gpurun_out/n_gemm512_plain.log gpurun_out/n_gemm128_plain.log gpurun_out/n_ext_plain.log; cut -c1-300 gpurun_out/r2_live_sop3000.json
