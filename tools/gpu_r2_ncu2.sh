#!/bin/bash
# round 2, second profiling call: the int8 headline kernel and the probes behind DESIGN.md's K2i section
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU_L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
NCU_F="ncu --set full --clock-control none --import-source on -f"
timeout 600 python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/n2_bench_plain.log 2>&1 &&
timeout 900 $NCU_L -k 'regex:^(score_|merit_|pick_|best_|decide_|stats_|update_|take_|negate_|i8_)' -c 200 --log-file gpurun_out/r02_launches_bench_i8.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/n2_bench_ncu.log 2>&1
timeout 900 $NCU_F -k regex:score_i8 -s 1 -c 1 -o gpurun_out/r02_prof_score_i8_final python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/n2_score_full.log 2>&1
( echo "== tools/umma_probe.cu"; timeout 60 build/tune/umma_probe; echo "== tools/umma_rate2.cu"; timeout 60 build/tune/umma_rate2; echo "== tools/umma_overlap.cu"; timeout 60 build/tune/umma_overlap; echo "== tools/i8_sweep.py"; timeout 600 python tools/i8_sweep.py ) > gpurun_out/r02_i8_probes.txt 2>&1
timeout 600 python tools/replay_timing.py --no-oracle > gpurun_out/r2_replay_timing_final.log 2>&1
tail -1 gpurun_out/n2_bench_plain.log | cut -c1-300; ls -la gpurun_out/r02_prof_score_i8_final.ncu-rep; tail -25 gpurun_out/r02_i8_probes.txt | cut -c1-160; tail -3 gpurun_out/r2_replay_timing_final.log | cut -c1-300
