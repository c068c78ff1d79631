#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/fit_profile.py 20000 3 0 0 0 > gpurun_out/fit_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_fit_chol2.csv python tools/fit_profile.py 20000 1 0 0 0 > gpurun_out/ncu_fit.log 2>&1
tail -4 gpurun_out/fit_plain.log; tail -2 gpurun_out/ncu_fit.log | cut -c1-200
