#!/bin/bash
mkdir -p gpurun_out
for cfg in "3 40" "3 56" "4 40" "4 56" "4 72"; do set -- $cfg; timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 $2 freestyle15 $1 2>&1 | tail -1; done | tee gpurun_out/r02_steady_groups.txt
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:solve_games -s 120 -c 2 --csv --log-file gpurun_out/r02_k5_inst_a.csv python tools/profile_solver.py bench_data/steady_freestyle15.npz 60 > /dev/null 2>&1
tail -4 gpurun_out/r02_k5_inst_a.csv
