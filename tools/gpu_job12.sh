#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_selfplay_gpu.py -m gpu -x -q -k "not benched and not small_solver" > gpurun_out/r02_pytest_gpu_c.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_c.log
for cfg in "2 56" "2 48" "2 40" "2 64"; do set -- $cfg; timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 $2 freestyle15 $1 2>&1 | tail -1; done | tee gpurun_out/r02_steady_smem.txt
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_active,l1tex__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio --clock-control none -k regex:solve_games -s 120 -c 1 --csv --log-file gpurun_out/r02_k5_inst_b.csv python tools/profile_solver.py bench_data/steady_freestyle15.npz 60 > /dev/null 2>&1
tail -6 gpurun_out/r02_k5_inst_b.csv | cut -d, -f12-
