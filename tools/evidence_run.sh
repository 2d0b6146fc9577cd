#!/bin/bash
# evidence run: ncu --set full over one steady-state launch of each engine kernel and the launch list of a short bench
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"resnet_board_kernel|select_kernel|expand_backup_kernel|make_move_kernel|solve_games_kernel|set_boards_kernel" -s 1098 -c 6 -f -o gpurun_out/r02_step_kernels_final \
  python tools/profile_solver.py bench_data/steady_freestyle15.npz 60 > gpurun_out/r02_step_kernels_final.log 2>&1
tail -2 gpurun_out/r02_step_kernels_final.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_bench_launches_final.csv python bench.py --steps 2 --warmup 3 --settle 20 --no-cpu-baseline --no-early-game > /dev/null 2>&1
tail -1 gpurun_out/r02_bench_launches_final.csv | cut -c1-200
