#!/bin/bash
# evidence run: one ncu --set full pass over one steady-state launch of each engine kernel, the launch list of a short bench, and a 1000-step soak
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"resnet_board_kernel|select_kernel|expand_backup_kernel|make_move_kernel|solve_games_kernel|set_boards_kernel" -s 720 -c 6 -f -o gpurun_out/r02_step_kernels \
  python tools/profile_solver.py bench_data/steady_freestyle15.npz 60 > gpurun_out/r02_step_kernels.log 2>&1
tail -3 gpurun_out/r02_step_kernels.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --settle 20 --no-cpu-baseline --no-early-game > /dev/null 2>&1
tail -2 gpurun_out/r02_bench_launches.csv | cut -c1-300
timeout 900 python tools/make_snapshot.py freestyle15 1000 gpurun_out/steady_freestyle15_1000.npz > gpurun_out/r02_soak_1000.txt 2>&1; tail -4 gpurun_out/r02_soak_1000.txt
timeout 600 python -m pytest tests/test_selfplay_gpu.py -m gpu -x -q -k "save_and_resume or symmetr or shard" 2>&1 | tail -3
