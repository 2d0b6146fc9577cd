#!/bin/bash
# round-2 call 1: steady-state snapshot of the freestyle workload, soak rates of the round-1 build, mid-game ncu profile of K5
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python tools/make_snapshot.py freestyle15 700 gpurun_out/steady_freestyle15.npz > gpurun_out/r02_snapshot_soak.txt 2>&1
tail -12 gpurun_out/r02_snapshot_soak.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_games -s 120 -c 1 -f -o gpurun_out/r02_k5_steady_before \
  python tools/profile_solver.py gpurun_out/steady_freestyle15.npz 60 > gpurun_out/r02_k5_profile_before.log 2>&1
tail -5 gpurun_out/r02_k5_profile_before.log
