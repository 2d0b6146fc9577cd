#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_games -s 120 -c 1 -f -o gpurun_out/r02_k5_steady_after python tools/profile_solver.py bench_data/steady_freestyle15.npz 60 > gpurun_out/r02_k5_profile_after.log 2>&1
tail -3 gpurun_out/r02_k5_profile_after.log
