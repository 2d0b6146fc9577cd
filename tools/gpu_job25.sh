#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_games -s 120 -c 1 -f -o gpurun_out/r02_k5_steady_c python tools/profile_solver.py bench_data/steady_freestyle15.npz 60 > gpurun_out/r02_k5_profile_c.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c.json 2>gpurun_out/r02_bench_c.err; tail -c 2500 gpurun_out/r02_bench_c.json
