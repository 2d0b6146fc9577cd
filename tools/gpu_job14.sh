#!/bin/bash
mkdir -p gpurun_out
AGB_SOLVER_WIDE=1 timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 56 freestyle15 2 2>&1 | tail -1 | tee gpurun_out/r02_steady_wide.txt
AGB_SOLVER_WIDE=1 timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 48 freestyle15 2 2>&1 | tail -1 | tee -a gpurun_out/r02_steady_wide.txt
timeout 2400 python -m pytest tests/test_host_gpu.py tests/test_selfplay_gpu.py -m gpu -x -q -s -k "host or benched or small_solver or generator" > gpurun_out/r02_pytest_gpu_d.log 2>&1; tail -12 gpurun_out/r02_pytest_gpu_d.log
