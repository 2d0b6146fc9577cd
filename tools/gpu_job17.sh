#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 56 freestyle15 2 2>&1 | tail -1 | tee gpurun_out/r02_steady_plain.txt
timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 48 freestyle15 2 2>&1 | tail -1 | tee -a gpurun_out/r02_steady_plain.txt
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_e.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_e.log
