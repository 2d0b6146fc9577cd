#!/bin/bash
mkdir -p gpurun_out
AGB_VERBOSE=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; tail -c 600 gpurun_out/r02_bench_n1_b.json; tail -5 gpurun_out/r02_bench_n1_b.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_b.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_b.log
