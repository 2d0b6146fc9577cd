#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_b.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_b.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err; tail -c 3000 gpurun_out/r02_bench_n1_a.json; tail -5 gpurun_out/r02_bench_n1_a.err
