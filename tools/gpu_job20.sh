#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_g.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_g.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_e.json 2> gpurun_out/r02_bench_n1_e.err; tail -c 200 gpurun_out/r02_bench_n1_e.json; tail -3 gpurun_out/r02_bench_n1_e.err
