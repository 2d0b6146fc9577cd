#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_host_gpu.py tests/test_selfplay_gpu.py -m gpu -x -q -s -k "host or benched or small_solver or generator" > gpurun_out/r02_pytest_gpu_d.log 2>&1; tail -25 gpurun_out/r02_pytest_gpu_d.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err; tail -c 300 gpurun_out/r02_bench_n1_d.json; tail -3 gpurun_out/r02_bench_n1_d.err
