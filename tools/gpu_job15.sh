#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_host_gpu.py tests/test_selfplay_gpu.py -m gpu -x -q -s -k "device_engine or benched or small_solver" > gpurun_out/r02_pytest_gpu_d.log 2>&1; tail -12 gpurun_out/r02_pytest_gpu_d.log
