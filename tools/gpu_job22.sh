#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_solver_gpu.py tests/test_selfplay_gpu.py -m gpu -x -q -k "not benched and not small_solver" 2>&1 | tail -3
for sms in 56 48; do timeout 600 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 $sms freestyle15 2 2>&1 | tail -1; done
