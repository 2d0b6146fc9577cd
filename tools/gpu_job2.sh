#!/bin/bash
# K5 step 1 (HALF_OPEN_3 lists dropped from the search, lane-split add/undo): parity tests, then steady-state rates
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_a.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_a.log
for sms in 74 0; do timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 $sms 2>&1 | tail -1; done | tee gpurun_out/r02_steady_a.txt
