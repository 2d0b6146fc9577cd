#!/bin/bash
mkdir -p gpurun_out
for sms in 20 28 36 44 56; do timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 $sms 2>&1 | tail -1; done | tee gpurun_out/r02_steady_sms_sweep.txt
timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 0 freestyle15 1 2>&1 | tail -1 | tee -a gpurun_out/r02_steady_sms_sweep.txt
