#!/bin/bash
mkdir -p gpurun_out
AGB_NET_TRACE=gpurun_out/r02_k4_layer_timeline_pipelined.txt timeout 300 python tools/bench_forward.py 20 128 1 148 1 15 2>&1 | tail -1
timeout 300 python tools/bench_forward.py 20 128 1 4096 5 15 2>&1 | tail -1
for sms in 56 52; do timeout 600 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 $sms freestyle15 2 2>&1 | tail -1; done
