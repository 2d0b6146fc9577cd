#!/bin/bash
mkdir -p gpurun_out
for cfg in "6 64 22" "6 56 22" "4 56 22" "8 56 22" "6 56 24" "6 56 16" "6 48 22"; do set -- $cfg; echo -n "resident $3: "; AGB_SOLVER_RESIDENT=$3 timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 $2 freestyle15 $1 2>&1 | tail -1; done | tee gpurun_out/r02_steady_green3.txt
echo -n "per-group nn streams: "; AGB_NET_PER_GROUP=1 timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 56 freestyle15 6 2>&1 | tail -1 | tee -a gpurun_out/r02_steady_green3.txt
rm -f gpurun_out/r02_trace_g6.txt; AGB_STEP_TRACE=gpurun_out/r02_trace_g6.txt timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 20 56 freestyle15 6 2>&1 | tail -1
tail -36 gpurun_out/r02_trace_g6.txt > gpurun_out/r02_trace_g6_tail.txt
