#!/bin/bash
mkdir -p gpurun_out
for cfg in "6 56 28" "6 48 28" "4 56 28" "8 56 28" "8 48 28" "6 56 16" "6 64 28"; do set -- $cfg; echo -n "resident $3: "; AGB_SOLVER_RESIDENT=$3 timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 $2 freestyle15 $1 2>&1 | tail -1; done | tee gpurun_out/r02_steady_green4.txt
rm -f gpurun_out/r02_trace_g6.txt; AGB_STEP_TRACE=gpurun_out/r02_trace_g6.txt timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 20 56 freestyle15 6 2>&1 | tail -1
tail -36 gpurun_out/r02_trace_g6.txt > gpurun_out/r02_trace_g6_tail.txt
