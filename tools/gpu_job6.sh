#!/bin/bash
mkdir -p gpurun_out
for cfg in "6 64 28" "6 64 16" "6 64 12" "6 64 8" "6 72 12" "8 72 12" "4 64 12" "8 64 16"; do set -- $cfg; echo -n "resident $3: "; AGB_SOLVER_RESIDENT=$3 timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 $2 freestyle15 $1 2>&1 | tail -1; done | tee gpurun_out/r02_steady_green2.txt
rm -f gpurun_out/r02_trace_g6.txt; AGB_SOLVER_RESIDENT=12 AGB_STEP_TRACE=gpurun_out/r02_trace_g6.txt timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 20 64 freestyle15 6 2>&1 | tail -1
tail -60 gpurun_out/r02_trace_g6.txt > gpurun_out/r02_trace_g6_tail.txt
