#!/bin/bash
mkdir -p gpurun_out
for cfg in "2 56" "4 56" "6 56" "8 56" "6 48" "6 64" "8 64" "8 48"; do set -- $cfg; timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 $2 freestyle15 $1 2>&1 | tail -1; done | tee gpurun_out/r02_steady_green.txt
