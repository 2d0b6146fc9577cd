#!/bin/bash
mkdir -p gpurun_out
for cfg in "60 3" "64 3" "52 3" "60 4" "68 4"; do set -- $cfg; timeout 600 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 $1 freestyle15 $2 2>&1 | tail -1; done
