#!/bin/bash
mkdir -p gpurun_out
for cfg in "56 3" "60 3" "64 3" "0 0"; do set -- $cfg; timeout 600 python tools/steady_bench.py bench_data/steady_freestyle15.npz 100 40 $1 freestyle15 $2 2>&1 | tail -1; done
