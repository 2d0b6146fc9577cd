#!/bin/bash
mkdir -p gpurun_out
for cfg in "56 3" "56 4" "48 6" "56 6"; do set -- $cfg; AGB_GREEN_CONTEXTS=1 timeout 600 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 $1 freestyle15 $2 2>&1 | tail -2; done
