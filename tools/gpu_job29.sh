#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/bench_forward.py 20 128 1 4096 5 15 2>&1 | tail -1
for cfg in "56 3" "60 3" "48 2" "0 0"; do set -- $cfg; timeout 600 python tools/steady_bench.py bench_data/steady_freestyle15.npz 100 40 $1 freestyle15 $2 2>&1 | tail -1; done
