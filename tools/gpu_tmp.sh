#!/bin/bash
timeout 600 python -m pytest tests/test_resnet_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/bench_forward.py 20 128 0 4096 5 15 2>&1 | tail -1
for i in 1 2; do timeout 600 python tools/steady_bench.py bench_data/steady_freestyle15.npz 100 40 60 freestyle15 3 2>&1 | tail -1; done
