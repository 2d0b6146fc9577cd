#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resnet_gpu.py -m gpu -x -q 2>&1 | tail -8
for cfg in "20 128 1 4096 5 15" "10 64 0 8192 5 15" "20 128 1 2048 5 20"; do timeout 300 python tools/bench_forward.py $cfg 2>&1 | tail -1; done
AGB_NET_TRACE=gpurun_out/r02_k4_layer_timeline_pipelined.txt timeout 300 python tools/bench_forward.py 20 128 1 148 1 15 2>&1 | tail -1
