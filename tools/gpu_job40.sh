#!/bin/bash
echo "constexpr 4: $(timeout 300 python tools/bench_forward.py 20 128 0 4096 5 20 2>&1 | tail -1)"
echo "16x16: $(timeout 300 python tools/bench_forward.py 20 128 0 4096 5 16 2>&1 | tail -1)"
timeout 600 python -m pytest tests/test_resnet_gpu.py -m gpu -x -q 2>&1 | tail -2
