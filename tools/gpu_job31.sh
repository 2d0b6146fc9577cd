#!/bin/bash
timeout 600 python tools/k4_determinism.py 2>&1 | tail -30
timeout 900 python -m pytest tests/test_host_gpu.py -m gpu -x -q 2>&1 | grep -E "AssertionError|assert|games|passed|failed" | head -20
