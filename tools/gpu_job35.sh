#!/bin/bash
timeout 900 python -m pytest tests/test_resnet_gpu.py tests/test_host_gpu.py -m gpu -x -q 2>&1 | tail -4
