mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_selfplay_gpu.py -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_selfplay.log
