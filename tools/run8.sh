mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_forward.py 20 128 0 4096 3 2>&1 | tail -1 | tee gpurun_out/bench_forward.log
AGB_NET_TRACE=gpurun_out/trace.txt python tools/bench_forward.py 20 128 0 4096 1 2>&1 | tail -1
python tools/bench_forward.py 10 64 0 16384 3 2>&1 | tail -1 | tee -a gpurun_out/bench_forward.log
python tools/bench_forward.py 20 128 1 4096 3 2>&1 | tail -1 | tee -a gpurun_out/bench_forward.log
