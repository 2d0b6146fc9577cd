set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_resnet.log
python tools/bench_forward.py 20 128 0 4096 2>&1 | tail -2 | tee gpurun_out/bench_forward.log
python tools/bench_forward.py 20 128 0 16384 2>&1 | tail -1 | tee -a gpurun_out/bench_forward.log
python tools/bench_forward.py 10 64 0 16384 2>&1 | tail -1 | tee -a gpurun_out/bench_forward.log
ncu --set full --clock-control none --import-source on -k regex:resnet_board -c 1 -o gpurun_out/prof_resnet python tools/bench_forward.py 20 128 0 1184 1 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
