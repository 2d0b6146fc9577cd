mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json
tail -3 gpurun_out/bench_err.log
ncu --set full --clock-control none --import-source on -k regex:resnet_board -c 1 -o gpurun_out/prof_resnet_pair python tools/bench_forward.py 20 128 0 1184 1 > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
