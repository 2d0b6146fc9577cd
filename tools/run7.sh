mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err.log | tee gpurun_out/bench.json
tail -5 gpurun_out/bench_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --games 4096 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
