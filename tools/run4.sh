mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py -m gpu -x -q 2>&1 | tail -3
for cfg in "32 2" "16 4" "8 8" "4 8" "8 4"; do set -- $cfg; echo "stage_kb=$1 stages=$2"; AGB_NET_STAGE_KB=$1 AGB_NET_STAGES=$2 python tools/bench_forward.py 20 128 0 4096 3 2>&1 | tail -1; done | tee gpurun_out/bench_stages.log
AGB_NET_STAGE_KB=4 AGB_NET_STAGES=8 python tools/bench_forward.py 10 64 0 16384 3 2>&1 | tail -1 | tee -a gpurun_out/bench_stages.log
AGB_NET_STAGE_KB=8 AGB_NET_STAGES=8 python tools/bench_forward.py 10 64 0 16384 3 2>&1 | tail -1 | tee -a gpurun_out/bench_stages.log
