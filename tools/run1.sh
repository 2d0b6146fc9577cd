set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
mkdir -p gpurun_out
for args in "0 0 128 0" "1 0 128 0" "0 1 128 0" "0 17 128 0" "0 0 128 1" "0 3 256 1" "0 0 64 0"; do timeout 60 ./tools/umma_probe $args >> gpurun_out/umma_probe.log 2>&1; echo "exit $?" >> gpurun_out/umma_probe.log; done
cat gpurun_out/umma_probe.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
