#!/bin/bash
mkdir -p gpurun_out
for n in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n${n}_final.json 2> gpurun_out/r02_bench_n${n}_final.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n${n}_final.json').read().strip().splitlines()[-1])
print('N=$n value',round(d['value']),'ms/step',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value']),d['records'],d['sharding']['per_rank_ms_per_step'])
PY
done
