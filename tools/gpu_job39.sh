#!/bin/bash
mkdir -p gpurun_out
for w in standard15 renju15 caro20; do
timeout 900 python bench.py --workload $w --start openings --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_$w.json 2> gpurun_out/r02_bench_n1_$w.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n1_$w.json').read().strip().splitlines()[-1])
    print('$w value',round(d['value']),'ms/step',round(d['ms_per_step'],1),d['config']['pipeline'][:70],'overflow',d['overflow_flags'])
except Exception as e:
    print('$w failed',e); print(open('gpurun_out/r02_bench_n1_$w.err').read()[-800:])
PY
done
