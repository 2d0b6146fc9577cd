#!/bin/bash
mkdir -p gpurun_out
for bias in 0.85 1.0; do AGB_BALANCE_BIAS=$bias timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_bias_$bias.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_bias_$bias.json').read().strip().splitlines()[-1])
print('bias $bias: value',round(d['value']),'ms/step',round(d['ms_per_step'],1),d['config']['pipeline'][:40],'early',round(d['early_game']['value']),'e2e',round(d['e2e']['value']))
PY
done
