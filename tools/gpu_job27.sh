#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_d.json 2>gpurun_out/r02_bench_d.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_d.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],1),d['config']['pipeline'][:60],'early',round(d['early_game']['value']),'e2e',round(d['e2e']['value']), 'k4',d['sharding']['per_rank_k4_ms_per_step'],'k5',d['sharding']['per_rank_k5_ms_per_step'], 'sms', d['solver_sms_during_timed_steps'])
PY
AGB_STEP_TRACE=gpurun_out/r02_step_trace_3groups.txt timeout 600 python tools/steady_bench.py bench_data/steady_freestyle15.npz 100 40 0 freestyle15 0 2>&1 | tail -1
tail -32 gpurun_out/r02_step_trace_3groups.txt
