#!/bin/bash
mkdir -p gpurun_out
for cfg in "6 64" "6 56" "6 72" "4 64" "8 64"; do set -- $cfg; timeout 300 python tools/steady_bench.py bench_data/steady_freestyle15.npz 60 40 $2 freestyle15 $1 2>&1 | tail -1; done | tee gpurun_out/r02_steady_green5.txt
AGB_VERBOSE=1 timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err; tail -c 400 gpurun_out/r02_bench_n1_c.json; tail -3 gpurun_out/r02_bench_n1_c.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_ref_a.json 2> gpurun_out/r02_bench_ref_a.err; tail -c 600 gpurun_out/r02_bench_ref_a.json
