#!/bin/bash
# final check of a build on a B200 box: the whole GPU suite, then the default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 1200 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -c 300 gpurun_out/r02_bench_n1_final.json
for w in caro20; do timeout 900 python bench.py --workload $w --start openings --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_$w.json 2> gpurun_out/r02_bench_n1_$w.err; done
