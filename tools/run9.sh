mkdir -p gpurun_out
for cfg in "32 2" "16 4" "8 8"; do set -- $cfg; echo "stage_kb=$1 stages=$2"; AGB_NET_STAGE_KB=$1 AGB_NET_STAGES=$2 AGB_NET_TRACE=gpurun_out/trace_$1x$2.txt python tools/bench_forward.py 20 128 0 4096 2 2>&1 | tail -1; done
