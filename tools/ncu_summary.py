"""Per-kernel summary of an ncu report (ncu --set full): the counters DESIGN.md and bench.py quote, as JSON.
  python tools/ncu_summary.py gpurun_out/r02_step_kernels.ncu-rep > profiles/r02_step_kernels_ncu.json"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_ns", "launch__grid_size": "grid", "launch__block_size": "block", "launch__registers_per_thread": "registers_per_thread",
    "smsp__inst_executed.sum": "warp_instructions", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed.avg.per_cycle_active": "ipc_per_sm_active", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "dram__bytes_read.sum.per_second": "dram_read_per_s",
    "dram__bytes_write.sum.per_second": "dram_write_per_s", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__t_bytes.sum": "l2_bytes", "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read_bytes", "sm__icc_request_hit_rate.pct": "instruction_cache_hit_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_instructions", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed": "memory_throughput_pct",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving",
}

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
header, units = rows[0], rows[1]
result = []
for r in rows[2:]:
    rec = {"kernel": r[header.index("Kernel Name")].split("(")[0].replace("void ", "").replace("unnamed>::", "")}
    for name, key in WANT.items():
        if name in header:
            i = header.index(name)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            unit = units[i]
            scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12, "us": 1e3, "ms": 1e6,
                     "s": 1e9}.get(unit, 1.0) if key != "duration_ns" or unit != "ns" else 1.0
            rec[key] = v * scale
    if "dram_read" in rec and "dram_write" in rec:
        rec["dram_bytes"] = rec["dram_read"] + rec["dram_write"]
        if rec.get("duration_ns"):
            rec["dram_gb_per_s"] = rec["dram_bytes"] / rec["duration_ns"]
    result.append(rec)
print(json.dumps(result, indent=1))
