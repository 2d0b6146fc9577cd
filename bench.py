#!/usr/bin/env python
"""Headline benchmark: self-play positions/sec (= NN evals/sec) of the lockstep engine on BASELINE.json configs[1]
(standard 15x15, ResNet 20 blocks x 128 channels bf16, 4096 concurrent games per GPU).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                      the reference's own CPU search code on the host cores

A "step" is one lockstep iteration over all games of a rank: PUCT select (K6) -> set board + encode (K1+K3) of the leaf
positions -> network forward (K4) -> edge generation / expand / backup (K7) -> final move and subtree reuse where a game
reached its simulation budget. `value` = NN evaluations of all ranks / device time (CUDA events on the engine's stream,
max over ranks). `e2e` = the same metric through the host-buffer C-ABI call agb_evaluate (the NNEvaluator drop-in: host
boards in, host policy/value out, copies inside the timed region).
"""
import argparse
import ctypes
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "self-play positions/sec (NN evals/sec)"
UNIT = "positions/s"
SIZE, RULES = 15, 1  # standard 15x15
BLOCKS, FILTERS = 20, 128
GAMES, BATCH, SIMS = 4096, 8, 400
SOLVER_POSITIONS, SOLVER_TABLE_ENTRIES = 100, 65536  # tss max_positions as in the reference config; per-game table 1 MiB (reference: 64 MiB)
FLOP_PER_POSITION = 2 * 1383.70e6  # BASELINE.md section 3 (algorithmic, ResNet 20x128 @ 15x15, heads p+v)


# --workload: the headline is configs[1]; configs[2] and configs[3] of BASELINE.json can be measured with the same harness (extra lines,
# e.g. profiles/r01_bench_renju15.json), name -> (label, rules, size, simulations, algorithmic FLOP per position from BASELINE.md section 3)
WORKLOADS = {"freestyle15": ("BASELINE.json metric config: freestyle 15x15 self-play, ResNet 20x128 bf16, 4096 concurrent games per GPU", 0, 15, 400, 2 * 1383.70e6),
             "standard15": ("configs[1]: standard 15x15 self-play, ResNet 20x128 bf16, 4096 concurrent games per GPU", 1, 15, 400, 2 * 1383.70e6),
             "renju15": ("configs[2]: renju 15x15 self-play with forbidden-move detection and the solver in the loop, ResNet 20x128 bf16", 2, 15, 400, 2 * 1383.70e6),
             "caro20": ("configs[3]: caro 20x20 self-play, ResNet 20x128 bf16, 800 playouts/move", 3, 20, 800, 2 * 2459.90e6)}
WORKLOAD = WORKLOADS["standard15"][0]
RULE_NAMES = ["FREESTYLE", "STANDARD", "RENJU", "CARO5", "CARO6"]


def select_workload(name):
    global WORKLOAD, RULES, SIZE, SIMS, FLOP_PER_POSITION
    WORKLOAD, RULES, SIZE, SIMS, FLOP_PER_POSITION = WORKLOADS[name]


def workload_config(n_gpus, impl="ours", solver=0):
    return {"workload": WORKLOAD, "rules": RULE_NAMES[RULES],
            "board": f"{SIZE}x{SIZE}", "network": "ResnetPV 20x128", "games_per_gpu": GAMES, "max_batch_size": BATCH, "max_simulations": SIMS, "use_symmetries": True,
            "solver": ("on (AlphaBetaSearch, max_positions 100, 4 Mi-entry table per game)" if impl == "reference" else
                       f"on (K5 alpha-beta, max_positions {solver}, {SOLVER_TABLE_ENTRIES}-entry table per game)" if solver > 0 else "off"),
            "parallelism": f"games sharded over {n_gpus} GPU(s), no data-path collective",
            "cache_note": "each step streams ~12 MB of weights per board from L2 and touches >1 GB of tree/pattern state, larger than L2"}


def random_openings(rng, n):
    boards = np.zeros((n, SIZE * SIZE), np.int8)
    stm = np.ones(n, np.int8)
    c = SIZE // 2
    for g in range(n):
        k = int(rng.integers(0, 9))
        cells = set()
        while len(cells) < k:
            cells.add((c + int(rng.integers(-4, 5))) * SIZE + c + int(rng.integers(-4, 5)))
        for j, cell in enumerate(cells):
            boards[g, cell] = 1 + (j % 2)
        stm[g] = 1 if k % 2 == 0 else 2
    return boards, stm


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self._stop = index, [], set(), threading.Event()
        self.max_mhz = None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for name, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---- reference arm / cpu_baseline: the reference's own search code on host cores ---------------------------------------
def _reference_worker(args):
    """One process = one GeneratorThread of the reference (games_per_thread games, one evaluator, CPU device)."""
    seed, seconds, games_per_thread, batch = args
    import torch
    torch.set_num_threads(1)  # NNEvaluator.cpp:151 forces one thread per evaluator
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refapi
    import nn_oracle
    from alphagomoku_b200 import netblob
    tensors = netblob.random_tensors(SIZE, SIZE, BLOCKS, FILTERS, False)

    def evaluate(features):
        p, v, _ = nn_oracle.forward(tensors, features, SIZE, SIZE, BLOCKS, False)
        return p, v, None

    rng = np.random.default_rng(seed)
    boards, stm = random_openings(rng, games_per_thread)
    games = []
    for g in range(games_per_thread):
        sp = refapi.RefSelfplay(RULES, SIZE, evaluate, max_batch_size=batch, max_simulations=SIMS, init_to="parent", use_solver=True,
                                solver_max_positions=100, fast=True)
        sp.set_position(boards[g], stm[g])
        games.append(sp)
    evals0 = sum(g.evaluations for g in games)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for g, sp in enumerate(games):
            if sp.step() == 2:
                sp.set_position(boards[g], stm[g])
    dt = time.perf_counter() - t0
    return sum(g.evaluations for g in games) - evals0, dt


def run_reference(seconds, threads=None):
    """Reference self-play (unmodified Tree/Search/AlphaBetaSearch from oracle/_ref) on all host cores. The network inside it
    is a stand-in (torch CPU fp32 graph of the same ResNet) because the reference's MinML backend is not in the tree."""
    threads = threads or os.cpu_count()
    with mp.get_context("spawn").Pool(threads) as pool:
        results = pool.map(_reference_worker, [(1000 + i, seconds, 8, 8) for i in range(threads)])
    evals = sum(r[0] for r in results)
    dt = max(r[1] for r in results)
    return evals / dt, threads, evals


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    seconds = min(20.0, max(6.0, 120.0 / (steps + args.warmup)))
    t0 = time.perf_counter()
    values = []
    for _ in range(min(steps, 3)):  # each step = a bounded sample of the workload
        v, cores, evals = run_reference(seconds)
        values.append(v)
    value = float(np.mean(values))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * seconds, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, "reference"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": f"{min(steps, 3)} x {seconds:.0f} s of reference self-play (oracle/_ref: Tree, Search, AlphaBetaSearch with "
                                       f"max_positions 100; 8 games x batch 8 per core; NN = torch-CPU fp32 stand-in for MinML)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


def reference_integer_path(seconds=1.0):
    """rates of the pure reference code on ONE host core (no network involved)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refapi
    ref = refapi.RefOracle(fast=True)
    rng = np.random.default_rng(0)
    boards, _ = random_openings(rng, 256)
    rates = np.zeros(3)
    ref.lib.agref_bench_integer_path.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
    ref.lib.agref_bench_integer_path(RULES, SIZE, SIZE, refapi._p(boards), 256, seconds, 100, refapi._p(rates))
    return {"set_board_and_encode_per_s": float(rates[0]), "add_undo_pairs_per_s": float(rates[1]), "solve_100_positions_per_s": float(rates[2]),
            "cores": 1, "note": "oracle/_ref (reference sources, -O3 -DNDEBUG), opening positions of this bench"}


# ---- our arm -------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--games", type=int, default=GAMES)
    ap.add_argument("--batch", type=int, default=BATCH, help="SearchConfig::max_batch_size (leaves per game and step)")
    ap.add_argument("--groups", type=int, default=0, help="AgbConfig::pipeline_groups (0 = engine default: 2 with the alpha-beta solver on)")
    ap.add_argument("--solver-sms", type=int, default=0, help="AgbConfig::solver_sms (0 = engine default: 28 of 148 SMs, -1 = no partition)")
    ap.add_argument("--solver", type=int, default=SOLVER_POSITIONS, help="TSSConfig::max_positions of the device solver (0 = off)")
    ap.add_argument("--workload", default="standard15", choices=sorted(WORKLOADS), help="standard15 = BASELINE.json configs[1] (the headline)")
    ap.add_argument("--shard", type=int, default=0, help="play the games rank SHARD of a larger job would play (openings, game ids); for variance checks")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(3, args.warmup)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    games = args.games
    nodes_per_game = 1536 * SIMS // 400  # the tree of one move plus the subtree kept from the previous one
    eng = agb.Engine(agb.GameConfig(agb.GameRules(RULES), SIZE, SIZE), max_boards=games * args.batch, device=local_rank, blocks=BLOCKS, filters=FILTERS,
                     q_head=False, games=games, max_batch_size=args.batch, max_simulations=SIMS, init_to="parent", max_nodes_per_game=nodes_per_game,
                     max_edges_per_game=nodes_per_game * 200, seed=1234, first_game_id=(rank + args.shard) * games, solver_max_positions=args.solver,
                     solver_table_entries=SOLVER_TABLE_ENTRIES, pipeline_groups=args.groups, solver_sms=args.solver_sms, use_symmetries=True)
    # C1: rank 0 owns the weights and broadcasts them over NCCL (NetworkLoader::get per thread in the reference)
    blob = netblob.pack(netblob.random_tensors(SIZE, SIZE, BLOCKS, FILTERS, False), SIZE, SIZE, BLOCKS, FILTERS, False) if rank == 0 else None
    eng.load_weights(sharding.broadcast_weights(blob))
    rng = np.random.default_rng(99 + rank + args.shard)
    boards, stm = random_openings(rng, games)
    eng.selfplay_reset(boards, stm)

    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.step(warmup)
    barrier()
    st0 = eng.stats()
    with ClockSampler(local_rank) as clocks:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record(stream)
        eng.step(args.steps)  # exactly K lockstep steps; no host sync inside
        end.record(stream)
        end.synchronize()
        barrier()
    ms = start.elapsed_time(end)
    st1 = eng.stats()
    evals = st1["nb_network_evaluations"] - st0["nb_network_evaluations"]
    launches = st1["nb_kernel_launches"] - st0["nb_kernel_launches"]
    nn_ns = st1["nn_kernel_ns"] - st0["nn_kernel_ns"]
    nn_launches = st1["nn_kernel_launches"] - st0["nn_kernel_launches"]
    nn_positions = st1["nn_positions"] - st0["nn_positions"]

    # e2e: NNEvaluator drop-in through host buffers (pinned), copies inside the timed region
    n_e2e = games * args.batch
    e_boards = torch.from_numpy(np.repeat(boards, args.batch, axis=0)).pin_memory()
    e_stm = torch.from_numpy(np.repeat(stm, args.batch)).pin_memory()
    e_policy = torch.empty((n_e2e, SIZE * SIZE), dtype=torch.float32).pin_memory()
    e_value = torch.empty((n_e2e, 3), dtype=torch.float32).pin_memory()
    lib = eng._lib

    def e2e_step():
        rc = lib.agb_evaluate(eng._h, ctypes.c_void_p(e_boards.data_ptr()), ctypes.c_void_p(e_stm.data_ptr()), None, n_e2e,
                              ctypes.c_void_p(e_policy.data_ptr()), ctypes.c_void_p(e_value.data_ptr()), None)
        assert rc == 0, lib.agb_last_error(eng._h)

    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # device rates of the integer path alone (no network): K1+K3 on the e2e boards, K5 from the solver kernel's event times
    d_boards, d_stm = torch.from_numpy(np.repeat(boards, args.batch, axis=0)).cuda(), torch.from_numpy(np.repeat(stm, args.batch)).cuda()
    d_feat = torch.empty((n_e2e, SIZE * SIZE), dtype=torch.int32, device="cuda")

    def k1():
        assert lib.agb_set_boards_dev(eng._h, ctypes.c_void_p(d_boards.data_ptr()), ctypes.c_void_p(d_stm.data_ptr()), n_e2e, ctypes.c_void_p(d_feat.data_ptr())) == 0

    k1()
    k1_start, k1_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k1_start.record(stream)
    for _ in range(5):
        k1()
    k1_end.record(stream)
    k1_end.synchronize()
    # K4 (+ value head) alone on all SMs, on the same boards: the kernel's own roofline point, measured live next to the in-step one
    d_policy = torch.empty((n_e2e, SIZE * SIZE), dtype=torch.float32, device="cuda")
    d_value = torch.empty((n_e2e, 3), dtype=torch.float32, device="cuda")

    def k4():
        assert lib.agb_forward_dev(eng._h, ctypes.c_void_p(d_feat.data_ptr()), n_e2e, ctypes.c_void_p(d_policy.data_ptr()), ctypes.c_void_p(d_value.data_ptr()), None) == 0

    k4()
    k4_start, k4_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k4_start.record(stream)
    for _ in range(3):
        k4()
    k4_end.record(stream)
    k4_end.synchronize()
    k4_alone_ms = k4_start.elapsed_time(k4_end) / 3
    solver_ns = st1["solver_kernel_ns"] - st0["solver_kernel_ns"]
    device_integer_path = {"set_board_and_encode_per_s": 5 * n_e2e / (k1_start.elapsed_time(k1_end) * 1e-3),
                           "solve_100_positions_per_s": ((st1["nb_node_count"] - st0["nb_node_count"]) / (solver_ns * 1e-9)) if solver_ns else None,
                           "note": "one GPU; leaf positions solved per second of K5 kernel time, K1+K3 on 32 k boards per launch"}
    # C2 (outside the timed regions): finished-game records of every rank, all-gathered like the reference's shared GameDataBuffer
    records, n_finished = eng.pop_finished()
    gathered = sharding.gather_records(records)
    c2_bytes = sum(len(g) for g in gathered)
    # whole-job totals: sum of the units every rank processed, max of the device times
    sums, maxes = sharding.reduce_counters([float(evals), ms, float(n_e2e * e2e_steps), e2e_s, float(launches)])
    if rank == 0:
        total_evals, max_ms = sums[0], maxes[1]
        value = total_evals / (max_ms / 1e3)
        e2e_value = sums[2] / maxes[3]
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["bf16_tflops_sustained"], "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
        else:
            peak, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained)"
        total_sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        groups_eff = args.groups if args.groups > 0 else (2 if args.solver > 1 and games >= 1024 else 1)
        partition_sms = int(st0["solver_sms"])  # as the timed call ran (the automatic mode re-balances after each agb_step call)
        achieved = (FLOP_PER_POSITION * nn_positions / max(nn_launches, 1)) / (nn_ns / max(nn_launches, 1) * 1e-9) / 1e12 if nn_ns else None
        traffic = None
        prof = os.path.join(ROOT, "profiles", "r01_k4_ncu_summary.json")
        if os.path.exists(prof):
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": max_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": dict(workload_config(world, solver=args.solver),
                               pipeline=("one group, K5 then K4 on all SMs" if groups_eff == 1 or args.solver <= 1 else
                                         f"{groups_eff} groups of games; K5 on {partition_sms} SMs side by side with K4 on the others"
                                         if partition_sms > 0 else f"{groups_eff} groups of games, no SM partition")),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_e2e * (SIZE * SIZE + 1)), "d2h_bytes_per_step": int(n_e2e * (SIZE * SIZE + 3) * 4),
                        "api": "agb_evaluate (NNEvaluator::evaluateGraph drop-in): pinned host boards -> K1+K3+K4 -> host policy/value"},
                "gpu_launches": int(sums[4]), "overflow_flags": int(st1["overflow_flags"]), "solver_sms_during_timed_steps": int(st0["solver_sms"]),
                "roofline": {"bound": "tensor", "kernel": "resnet_board_kernel (+ value head)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                             "note": ("K4 timed alone on all SMs" if partition_sms == 0 else
                                      f"K4 runs on {total_sms - partition_sms} of {total_sms} SMs, side by side with K5 on the other {partition_sms}; frac is against the "
                                      f"whole GPU's peak ({achieved / (peak * (total_sms - partition_sms) / total_sms):.3f} of its own SMs' share); "
                                      f"alone on all SMs it reaches the fraction under 'alone' (also --groups 1, profiles/r01_bench_n1_groups1.json)"),
                             "positions_per_launch": nn_positions / max(nn_launches, 1), "ms_per_launch": nn_ns / max(nn_launches, 1) * 1e-6,
                             "alone": {"what": "the same kernel (+ value head) alone on all SMs, agb_forward_dev on device-resident features, CUDA events on its stream",
                                       "positions_per_launch": n_e2e, "ms_per_launch": k4_alone_ms, "achieved": FLOP_PER_POSITION * n_e2e / (k4_alone_ms * 1e-3) / 1e12,
                                       "frac": FLOP_PER_POSITION * n_e2e / (k4_alone_ms * 1e-3) / 1e12 / peak},
                             "share_of_step": (nn_ns * 1e-6) / ms, "solver_share_of_step": (st1["solver_kernel_ns"] - st0["solver_kernel_ns"]) * 1e-6 / ms,
                             "leaf_positions_per_step": (st1["nb_node_count"] - st0["nb_node_count"]) / args.steps},
                "sharding": {"c1_weight_bytes_broadcast": int(blob.nbytes), "c2_record_bytes_gathered": int(c2_bytes), "ranks": world},
                "clocks": clocks.summary()}
        if world == 1 and not args.no_cpu_baseline:
            v, cores, _ = run_reference(args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                                    "sample": f"{args.cpu_seconds:.0f} s of reference self-play on the host cores (oracle/_ref Tree/Search/AlphaBetaSearch, "
                                              f"8 games x batch 8 per core, solver on; NN = torch-CPU fp32 stand-in for MinML)",
                                    "integer_path_one_core": reference_integer_path(), "integer_path_device": device_integer_path}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
