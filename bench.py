#!/usr/bin/env python
"""Headline benchmark: self-play positions/sec (= NN evals/sec) of the lockstep engine on the configuration BASELINE.json's metric is
quoted on: freestyle 15x15, ResNet 20 blocks x 128 channels bf16, 4096 concurrent games per GPU, 400 simulations, solver on.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                      the reference's own GeneratorThread::run loop on the host cores

A "step" is one lockstep iteration over all games of a rank: PUCT select (K6) -> set board + encode (K1+K3) of the leaf positions ->
solver (K5) -> network forward (K4) on the unproven leaves -> edge generation / expand / backup (K7) -> final move, record and subtree
reuse where a game reached its simulation budget (K8).

The timed steps are STEADY STATE: every game starts from a committed mid-game snapshot (bench_data/steady_<workload>.npz, written by
tools/make_snapshot.py from a 700-step soak: a stationary mixture of game phases), plays `--settle` untimed steps to refill the search
trees and settle the SM split, then W warm-up and K timed steps. The early-game rate (random openings, the round-1 measurement) is
printed next to it under `early_game`. Inside the timed region finished games are popped every `--pop-every` steps and their records
all-gathered over the ranks together with the finished-game counter (C2 and the hasEnoughGames reduction of GeneratorManager.cpp:160-181).

`value` = NN evaluations of all ranks / device time (CUDA events on the engine's stream, max over ranks).
`e2e`   = the same metric as host wall-clock around agb_step(K) + agb_pop_finished through the C ABI (records land in pinned host memory).
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
SIZE, RULES = 15, 0  # freestyle 15x15
BLOCKS, FILTERS = 20, 128
GAMES, BATCH, SIMS = 4096, 8, 400
SOLVER_POSITIONS, SOLVER_TABLE_ENTRIES = 100, 65536  # tss max_positions as in the reference config; per-game table 1 MiB (reference: 64 MiB)
FLOP_PER_POSITION = 2 * 1383.70e6  # BASELINE.md section 3 (algorithmic, ResNet 20x128 @ 15x15, heads p+v)

# --workload: name -> (label, rules, size, simulations, algorithmic FLOP per position from BASELINE.md section 3)
WORKLOADS = {"freestyle15": ("BASELINE.json metric config: freestyle 15x15 self-play, ResNet 20x128 bf16, 4096 concurrent games per GPU", 0, 15, 400, 2 * 1383.70e6),
             "standard15": ("configs[1]: standard 15x15 self-play, ResNet 20x128 bf16, 4096 concurrent games per GPU", 1, 15, 400, 2 * 1383.70e6),
             "renju15": ("configs[2]: renju 15x15 self-play with forbidden-move detection and the solver in the loop, ResNet 20x128 bf16", 2, 15, 400, 2 * 1383.70e6),
             "caro20": ("configs[3]: caro 20x20 self-play, ResNet 20x128 bf16, 800 playouts/move", 3, 20, 800, 2 * 2459.90e6)}
WORKLOAD_NAME = "freestyle15"
WORKLOAD = WORKLOADS[WORKLOAD_NAME][0]
RULE_NAMES = ["FREESTYLE", "STANDARD", "RENJU", "CARO5", "CARO6"]


def select_workload(name):
    global WORKLOAD_NAME, WORKLOAD, RULES, SIZE, SIMS, FLOP_PER_POSITION
    WORKLOAD_NAME = name
    WORKLOAD, RULES, SIZE, SIMS, FLOP_PER_POSITION = WORKLOADS[name]


def snapshot_path():
    return os.path.join(ROOT, "bench_data", f"steady_{WORKLOAD_NAME}.npz")


def workload_config(n_gpus, impl="ours", solver=0, start="snapshot"):
    return {"workload": WORKLOAD, "rules": RULE_NAMES[RULES],
            "board": f"{SIZE}x{SIZE}", "network": "ResnetPV 20x128", "games_per_gpu": GAMES, "max_batch_size": BATCH, "max_simulations": SIMS, "use_symmetries": True,
            "solver": ("on (AlphaBetaSearch, max_positions 100, 4 Mi-entry table per game)" if impl == "reference" else
                       f"on (K5 alpha-beta, max_positions {solver}, {SOLVER_TABLE_ENTRIES}-entry table per game)" if solver > 0 else "off"),
            "start": ("steady state: games start from the committed mid-game snapshot bench_data/steady_%s.npz" % WORKLOAD_NAME) if start == "snapshot"
            else "early game: random openings of 0..8 stones",
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


def unpack_boards(packed, cells):
    """bench_data snapshots hold 2 bits per cell (tools/make_snapshot.py:pack_boards)."""
    n = packed.shape[0]
    out = np.zeros((n, packed.shape[1], 4), np.int8)
    for k in range(4):
        out[:, :, k] = (packed >> (2 * k)) & 3
    return out.reshape(n, -1)[:, :cells].copy()


def snapshot_positions(n, shard=0):
    """`n` steady-state positions for shard `shard`: the snapshot's games, rotated by shard * n and seen through board symmetry
    (shard mod 8), so that every rank plays different games of the same distribution. None when no snapshot is committed."""
    path = snapshot_path()
    if not os.path.exists(path):
        return None
    snap = np.load(path)
    boards, stm = unpack_boards(snap["boards"], SIZE * SIZE), snap["sign_to_move"].astype(np.int8)
    idx = (np.arange(n) + shard * n) % boards.shape[0]
    boards, stm = boards[idx], stm[idx]
    b = boards.reshape(n, SIZE, SIZE)
    sym = ((shard * n) // boards.shape[0]) % 8  # every full pass over the snapshot sees it through the next symmetry
    if sym & 1:
        b = b[:, ::-1, :]
    if sym & 2:
        b = b[:, :, ::-1]
    if sym & 4:
        b = b.transpose(0, 2, 1)
    return np.ascontiguousarray(b).reshape(n, SIZE * SIZE), np.ascontiguousarray(stm)


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


# ---- reference arm / cpu_baseline: the reference's own self-play loop on host cores ------------------------------------------
def _reference_worker(args):
    """One process = one GeneratorThread of the reference: the unmodified GeneratorThread::run loop (GeneratorManager.cpp:120-141,
    through oracle/ref_shim_manager.cpp) with games_per_thread GameGenerators and one NNEvaluator of batch 64, on one host core.
    Returns the evaluated positions of every `interval`-second slice of the run."""
    index, workload, interval, n_intervals, games_per_thread, batch, evaluator_batch, start = args
    select_workload(workload)
    import torch
    torch.set_num_threads(1)  # NNEvaluator.cpp:151 forces one thread per evaluator
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refapi
    import nn_oracle
    from alphagomoku_b200 import netblob
    tensors = netblob.random_tensors(SIZE, SIZE, BLOCKS, FILTERS, False)
    stamps = []

    def evaluate(features):
        p, v, _ = nn_oracle.forward(tensors, features, SIZE, SIZE, BLOCKS, False)
        stamps.append((time.perf_counter(), features.shape[0]))
        return p, v, None

    start_boards = None
    if start == "snapshot":
        pos = snapshot_positions(games_per_thread, shard=index)
        start_boards = None if pos is None else pos[0]
    elif start == "openings":
        start_boards = random_openings(np.random.default_rng(1000 + index), games_per_thread)[0]
    t0 = time.perf_counter()
    stats = refapi.run_generator_threads(RULES, SIZE, evaluate, interval * n_intervals, threads=1, games_per_thread=games_per_thread, max_batch_size=batch,
                                         evaluator_batch=evaluator_batch, max_simulations=SIMS, solver_max_positions=SOLVER_POSITIONS, use_opening=True,
                                         use_symmetries=True, init_to="parent", start_boards=start_boards, fast=True)
    per_interval = np.zeros(n_intervals)
    for t, n in stamps:
        k = int((t - t0) / interval)
        if 0 <= k < n_intervals:
            per_interval[k] += n
    return per_interval, stats["nb_network_evaluations"], stats["seconds"], stats["games_finished"]


def run_reference(interval, n_intervals, threads=None, start="snapshot"):
    """Reference self-play on all host cores: one process per core (= one GeneratorThread per CPU DeviceConfig entry), 8 games x batch 8
    per thread, evaluator batch 64. The network inside it is a stand-in (torch CPU fp32 graph of the same ResNet, one thread per
    evaluator) because the reference's MinML backend is not in its tree. Returns positions/s per interval (summed over the cores)."""
    threads = threads or os.cpu_count()
    with mp.get_context("spawn").Pool(threads) as pool:
        results = pool.map(_reference_worker, [(i, WORKLOAD_NAME, interval, n_intervals, 8, 8, 64, start) for i in range(threads)])
    per_interval = sum(r[0] for r in results) / interval
    search_evals = sum(r[1] for r in results)
    seconds = max(r[2] for r in results)
    return per_interval, threads, search_evals / seconds, sum(r[3] for r in results)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    interval = min(8.0, max(2.0, 150.0 / (steps + warmup)))  # each step = `interval` seconds of the reference's loop; the whole run fits in minutes
    t0 = time.perf_counter()
    per_interval, cores, search_rate, finished = run_reference(interval, steps + warmup, start=args.start)
    timed = per_interval[warmup:]
    value = float(np.mean(timed))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1000.0 * interval, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.gpus, "reference", start=args.start), games_per_thread=8, threads=cores, evaluator_batch=64),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": f"{steps} timed + {warmup} warm-up slices of {interval:.1f} s of the reference's own GeneratorThread::run loop (oracle/_ref: "
                                       f"GeneratorManager, GameGenerator, Search, Tree, AlphaBetaSearch with max_positions 100, NNEvaluator batch 64; one thread per "
                                       f"host core, 8 games x batch 8 each; NN = torch-CPU fp32 stand-in for MinML)"},
            "per_step_values": [float(v) for v in timed],
            "search_stats_rate": float(search_rate), "games_finished": int(finished),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


def reference_integer_path(seconds=1.0):
    """rates of the pure reference code on ONE host core (no network involved), on the bench's positions."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refapi
    ref = refapi.RefOracle(fast=True)
    pos = snapshot_positions(256)
    boards = pos[0] if pos is not None else random_openings(np.random.default_rng(0), 256)[0]
    rates = np.zeros(3)
    ref.lib.agref_bench_integer_path.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
    ref.lib.agref_bench_integer_path(RULES, SIZE, SIZE, refapi._p(boards), 256, seconds, 100, refapi._p(rates))
    return {"set_board_and_encode_per_s": float(rates[0]), "add_undo_pairs_per_s": float(rates[1]), "solve_100_positions_per_s": float(rates[2]),
            "cores": 1, "note": "oracle/_ref (reference sources, -O3 -DNDEBUG) on the positions this bench starts from"}


# ---- our arm -------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--games", type=int, default=GAMES)
    ap.add_argument("--batch", type=int, default=BATCH, help="SearchConfig::max_batch_size (leaves per game and step)")
    ap.add_argument("--groups", type=int, default=0, help="AgbConfig::pipeline_groups (0 = engine default)")
    ap.add_argument("--solver-sms", type=int, default=0, help="AgbConfig::solver_sms (0 = engine default / automatic, -1 = no partition)")
    ap.add_argument("--solver", type=int, default=SOLVER_POSITIONS, help="TSSConfig::max_positions of the device solver (0 = off)")
    ap.add_argument("--workload", default="freestyle15", choices=sorted(WORKLOADS), help="freestyle15 = the configuration BASELINE.json's metric names (the headline)")
    ap.add_argument("--start", default="snapshot", choices=["snapshot", "openings"], help="snapshot = steady state (headline); openings = early game only")
    ap.add_argument("--settle", type=int, default=100, help="untimed steps after the reset to the snapshot (trees refill in 50 steps at 8 of 400 simulations per step; the automatic SM split needs about ten calls of 10 steps to settle)")
    ap.add_argument("--pop-every", type=int, default=10, help="steps between agb_pop_finished + record all-gather + finished-counter all-reduce in the timed region")
    ap.add_argument("--shard", type=int, default=0, help="play the games rank SHARD of a larger job would play (openings, game ids); for variance checks")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-early-game", action="store_true")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.start == "snapshot" and not os.path.exists(snapshot_path()):
        args.start = "openings"  # no snapshot committed for this workload: early game only, and the line says so
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
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local_rank))
    lib = eng._lib

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # pinned host buffer the finished-game records are popped into (agb_pop_finished is the device -> host read of the step's result)
    rec_cap = 64 << 20
    rec_host = torch.empty(rec_cap, dtype=torch.uint8).pin_memory()
    popped = {"bytes": 0, "games": 0, "gathered_bytes": 0, "pops": 0, "finished_total": 0.0}

    def pop_and_gather():
        used, n_games = ctypes.c_size_t(0), ctypes.c_int(0)
        rc = lib.agb_pop_finished(eng._h, ctypes.c_void_p(rec_host.data_ptr()), ctypes.c_size_t(rec_cap), ctypes.byref(used), ctypes.byref(n_games))
        assert rc == 0, lib.agb_last_error(eng._h)
        popped["bytes"] += used.value
        popped["games"] += n_games.value
        popped["pops"] += 1
        if world > 1:  # C2 + hasEnoughGames: every rank sees every record and the global count of finished games
            gathered = sharding.gather_records(rec_host[:used.value].numpy().tobytes())
            popped["gathered_bytes"] += sum(len(g) for g in gathered)
            sums, _ = sharding.reduce_counters([float(n_games.value)])
            popped["finished_total"] += float(sums[0])
        else:
            popped["gathered_bytes"] += used.value
            popped["finished_total"] += n_games.value

    def timed_steps(n_steps):
        """Exactly n_steps lockstep steps with the production loop's pops in between; device time by CUDA events on the engine's stream."""
        st0 = eng.stats()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        start.record(stream)
        done = 0
        while done < n_steps:
            chunk = min(args.pop_every if args.pop_every > 0 else n_steps, n_steps - done)
            eng.step(chunk)
            done += chunk
            if args.pop_every > 0:
                pop_and_gather()
        end.record(stream)
        end.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        st1 = eng.stats()
        return start.elapsed_time(end), wall, {k: st1[k] - st0[k] for k in st1}, st0, st1

    # ---- early game (the round-1 measurement): random openings, a few steps ----
    early = None
    if not args.no_early_game and args.start == "snapshot":
        boards, stm = random_openings(np.random.default_rng(99 + rank + args.shard), games)
        eng.selfplay_reset(boards, stm)
        eng.step(warmup)
        e_steps = min(args.steps, 10)
        ms, _, d, _, _ = timed_steps(e_steps)
        eng.pop_finished()
        sums, maxes = sharding.reduce_counters([float(d["nb_network_evaluations"]), ms])
        early = {"value": sums[0] / (maxes[1] / 1e3), "unit": UNIT, "ms_per_step": maxes[1] / e_steps, "steps": e_steps,
                 "what": "the same engine from random openings of 0..8 stones (steps %d..%d of a game): few threats, the solver is cheap" % (warmup + 1, warmup + e_steps)}

    # ---- steady state (the headline) ----
    if args.start == "snapshot":
        boards, stm = snapshot_positions(games, shard=rank + args.shard)
    else:
        boards, stm = random_openings(np.random.default_rng(99 + rank + args.shard), games)
    eng.selfplay_reset(boards, stm)
    settle = args.settle if args.start == "snapshot" else 0
    for _ in range(settle // 20):
        eng.step(20)
    if settle % 20:
        eng.step(settle % 20)
    eng.step(warmup)
    eng.pop_finished()
    popped.update(bytes=0, games=0, gathered_bytes=0, pops=0, finished_total=0.0)
    with ClockSampler(local_rank) as clocks:
        ms, _, d, st0, st1 = timed_steps(args.steps)
    evals = d["nb_network_evaluations"]
    launches = d["nb_kernel_launches"]
    nn_ns, nn_launches, nn_positions, solver_ns = d["nn_kernel_ns"], d["nn_kernel_launches"], d["nn_positions"], d["solver_kernel_ns"]
    leaves = d["nb_node_count"]
    pop_stats = dict(popped)

    # ---- e2e: host wall-clock of agb_step(K) + agb_pop_finished through the C ABI (the call sequence a GeneratorManager-style host makes) ----
    popped.update(bytes=0, games=0, gathered_bytes=0, pops=0, finished_total=0.0)
    e2e_steps = args.steps
    _, e2e_wall, e2e_d, _, _ = timed_steps(e2e_steps)
    e2e_evals = e2e_d["nb_network_evaluations"]
    e2e_d2h = popped["bytes"] / e2e_steps + 24.0 * max(1, popped["pops"]) / e2e_steps  # records + the counter / status words each call reads back

    # ---- evaluator drop-in (agb_evaluate: NNEvaluator::evaluateGraph through host buffers), kept next to the e2e number ----
    n_eval = games * args.batch
    e_boards = torch.from_numpy(np.repeat(boards, args.batch, axis=0)).pin_memory()
    e_stm = torch.from_numpy(np.repeat(stm, args.batch)).pin_memory()
    e_policy = torch.empty((n_eval, SIZE * SIZE), dtype=torch.float32).pin_memory()
    e_value = torch.empty((n_eval, 3), dtype=torch.float32).pin_memory()

    def evaluator_step():
        rc = lib.agb_evaluate(eng._h, ctypes.c_void_p(e_boards.data_ptr()), ctypes.c_void_p(e_stm.data_ptr()), None, n_eval,
                              ctypes.c_void_p(e_policy.data_ptr()), ctypes.c_void_p(e_value.data_ptr()), None)
        assert rc == 0, lib.agb_last_error(eng._h)

    for _ in range(2):
        evaluator_step()
    barrier()
    ev_steps = 5
    t0 = time.perf_counter()
    for _ in range(ev_steps):
        evaluator_step()
    torch.cuda.synchronize()
    ev_s = time.perf_counter() - t0

    # device rates of the integer path alone (no network): K1+K3 on the same boards, K5 from the solver kernel's event times
    d_boards, d_stm = torch.from_numpy(np.repeat(boards, args.batch, axis=0)).cuda(), torch.from_numpy(np.repeat(stm, args.batch)).cuda()
    d_feat = torch.empty((n_eval, SIZE * SIZE), dtype=torch.int32, device="cuda")

    def k1():
        assert lib.agb_set_boards_dev(eng._h, ctypes.c_void_p(d_boards.data_ptr()), ctypes.c_void_p(d_stm.data_ptr()), n_eval, ctypes.c_void_p(d_feat.data_ptr())) == 0

    k1()
    k1_start, k1_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k1_start.record(stream)
    for _ in range(5):
        k1()
    k1_end.record(stream)
    k1_end.synchronize()
    # K4 (+ value head) alone on all SMs, on the same boards: the kernel's own roofline point, measured live next to the in-step one
    d_policy = torch.empty((n_eval, SIZE * SIZE), dtype=torch.float32, device="cuda")
    d_value = torch.empty((n_eval, 3), dtype=torch.float32, device="cuda")

    def k4():
        assert lib.agb_forward_dev(eng._h, ctypes.c_void_p(d_feat.data_ptr()), n_eval, ctypes.c_void_p(d_policy.data_ptr()), ctypes.c_void_p(d_value.data_ptr()), None) == 0

    k4()
    k4_start, k4_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k4_start.record(stream)
    for _ in range(3):
        k4()
    k4_end.record(stream)
    k4_end.synchronize()
    k4_alone_ms = k4_start.elapsed_time(k4_end) / 3
    device_integer_path = {"set_board_and_encode_per_s": 5 * n_eval / (k1_start.elapsed_time(k1_end) * 1e-3),
                           "solve_100_positions_per_s": (leaves / (solver_ns * 1e-9)) if solver_ns else None,
                           "note": "one GPU; leaf positions solved per second of K5 kernel time in the timed steady-state steps, K1+K3 on 32 k boards per launch"}
    # whole-job totals: sum of the units every rank processed, max of the device times; per-rank kernel times for the spread
    sums, maxes = sharding.reduce_counters([float(evals), ms, float(e2e_evals), e2e_wall, float(launches), float(n_eval * ev_steps), ev_s,
                                            float(pop_stats["bytes"]), float(pop_stats["games"])])
    _, neg_mins = sharding.reduce_counters([-float(solver_ns), -float(nn_ns), -ms])
    _, k_maxes = sharding.reduce_counters([float(solver_ns), float(nn_ns)])
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
        partition_sms = int(st0["solver_sms"])  # as the timed call ran (the automatic mode re-balances after each agb_step call)
        positions_per_launch = nn_positions / max(nn_launches, 1)
        achieved = (FLOP_PER_POSITION * positions_per_launch) / (nn_ns / max(nn_launches, 1) * 1e-9) / 1e12 if nn_ns else None
        traffic, traffic_src = None, None
        prof = os.path.join(ROOT, "profiles", "r02_k4_ncu_instep.json")
        if os.path.exists(prof):
            pj = json.load(open(prof))
            traffic = pj.get("dram_bytes_per_launch")
            traffic_src = f"profiles/r02_k4_ncu_instep.json: ncu --set full of one in-step launch of {pj.get('positions_per_launch')} positions"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": max_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": dict(workload_config(world, solver=args.solver, start=args.start), settle_steps=settle, pop_every=args.pop_every,
                               pipeline=(f"{int(st1.get('pipeline_groups', 0)) or 'engine default'} groups of games; K5 on {partition_sms} SMs side by side with K4 on the others"
                                         if partition_sms > 0 else "one group, K5 then K4 on all SMs"),
                               differences_from_reference_arm=["per-game solver table: 65 536 entries here, 4 Mi in the reference (results identical when given the reference's size; "
                                                               "soundness and move agreement of the small table: tests/test_selfplay_gpu.py)",
                                                               "concurrency: 4096 games x 8 leaves per GPU here, 8 games x 8 leaves per host core there",
                                                               "the reference arm's network is a torch-CPU fp32 stand-in for MinML (not in the reference tree)",
                                                               "evaluation symmetries and root noise come from counter-based streams keyed by the game id, not a thread-local mt19937"]),
                "early_game": early,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(e2e_d2h),
                        "api": "host wall-clock of agb_step(%d) + agb_pop_finished into pinned host memory, %d steps; the games are device-resident, so a step "
                               "has no host inputs: what crosses the bus is the finished-game records and the counter / status words"
                               % (min(args.pop_every, args.steps) if args.pop_every > 0 else args.steps, e2e_steps)},
                "evaluator_e2e": {"value": sums[5] / maxes[6], "unit": UNIT, "h2d_bytes_per_step": int(n_eval * (SIZE * SIZE + 1)), "d2h_bytes_per_step": int(n_eval * (SIZE * SIZE + 3) * 4),
                                  "api": "agb_evaluate (NNEvaluator::evaluateGraph drop-in): pinned host boards -> K1+K3+K4 -> host policy/value; no tree, no solver"},
                "gpu_launches": int(sums[4]), "overflow_flags": int(st1["overflow_flags"]), "solver_sms_during_timed_steps": partition_sms,
                "records": {"popped_games": int(sums[8]), "popped_bytes": int(sums[7]), "pops_in_timed_region": pop_stats["pops"],
                            "c2_gathered_bytes_per_rank": int(pop_stats["gathered_bytes"]), "finished_games_allreduced": pop_stats["finished_total"]},
                "roofline": {"bound": "tensor", "kernel": "resnet_board_kernel (+ value head)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                             "note": ("K4 timed alone on all SMs" if partition_sms == 0 else
                                      f"K4 runs on {total_sms - partition_sms} of {total_sms} SMs, side by side with K5 on the other {partition_sms}; frac is against the "
                                      f"whole GPU's peak ({achieved / (peak * (total_sms - partition_sms) / total_sms):.3f} of its own SMs' share); "
                                      f"alone on all SMs it reaches the fraction under 'alone'"),
                             "positions_per_launch": positions_per_launch, "ms_per_launch": nn_ns / max(nn_launches, 1) * 1e-6,
                             "alone": {"what": "the same kernel (+ value head) alone on all SMs, agb_forward_dev on device-resident features, CUDA events on its stream",
                                       "positions_per_launch": n_eval, "ms_per_launch": k4_alone_ms, "achieved": FLOP_PER_POSITION * n_eval / (k4_alone_ms * 1e-3) / 1e12,
                                       "frac": FLOP_PER_POSITION * n_eval / (k4_alone_ms * 1e-3) / 1e12 / peak},
                             "share_of_step": (nn_ns * 1e-6) / ms, "solver_share_of_step": solver_ns * 1e-6 / ms,
                             "leaf_positions_per_step": leaves / args.steps, "network_positions_per_step": evals / args.steps},
                "sharding": {"c1_weight_bytes_broadcast": int(blob.nbytes), "ranks": world,
                             "per_rank_ms_per_step": {"min": -neg_mins[2] / args.steps, "max": max_ms / args.steps},
                             "per_rank_k5_ms_per_step": {"min": -neg_mins[0] * 1e-6 / args.steps, "max": k_maxes[0] * 1e-6 / args.steps},
                             "per_rank_k4_ms_per_step": {"min": -neg_mins[1] * 1e-6 / args.steps, "max": k_maxes[1] * 1e-6 / args.steps}},
                "clocks": clocks.summary()}
        if world == 1 and not args.no_cpu_baseline:
            n_int = max(2, int(round(args.cpu_seconds / 4.0)))
            per_interval, cores, search_rate, _ = run_reference(4.0, n_int + 1, start=args.start)
            v = float(np.mean(per_interval[1:]))
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                                    "sample": f"{n_int} x 4 s (after one 4 s warm-up slice) of the reference's own GeneratorThread::run loop on the host cores from the same "
                                              f"positions (oracle/_ref: GeneratorManager, GameGenerator, Search, Tree, AlphaBetaSearch, NNEvaluator batch 64; one thread per core, "
                                              f"8 games x batch 8 each, solver on; NN = torch-CPU fp32 stand-in for MinML)",
                                    "integer_path_one_core": reference_integer_path(), "integer_path_device": device_integer_path}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
