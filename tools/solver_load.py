"""Load balance of the solver kernel (K5): per-game SM clocks and positions visited by one launch, after a few lockstep steps."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import alphagomoku_b200 as agb
from alphagomoku_b200 import netblob
import bench

games = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
eng = agb.Engine(agb.GameConfig(agb.GameRules(bench.RULES), 15, 15), max_boards=games * batch, blocks=2, filters=64, games=games, max_batch_size=batch,
                 max_simulations=400, max_nodes_per_game=1536, max_edges_per_game=1536 * 200, solver_max_positions=100, solver_table_entries=65536, seed=1)
eng.load_weights(netblob.pack(netblob.random_tensors(15, 15, 2, 64, False), 15, 15, 2, 64, False))
boards, stm = bench.random_openings(np.random.default_rng(99), games)
eng.selfplay_reset(boards, stm)
for steps in (1, 5, 20):
    eng.step(steps - 1)
    ns0 = eng.stats()["solver_kernel_ns"]
    eng.step(1)
    kernel_ms = (eng.stats()["solver_kernel_ns"] - ns0) * 1e-6
    out = np.zeros((games, 2), np.uint64)
    eng._lib.agb_debug_solver_load.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    assert eng._lib.agb_debug_solver_load(eng._h, out.ctypes.data_as(ctypes.c_void_p)) == 0
    cyc, nodes = out[:, 0].astype(np.float64), out[:, 1].astype(np.float64)
    q = np.percentile(cyc, [10, 50, 90, 99, 100]) / 1.9e6
    print(f"after {steps:3d} more steps: ms per game p10/p50/p90/p99/max = " + " / ".join(f"{v:.2f}" for v in q) + f"; positions per game mean {nodes.mean():.0f} max {nodes.max():.0f}; "
          f"cycles per position p50 {np.median(cyc / np.maximum(nodes, 1)):.0f}; kernel {kernel_ms:.1f} ms, sum of per-game time / (SMs x kernel) = "
          f"{cyc.sum() / 1.9e6 / (148 * kernel_ms):.1f} warps busy per SM on average")
