"""Per-game cost of one steady-state K5 launch: SM clocks and search nodes per game (agb_debug_solver_load), from a snapshot."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import alphagomoku_b200 as agb
from alphagomoku_b200 import netblob
import bench
from make_snapshot import unpack_boards

snap = np.load(sys.argv[1])
sms = int(sys.argv[2]) if len(sys.argv) > 2 else 56
bench.select_workload("freestyle15")
S = 15
boards, stm = unpack_boards(snap["boards"], S * S), snap["sign_to_move"]
games, nodes = boards.shape[0], 1536
eng = agb.Engine(agb.GameConfig(agb.GameRules(0), S, S), max_boards=games * 8, blocks=bench.BLOCKS, filters=bench.FILTERS, games=games,
                 max_batch_size=8, max_simulations=400, max_nodes_per_game=nodes, max_edges_per_game=nodes * 200, solver_max_positions=100,
                 solver_table_entries=65536, seed=1, use_symmetries=True, solver_sms=sms, pipeline_groups=2)
eng.load_weights(netblob.pack(netblob.random_tensors(S, S, bench.BLOCKS, bench.FILTERS, False), S, S, bench.BLOCKS, bench.FILTERS, False))
eng.selfplay_reset(boards, stm)
eng.step(60)
prev = None
for rep in range(3):
    ns0 = eng.stats()["solver_kernel_ns"]
    eng.step(1)
    kernel_ms = (eng.stats()["solver_kernel_ns"] - ns0) * 1e-6
    out = np.zeros((games, 2), np.uint64)
    eng._lib.agb_debug_solver_load.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    assert eng._lib.agb_debug_solver_load(eng._h, out.ctypes.data_as(ctypes.c_void_p)) == 0
    cyc, nd = out[:, 0].astype(np.float64) / 1.95e6, out[:, 1].astype(np.float64)
    q = np.percentile(cyc, [10, 50, 75, 90, 95, 99, 100])
    order = np.sort(cyc)[::-1]
    share = np.cumsum(order) / order.sum()
    print(f"launch pair {rep}: K5 {kernel_ms:.1f} ms (both groups); ms per game p10/p50/p75/p90/p95/p99/max = " + " / ".join(f"{v:.1f}" for v in q)
          + f"; nodes per game mean {nd.mean():.0f}; heaviest 5 % of the games hold {100 * share[games // 20]:.0f} % of the warp time, 10 %: {100 * share[games // 10]:.0f} %, 25 %: {100 * share[games // 4]:.0f} %")
    if prev is not None:
        print("   correlation with the previous step's per-game time: %.2f" % np.corrcoef(prev, cyc)[0, 1])
    prev = cyc
eng.close()
