"""One K5 launch at steady state for ncu: resets the bench workload to a snapshot (tools/make_snapshot.py), plays `warm` steps, then a few more.
  ncu --set full --import-source on -k regex:solve_games -s <2 * warm> -c 1 python tools/profile_solver.py <snapshot.npz> <warm> [workload]"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import alphagomoku_b200 as agb
from alphagomoku_b200 import netblob
import bench
from make_snapshot import unpack_boards

snap = np.load(sys.argv[1])
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 60
bench.select_workload(sys.argv[3] if len(sys.argv) > 3 else "freestyle15")
S = bench.SIZE
boards, stm = unpack_boards(snap["boards"], S * S), snap["sign_to_move"]
games, nodes = boards.shape[0], 1536 * bench.SIMS // 400
eng = agb.Engine(agb.GameConfig(agb.GameRules(bench.RULES), S, S), max_boards=games * 8, blocks=bench.BLOCKS, filters=bench.FILTERS, games=games,
                 max_batch_size=8, max_simulations=bench.SIMS, max_nodes_per_game=nodes, max_edges_per_game=nodes * 200, solver_max_positions=100,
                 solver_table_entries=65536, seed=1, use_symmetries=True)
eng.load_weights(netblob.pack(netblob.random_tensors(S, S, bench.BLOCKS, bench.FILTERS, False), S, S, bench.BLOCKS, bench.FILTERS, False))
eng.selfplay_reset(boards, stm)
eng.step(warm)
st0 = eng.stats()
eng.step(4)
st = eng.stats()
print({k: st[k] - st0[k] for k in ("nb_network_evaluations", "nb_node_count", "nn_kernel_ns", "solver_kernel_ns")}, "solver_sms", st["solver_sms"])
eng.close()
