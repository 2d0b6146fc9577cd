"""Steady-state rates of the lockstep engine from a snapshot (tools/make_snapshot.py): K4 / K5 device time per step, evaluations per second.
  python tools/steady_bench.py <snapshot.npz> [warm=60] [steps=40] [solver_sms=0 (automatic)] [workload=freestyle15] [groups=0]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import alphagomoku_b200 as agb
from alphagomoku_b200 import netblob
import bench
from make_snapshot import unpack_boards

snap = np.load(sys.argv[1])
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 60
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
solver_sms = int(sys.argv[4]) if len(sys.argv) > 4 else 0
bench.select_workload(sys.argv[5] if len(sys.argv) > 5 else "freestyle15")
groups = int(sys.argv[6]) if len(sys.argv) > 6 else 0
S = bench.SIZE
boards, stm = unpack_boards(snap["boards"], S * S), snap["sign_to_move"]
games, nodes = boards.shape[0], 1536 * bench.SIMS // 400
eng = agb.Engine(agb.GameConfig(agb.GameRules(bench.RULES), S, S), max_boards=games * 8, blocks=bench.BLOCKS, filters=bench.FILTERS, games=games,
                 max_batch_size=8, max_simulations=bench.SIMS, max_nodes_per_game=nodes, max_edges_per_game=nodes * 200, solver_max_positions=100,
                 solver_table_entries=65536, seed=1, use_symmetries=True, solver_sms=solver_sms, pipeline_groups=groups)
eng.load_weights(netblob.pack(netblob.random_tensors(S, S, bench.BLOCKS, bench.FILTERS, False), S, S, bench.BLOCKS, bench.FILTERS, False))
eng.selfplay_reset(boards, stm)
for _ in range(warm // 20):
    eng.step(20)
st0 = eng.stats()
t0 = time.time()
for _ in range(steps // 20):
    eng.step(20)
eng.synchronize()
dt = time.time() - t0
st = eng.stats()
d = {k: st[k] - st0[k] for k in st}
print(f"solver_sms arg {solver_sms} groups {groups}: {d['nb_network_evaluations'] / dt:.0f} evals/s, {dt / steps * 1e3:.1f} ms/step wall, K4 {d['nn_kernel_ns'] / steps / 1e6:.1f} ms/step, "
      f"K5 {d['solver_kernel_ns'] / steps / 1e6:.1f} ms/step, leaves/step {d['nb_node_count'] / steps:.0f}, evals/step {d['nb_network_evaluations'] / steps:.0f}, "
      f"solver SMs now {st['solver_sms']}, overflow {st['overflow_flags']}", flush=True)
eng.close()
