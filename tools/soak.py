"""Soak test of the lockstep engine at the bench configuration: many steps, finished games popped as they come, no overflow flags."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import alphagomoku_b200 as agb
from alphagomoku_b200 import netblob, dataset
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
games = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
bench.select_workload(sys.argv[3] if len(sys.argv) > 3 else "standard15")
S, nodes = bench.SIZE, 1536 * bench.SIMS // 400
eng = agb.Engine(agb.GameConfig(agb.GameRules(bench.RULES), S, S), max_boards=games * 8, blocks=bench.BLOCKS, filters=bench.FILTERS, games=games,
                 max_batch_size=8, max_simulations=bench.SIMS, max_nodes_per_game=nodes, max_edges_per_game=nodes * 200, solver_max_positions=100,
                 solver_table_entries=65536, seed=1, use_symmetries=True)
eng.load_weights(netblob.pack(netblob.random_tensors(S, S, bench.BLOCKS, bench.FILTERS, False), S, S, bench.BLOCKS, bench.FILTERS, False))
print("workload:", bench.WORKLOAD)
boards, stm = eng.generate_openings(games)
print("openings:", games, "stones per opening mean", float((boards != 0).sum(1).mean()))
eng.selfplay_reset(boards, stm)
t0 = time.time()
total_games, total_bytes, lengths = 0, 0, []
for chunk in range(steps // 100):
    eng.step(100)
    blob, n = eng.pop_finished()
    total_games += n
    total_bytes += len(blob)
    for rec in dataset.split_records(blob, n)[:50]:
        lengths.append(len(dataset.parse_record(rec)["moves"]))
    prev = st if chunk > 0 else None
    st = eng.stats()
    assert st["overflow_flags"] == 0, st
    if prev is not None:
        print(f"  last 100 steps: K4 {(st['nn_kernel_ns'] - prev['nn_kernel_ns']) / 1e8:.1f} ms/step in {(st['nn_kernel_launches'] - prev['nn_kernel_launches']) / 100:.0f} launches, "
              f"K5 {(st['solver_kernel_ns'] - prev['solver_kernel_ns']) / 1e8:.1f} ms/step, network positions {(st['nn_positions'] - prev['nn_positions']) / 100:.0f}/step, "
              f"solver SMs now {st['solver_sms']}")
    print(f"step {(chunk + 1) * 100}: evals {st['nb_network_evaluations']} moves {st['nb_moves_played']} finished {st['nb_games_finished']} popped {total_games} "
          f"({total_bytes / 1e6:.1f} MB) proven {st['nb_proven_states']} leaks {st['nb_information_leaks']} t={time.time() - t0:.0f}s", flush=True)
print("game length (sample): mean", np.mean(lengths) if lengths else None, "max", max(lengths) if lengths else None)
eng.close()
