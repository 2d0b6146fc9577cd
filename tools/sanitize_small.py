"""A few lockstep steps of a small engine (solver on, pipeline groups and the SM partition on) for compute-sanitizer:
  compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
import alphagomoku_b200 as agb
from alphagomoku_b200 import netblob

S = 15
for groups, sms in ((1, 0), (2, 8)):
    eng = agb.Engine(agb.GameConfig(agb.GameRules.STANDARD, S, S), max_boards=64 * 8, blocks=2, filters=64, games=64, max_batch_size=8, max_simulations=60,
                     max_nodes_per_game=512, max_edges_per_game=512 * 200, solver_max_positions=100, solver_table_entries=4096, seed=1, use_symmetries=True,
                     pipeline_groups=groups, solver_sms=sms)
    eng.load_weights(netblob.pack(netblob.random_tensors(S, S, 2, 64, False), S, S, 2, 64, False))
    rng = np.random.default_rng(0)
    boards = np.zeros((64, S * S), np.int8)
    for g in range(64):
        cells = rng.permutation(S * S)[:10]
        boards[g, cells[:5]] = 1
        boards[g, cells[5:]] = 2
    eng.selfplay_reset(boards, np.ones(64, np.int8))
    eng.step(6)
    print("groups", groups, "ok", eng.stats()["nb_network_evaluations"])
    eng.close()
