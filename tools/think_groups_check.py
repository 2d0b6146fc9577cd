import sys, numpy as np
sys.path.insert(0, ".")
import alphagomoku_b200 as agb
from alphagomoku_b200 import netblob
import bench
games = 1536
eng = agb.Engine(agb.GameConfig(agb.GameRules.STANDARD, 15, 15), max_boards=games * 4, blocks=2, filters=64, games=games, max_batch_size=4,
                 max_simulations=50, solver_max_positions=50, seed=1)
eng.load_weights(netblob.pack(netblob.random_tensors(15, 15, 2, 64, False), 15, 15, 2, 64, False))
boards, stm = bench.random_openings(np.random.default_rng(1), games)
moves, values = eng.think(boards, stm)
print("think ok", (moves != 0).sum(), eng.stats()["solver_sms"], eng.stats()["overflow_flags"])
eng.selfplay_reset(boards, stm)
eng.step(30)
blob = eng.save_games()
eng.load_games(blob)
eng.step(10)
print("resume ok", eng.stats()["nb_moves_played"], eng.stats()["solver_sms"], eng.stats()["overflow_flags"])
