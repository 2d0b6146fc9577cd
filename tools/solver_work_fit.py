"""Calibration of the solver's scheduling weights: with one game per SM a game's clocks are its own cost; fit them on the work counters.
Needs a build whose kernel packs the counters into game_cycles[2g+1] (nodes | adds << 16 | quiet << 32 | generated actions << 48); the
shipped kernel stores the fitted estimate in SolverState::game_work instead. Result of the fit (B200, standard 15x15, 100 positions):
clocks = 23.6 k x moves added + 6.0 k x quiet child visits + 4.3 k x generated actions, R^2 0.97; step-to-step correlation 0.91."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import alphagomoku_b200 as agb
from alphagomoku_b200 import netblob
import bench

games, batch = 148, 8
eng = agb.Engine(agb.GameConfig(agb.GameRules(bench.RULES), 15, 15), max_boards=games * batch, blocks=2, filters=64, games=games, max_batch_size=batch,
                 max_simulations=400, max_nodes_per_game=1536, max_edges_per_game=1536 * 200, solver_max_positions=100, solver_table_entries=65536, seed=1)
eng.load_weights(netblob.pack(netblob.random_tensors(15, 15, 2, 64, False), 15, 15, 2, 64, False))
boards, stm = bench.random_openings(np.random.default_rng(99), games)
eng.selfplay_reset(boards, stm)
rows = []
prev = None
corr = []
for step in range(120):
    eng.step(1)
    out = np.zeros((games, 2), np.uint64)
    eng._lib.agb_debug_solver_load.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    assert eng._lib.agb_debug_solver_load(eng._h, out.ctypes.data_as(ctypes.c_void_p)) == 0
    cyc = out[:, 0].astype(np.float64)
    w = out[:, 1]
    feats = np.stack([(w & np.uint64(0xFFFF)), (w >> np.uint64(16)) & np.uint64(0xFFFF), (w >> np.uint64(32)) & np.uint64(0xFFFF), (w >> np.uint64(48)) & np.uint64(0xFFFF),
                      np.ones(games, np.uint64) * batch], 1).astype(np.float64)
    rows.append((feats, cyc))
    if prev is not None:
        corr.append(np.corrcoef(prev, cyc)[0, 1])
    prev = cyc
X = np.concatenate([r[0] for r in rows])
y = np.concatenate([r[1] for r in rows])
coef, *_ = np.linalg.lstsq(X, y, rcond=None)
pred = X @ coef
print("features: nodes, adds, quiet visits, generated actions, leaves; clocks per unit:", np.round(coef, 1))
print("fit R^2", 1 - ((y - pred) ** 2).sum() / ((y - y.mean()) ** 2).sum(), "mean clocks", y.mean(), "max", y.max())
print("step-to-step correlation of a game's cost: mean", np.mean(corr), "min", np.min(corr))
