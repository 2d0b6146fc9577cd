"""Probe: how much of the solver kernel's time is instruction supply? Solves 4096 DIFFERENT positions (every warp of an SM at a different
place of the program) and then 4096 copies of ONE position (all warps run the same instruction stream), and compares positions/s."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import alphagomoku_b200 as agb

n = 4096
S = 15
eng = agb.Engine(agb.GameConfig(agb.GameRules.STANDARD, S, S), max_boards=n, blocks=2, filters=64, seed=1)
rng = np.random.default_rng(5)
boards = np.zeros((n, S * S), np.int8)
stm = np.ones(n, np.int8)
for g in range(n):
    k = int(rng.integers(6, 30))
    cells = rng.choice(11 * 11, size=k, replace=False)
    for j, cell in enumerate(cells):
        boards[g, (2 + cell // 11) * S + 2 + cell % 11] = 1 + (j % 2)
    stm[g] = 1 if k % 2 == 0 else 2


def run(b, s, limit):
    eng.solve(b, s, limit)
    t0 = time.perf_counter()
    out = eng.solve(b, s, limit)
    dt = time.perf_counter() - t0
    nodes = (out[4] >> 8).astype(np.int64)
    return dt, nodes


for limit in (100, 1000):
    dt, nodes = run(boards, stm, limit)
    print(f"limit {limit}: distinct   {dt * 1e3:8.2f} ms, positions mean {nodes.mean():7.1f} max {nodes.max()}, {nodes.sum() / dt / 1e6:7.2f} M positions/s")
    order = np.argsort(nodes)
    for q in (0.5, 0.9, 0.99):
        g = order[int(q * (n - 1))]
        dt1, nodes1 = run(np.repeat(boards[g:g + 1], n, 0), np.repeat(stm[g:g + 1], n), limit)
        print(f"limit {limit}: identical  {dt1 * 1e3:8.2f} ms, positions each {nodes1[0]:5d}, {nodes1.sum() / dt1 / 1e6:7.2f} M positions/s")
