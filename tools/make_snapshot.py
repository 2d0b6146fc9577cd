"""Steady-state snapshot for bench.py: plays the bench workload from generated openings for `steps` lockstep steps (finished games
restart from the openings pool, so the mixture of game phases becomes stationary), then writes every game's current position.

  python tools/make_snapshot.py <workload> <steps> <out.npz> [games]

The output (boards int8 [games, cells] bit-packed 2 bits per cell, sign to move) is committed under bench_data/ and is what
`bench.py --start snapshot` (the default) resets both arms to. Prints the per-100-step device rates of the soak on the way."""
import struct
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import alphagomoku_b200 as agb
from alphagomoku_b200 import netblob
import bench


def parse_saved_positions(blob, cells):
    """Positions out of an agb_save_games blob (selfplay.cu: AgbSavedHeader, then per game board / sign / moves / samples)."""
    magic, version, games, rows, cols, rules = struct.unpack_from("<IIiiii", blob, 0)
    assert magic == 0x53424741 and rows * cols == cells
    off = 24
    boards, stm = np.zeros((games, cells), np.int8), np.zeros(games, np.int8)
    for g in range(games):
        boards[g] = np.frombuffer(blob, np.int8, cells, off)
        off += cells
        stm[g] = blob[off]
        off += 1
        (n_moves,) = struct.unpack_from("<i", blob, off)
        off += 4 + 2 * n_moves
        _, rec_len = struct.unpack_from("<ii", blob, off)
        off += 8 + rec_len
    return boards, stm


def pack_boards(boards):
    """2 bits per cell, 4 cells per byte (cell i in bits 2*(i%4) of byte i//4)."""
    n, cells = boards.shape
    padded = np.zeros((n, (cells + 3) // 4 * 4), np.uint8)
    padded[:, :cells] = boards
    q = padded.reshape(n, -1, 4)
    return (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)


def unpack_boards(packed, cells):
    n = packed.shape[0]
    out = np.zeros((n, packed.shape[1], 4), np.int8)
    for k in range(4):
        out[:, :, k] = (packed >> (2 * k)) & 3
    return out.reshape(n, -1)[:, :cells].copy()


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "freestyle15"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    out = sys.argv[3] if len(sys.argv) > 3 else f"gpurun_out/steady_{workload}.npz"
    games = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
    bench.select_workload(workload)
    S, nodes = bench.SIZE, 1536 * bench.SIMS // 400
    eng = agb.Engine(agb.GameConfig(agb.GameRules(bench.RULES), S, S), max_boards=games * 8, blocks=bench.BLOCKS, filters=bench.FILTERS, games=games,
                     max_batch_size=8, max_simulations=bench.SIMS, max_nodes_per_game=nodes, max_edges_per_game=nodes * 200, solver_max_positions=100,
                     solver_table_entries=65536, seed=1, use_symmetries=True)
    eng.load_weights(netblob.pack(netblob.random_tensors(S, S, bench.BLOCKS, bench.FILTERS, False), S, S, bench.BLOCKS, bench.FILTERS, False))
    print("workload:", bench.WORKLOAD)
    boards, stm = eng.generate_openings(games)
    eng.selfplay_reset(boards, stm)
    t0 = time.time()
    prev = eng.stats()
    for chunk in range(steps // 100):
        eng.step(100)
        eng.pop_finished()
        st = eng.stats()
        assert st["overflow_flags"] == 0, st
        ms = (time.time() - t0) * 1e3 / (100 * (chunk + 1))
        print(f"step {(chunk + 1) * 100}: {(st['nb_network_evaluations'] - prev['nb_network_evaluations']) / 100:.0f} evals/step, K4 "
              f"{(st['nn_kernel_ns'] - prev['nn_kernel_ns']) / 1e8:.1f} ms/step, K5 {(st['solver_kernel_ns'] - prev['solver_kernel_ns']) / 1e8:.1f} ms/step, "
              f"solver SMs {st['solver_sms']}, finished {st['nb_games_finished']}, wall {ms:.1f} ms/step (cumulative)", flush=True)
        prev = st
    b, s = parse_saved_positions(eng.save_games(), S * S)
    stones = (b != 0).sum(1)
    print("snapshot: stones per position mean %.1f min %d max %d" % (stones.mean(), stones.min(), stones.max()))
    np.savez_compressed(out, boards=pack_boards(b), sign_to_move=s, rows=S, cols=S, rules=bench.RULES, steps=steps)
    eng.close()


if __name__ == "__main__":
    main()
