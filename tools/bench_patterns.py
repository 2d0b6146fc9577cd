"""K1/K2/K3 at scale: N boards through agb_set_boards_dev (K1+K3), agb_add_moves / agb_undo_moves (K2) and agb_encode (K3).
Run under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` for per-kernel time and DRAM traffic; the
script itself prints CUDA-event times of the device-pointer entry point."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alphagomoku_b200 as agb


def random_boards(rng, size, count, max_fill=0.6):
    boards = np.zeros((count, size * size), np.int8)
    for i in range(count):
        n = int(rng.integers(0, int(max_fill * size * size) + 1))
        idx = rng.permutation(size * size)[:n]
        boards[i, idx[0::2]] = 1
        boards[i, idx[1::2]] = 2
    return boards


N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
S = int(sys.argv[2]) if len(sys.argv) > 2 else 15
rules = int(sys.argv[3]) if len(sys.argv) > 3 else 1
eng = agb.Engine(agb.GameConfig(agb.GameRules(rules), S, S), max_boards=N)
rng = np.random.default_rng(0)
base = random_boards(rng, S, 4096, max_fill=0.6)
boards = np.tile(base, (N // 4096, 1))
stm = rng.integers(1, 3, N).astype(np.int8)
d_boards = torch.from_numpy(boards).cuda()
d_stm = torch.from_numpy(stm).cuda()
d_feat = torch.empty((N, S * S), dtype=torch.int32, device="cuda")
stream = torch.cuda.ExternalStream(eng.stream())
lib = eng._lib


def set_boards():
    assert lib.agb_set_boards_dev(eng._h, ctypes.c_void_p(d_boards.data_ptr()), ctypes.c_void_p(d_stm.data_ptr()), N, ctypes.c_void_p(d_feat.data_ptr())) == 0


for _ in range(2):
    set_boards()
eng.synchronize()
times = []
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    set_boards()
    e.record(stream)
    e.synchronize()
    times.append(s.elapsed_time(e))
ms = float(np.median(times))
cells = S * S
print(f"K1+K3 set_boards: N={N} {S}x{S}: {ms:.3f} ms = {N / ms / 1e3:.1f} M positions/s; algorithmic {5 * cells} B/position -> {N * 5 * cells / ms / 1e6:.0f} GB/s; "
      f"with the persistent state written ({11 * cells + 6 * S * 8} B more) -> {N * (16 * cells + 48 * S) / ms / 1e6:.0f} GB/s")
# K2: one random empty cell per board, add then undo (host entry points: the copies are outside the kernels ncu times)
moves = np.zeros(N, np.uint16)
for i in range(4096):
    empty = np.flatnonzero(base[i] == 0)
    c = int(empty[rng.integers(len(empty))]) if len(empty) else 0
    moves[i] = (1 + i % 2) | ((c // S) << 2) | ((c % S) << 9)
moves = np.tile(moves[:4096], N // 4096)
eng.add_moves(moves)
eng.undo_moves(moves)
eng.encode(N)
print("K2 add + undo and K3 encode launched once each (see the ncu launch list)")
eng.close()
