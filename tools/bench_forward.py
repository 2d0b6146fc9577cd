"""Quick device-side timing of agb_forward_dev (K4) with CUDA events on the engine's stream."""
import ctypes, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alphagomoku_b200 as agb
from alphagomoku_b200 import netblob
blocks, filters, q = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
S = int(sys.argv[6]) if len(sys.argv) > 6 else 15  # board size (20: one board per CTA pair)
C = S * S
eng = agb.Engine(agb.GameConfig(agb.GameRules.STANDARD, S, S), max_boards=n, blocks=blocks, filters=filters, q_head=bool(q))
eng.load_weights(netblob.pack(netblob.random_tensors(S, S, blocks, filters, bool(q)), S, S, blocks, filters, bool(q)))
stream = torch.cuda.ExternalStream(eng.stream())
feats = torch.randint(0, 2**31 - 1, (n, C), dtype=torch.int32, device="cuda")
policy = torch.empty((n, C), dtype=torch.float32, device="cuda"); value = torch.empty((n, 3), dtype=torch.float32, device="cuda")
qo = torch.empty((n, C, 3), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
lib = eng._lib
def run():
    rc = lib.agb_forward_dev(eng._h, ctypes.c_void_p(feats.data_ptr()), n, ctypes.c_void_p(policy.data_ptr()), ctypes.c_void_p(value.data_ptr()), ctypes.c_void_p(qo.data_ptr()))
    assert rc == 0, lib.agb_last_error(eng._h)
for _ in range(3): run()
eng.synchronize()
times = []
with torch.cuda.stream(stream):
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream); run(); e.record(stream); e.synchronize(); times.append(s.elapsed_time(e))
ms = float(np.median(times))
# 2 x MACs per position: stem 5x5x32 -> F, 2 x blocks 3x3 F -> F, policy 3x3 F -> F + 1x1, value 1x1 F -> 4 + dense, Q head
D = min(256, 2 * filters)
macs = C * (25 * 32 * filters + 2 * blocks * 9 * filters * filters + 9 * filters * filters + filters + 4 * filters) + 4 * C * D + 3 * D + q * C * (9 * filters * filters + 3 * filters)
print(f"board={S}x{S} blocks={blocks} filters={filters} q={q} n={n}: {ms:.3f} ms  -> {n/ms*1e3:.0f} pos/s, {2*macs*n/ms/1e9:.1f} TFLOP/s algorithmic; all={['%.3f'%t for t in times]}")
