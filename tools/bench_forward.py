"""Quick device-side timing of agb_forward_dev (K4) with CUDA events on the engine's stream."""
import ctypes, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alphagomoku_b200 as agb
from alphagomoku_b200 import netblob
blocks, filters, q = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
eng = agb.Engine(agb.GameConfig(agb.GameRules.STANDARD, 15, 15), max_boards=n, blocks=blocks, filters=filters, q_head=bool(q))
eng.load_weights(netblob.pack(netblob.random_tensors(15, 15, blocks, filters, bool(q)), 15, 15, blocks, filters, bool(q)))
stream = torch.cuda.ExternalStream(eng.stream())
feats = torch.randint(0, 2**31 - 1, (n, 225), dtype=torch.int32, device="cuda")
policy = torch.empty((n, 225), dtype=torch.float32, device="cuda"); value = torch.empty((n, 3), dtype=torch.float32, device="cuda")
qo = torch.empty((n, 225, 3), dtype=torch.float32, device="cuda")
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
macs = {(20,128): 1383.70e6, (10,64): 185.89e6}.get((blocks, filters), 0) + (q * (225*9*filters*filters + 225*3*filters))
print(f"blocks={blocks} filters={filters} q={q} n={n}: {ms:.3f} ms  -> {n/ms*1e3:.0f} pos/s, {2*macs*n/ms/1e9:.1f} TFLOP/s algorithmic; all={['%.3f'%t for t in times]}")
