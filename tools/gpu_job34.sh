#!/bin/bash
mkdir -p /tmp/hosttest gpurun_out
python - <<PY
import numpy as np, sys
sys.path.insert(0,'.')
from alphagomoku_b200 import netblob
blob = netblob.pack(netblob.random_tensors(15, 15, 4, 64, False, seed=11), 15, 15, 4, 64, False)
np.ascontiguousarray(blob, np.float32).tofile('/tmp/hosttest/w.f32')
PY
rm -f gpurun_out/a.eval
AGB_EVAL_DUMP=gpurun_out/a.eval oracle/_ref/agb_host_b200 generator /tmp/hosttest/w.f32 /tmp/hosttest/a_new.bin 6 2>&1 | grep STALE | head -5
ls -la gpurun_out/a.eval
