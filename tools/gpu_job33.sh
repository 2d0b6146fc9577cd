#!/bin/bash
mkdir -p /tmp/hosttest
python - <<PY
import numpy as np, sys
sys.path.insert(0,'.')
from alphagomoku_b200 import netblob
blob = netblob.pack(netblob.random_tensors(15, 15, 4, 64, False, seed=11), 15, 15, 4, 64, False)
np.ascontiguousarray(blob, np.float32).tofile('/tmp/hosttest/w.f32')
PY
AGB_NET_DUMP=/tmp/hosttest/a.dump AGB_EVAL_DUMP=/tmp/hosttest/a.eval oracle/_ref/agb_host_b200 generator /tmp/hosttest/w.f32 /tmp/hosttest/a_new.bin 6 | tail -1
AGB_NET_DUMP=/tmp/hosttest/b.dump oracle/_ref/agb_host_shadow generator /tmp/hosttest/w.f32 /tmp/hosttest/b_new.bin 6 | tail -1
python tools/k4_sym_check.py /tmp/hosttest/a.dump /tmp/hosttest/a.eval /tmp/hosttest/b.dump 2>&1 | tail -40
