#!/bin/bash
mkdir -p /tmp/hosttest
python - <<PY
import numpy as np, sys
sys.path.insert(0,'.')
from alphagomoku_b200 import netblob
blob = netblob.pack(netblob.random_tensors(15, 15, 4, 64, False, seed=11), 15, 15, 4, 64, False)
np.ascontiguousarray(blob, np.float32).tofile('/tmp/hosttest/w.f32')
PY
for i in 1 2 3; do oracle/_ref/agb_host_b200 generator /tmp/hosttest/w.f32 /tmp/hosttest/a$i.bin 6 | tail -1; done
for i in 1 2 3; do oracle/_ref/agb_host_shadow generator /tmp/hosttest/w.f32 /tmp/hosttest/b$i.bin 6 | tail -1; done
md5sum /tmp/hosttest/*.bin; ls -la /tmp/hosttest/*.bin
