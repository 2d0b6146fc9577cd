"""a22: prepareOpening (src/utils/misc.cpp:142-170) restated in alphagomoku_b200/csrc/openings_logic.hpp draws the same openings as the
reference when both generators start from the same seed. The reference's generator is a thread-local std::mt19937(0) of its debug build
that cannot be reseeded, so each comparison runs in a fresh interpreter."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

SCRIPT = r"""
import ctypes, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from refapi import RefOracle, _p
rules, size, count = {rules}, {size}, {count}
ref = RefOracle()
hs = ctypes.CDLL({hostsim!r})
hs.hostsim_openings_create.restype = ctypes.c_void_p
pattern, threat = np.zeros(1 << 20, np.uint8), np.zeros(4096, np.uint8)
hs.hostsim_pattern_table(rules, _p(pattern)); hs.hostsim_threat_table(rules, _p(threat))
rng = ctypes.c_void_p(hs.hostsim_openings_create(0))
lengths = []
for i in range(count):
    a, b = np.zeros(size * size, np.uint16), np.zeros(size * size, np.uint16)
    na = ref.lib.agref_prepare_opening(rules, size, size, 1, _p(a))
    nb = hs.hostsim_prepare_opening(rng, rules, size, 1, _p(pattern), _p(threat), _p(b))
    assert na == nb and (a[:na] == b[:nb]).all(), (i, a[:na].tolist(), b[:nb].tolist())
    lengths.append(na)
print("OK", sum(lengths), max(lengths))
"""


@pytest.mark.parametrize("rules,size", [(0, 15), (1, 15), (2, 15), (3, 20), (4, 12)])
def test_prepare_opening_matches_reference(ref, hostsim, rules, size):
    script = SCRIPT.format(root=ROOT, tests=os.path.join(ROOT, "tests"), rules=rules, size=size, count=300, hostsim=hostsim._name)
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    status, total, longest = out.stdout.split()[-3:]
    assert status == "OK" and int(total) > 300 and int(longest) <= 15  # 3 x randInt(6) stones at most
