import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref():
    """The reference's own classes (oracle/_ref/libagref.so); built here when /root/reference exists, prebuilt on the GPU box."""
    import refapi
    if not refapi.available():
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", os.path.join(ROOT, "oracle", "_ref", "libagref.so")])
        else:
            pytest.skip("oracle/_ref/libagref.so not built and /root/reference absent")
    return refapi.RefOracle()


@pytest.fixture(scope="session")
def ref_fast(ref):
    """The same reference sources with the reference's Release flags (-O3 -DNDEBUG): for inputs on which its debug asserts abort."""
    import refapi
    if not os.path.exists(refapi.REF_LIB_FAST):
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", os.path.join(ROOT, "oracle", "_ref", "libagref_fast.so")])
        else:
            pytest.skip("oracle/_ref/libagref_fast.so not built and /root/reference absent")
    return refapi.RefOracle(fast=True)


@pytest.fixture(scope="session")
def hostsim(tmp_path_factory):
    """Host-compiled kernel logic (tests/hostsim): the per-cell device functions run on the CPU for logic checks."""
    import ctypes
    out = str(tmp_path_factory.mktemp("hostsim") / "hostsim.so")
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", out])
    lib = ctypes.CDLL(out)
    lib.hostsim_open3_promotions.restype = ctypes.c_uint32
    return lib


@pytest.fixture(scope="session")
def golden():
    import json
    gdir = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(gdir, "known_answers.json")) as f:
        known = json.load(f)
    states = dict(np.load(os.path.join(gdir, "reference_states.npz")))
    return known, states


def random_boards(rng, size, count, max_fill=0.6):
    boards = np.zeros((count, size * size), np.int8)
    for i in range(count):
        n = int(rng.integers(0, int(max_fill * size * size) + 1))
        idx = rng.permutation(size * size)[:n]
        boards[i, idx[0::2]] = 1
        boards[i, idx[1::2]] = 2
    return boards
