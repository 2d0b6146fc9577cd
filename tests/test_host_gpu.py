"""The boundary, run rather than compiled: the reference's own C++ host classes on top of libagb200.so, on the GPU box.

oracle/_ref/agb_host_b200 and agb_host_shadow are built where /root/reference exists (make -C oracle host, part of build()) from
tests/host/host_driver.cpp; they travel to the GPU box like the other prebuilt checkers."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B200 = os.path.join(ROOT, "oracle", "_ref", "agb_host_b200")
SHADOW = os.path.join(ROOT, "oracle", "_ref", "agb_host_shadow")


def _weights(tmp_path):
    from alphagomoku_b200 import netblob
    blob = netblob.pack(netblob.random_tensors(15, 15, 4, 64, False, seed=11), 15, 15, 4, 64, False)
    path = str(tmp_path / "weights.f32")
    np.ascontiguousarray(blob, np.float32).tofile(path)
    return path


def _run(binary, *args):
    if not os.path.exists(binary):
        pytest.skip(f"{binary} not built (needs /root/reference: make -C oracle host)")
    out = subprocess.run([binary, *args], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    return out.stdout


def test_reference_generator_thread_on_the_b200_evaluator(tmp_path):
    """GeneratorManager::generate -> GeneratorThread::run (GeneratorManager.cpp:120-141) -> GameGenerator::generate -> Search / Tree /
    AlphaBetaSearch, all unmodified reference code, with integration/NNEvaluator_b200.cpp in place of the reference's NNEvaluator.cpp: the
    double-buffered asyncEvaluateGraphLaunch / Join schedule, useSymmetries, isQueueFull, the opening generator's own evaluations. The same
    program with the reference's evaluator (its network = the device network called on the reference's features) must produce the same
    games: identical buffers, byte for byte."""
    weights = _weights(tmp_path)
    a, b = str(tmp_path / "b200.bin"), str(tmp_path / "shadow.bin")
    out_a = _run(B200, "generator", weights, a, "6")
    out_b = _run(SHADOW, "generator", weights, b, "6")
    last_a, last_b = out_a.strip().splitlines()[-1], out_b.strip().splitlines()[-1]  # the rest is GeneratorManager::printStats (timings differ)
    assert last_a == last_b and last_a.startswith("games "), (last_a, last_b)
    assert int(last_a.split()[1]) >= 6
    assert open(a, "rb").read() == open(b, "rb").read()


def test_generator_manager_over_the_device_engine(tmp_path, ref):
    """GeneratorManagerB200 (integration/agb200_shims.hpp): generate / getGameBuffer / saveState / loadState over agb_step and
    agb_pop_finished; its buffer files load in the reference's GameDataBuffer."""
    import refapi
    weights = _weights(tmp_path)
    out = _run(B200, "device", weights, str(tmp_path / "work"), "48")
    lines = [line for line in out.strip().splitlines() if line.startswith(("first half", "resumed with"))]
    assert lines[0].startswith("first half: games ") and lines[1].startswith("resumed with ")
    first = int(lines[0].split()[3])
    resumed_with, final = int(lines[1].split()[2]), int(lines[1].split()[7])
    assert first >= 24 and resumed_with == first and final >= 48
    for name, expect in (("saved_state/buffer.bin", first), ("buffer_final.bin", final)):
        spg, mpg, outcomes = np.zeros(4096, np.int32), np.zeros(4096, np.int32), np.zeros(4096, np.int32)
        rows, cols, rules = np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.int32)
        n = ref.lib.agref_buffer_load(str(tmp_path / "work" / name).encode(), refapi._p(spg), refapi._p(mpg), refapi._p(outcomes), 4096, refapi._p(rows), refapi._p(cols),
                                      refapi._p(rules))
        assert n == expect and (spg[:n] > 0).all() and (outcomes[:n] >= 1).all() and rows[0] == 15 and rules[0] == 1
