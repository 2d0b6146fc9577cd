"""Games in flight in the reference's saved_state/thread_<i>.bin layout (GeneratorManager::saveState / loadState,
src/selfplay/GeneratorManager.cpp:240-290; GameGenerator::save / load, GameGenerator.cpp:122-141)."""
import ctypes
import os
import struct

import numpy as np


def _in_flight_games(n, rng):
    """Synthetic engine state: random move lists and, as records, the sample part of games the reference's own search played."""
    import refapi
    cells = 225

    def evaluate(features):
        k = features.shape[0]
        p = rng.random((k, cells)).astype(np.float32)
        p /= p.sum(1, keepdims=True)
        w = rng.random(k).astype(np.float32)
        return p, np.stack([w, (1 - w) * 0.5, (1 - w) * 0.5], 1).astype(np.float32), None

    sp = refapi.RefSelfplay(1, 15, evaluate, max_batch_size=4, max_simulations=60, use_solver=True, solver_max_positions=50, draw_after=18)
    board = np.zeros(cells, np.int8)
    board[[7 * 15 + 7, 7 * 15 + 8, 8 * 15 + 7]] = [1, 2, 1]
    sp.set_position(board, 2)
    for _ in range(4000):
        if sp.step() == 2:
            break
    rec = sp.record()
    sp.close()
    (n_samples,) = struct.unpack_from("<I", rec, 0)
    end = 4
    for _ in range(n_samples):
        (n_entries,) = struct.unpack_from("<I", rec, end + 12)
        end += 16 + 6 * n_entries
    games = []
    for g in range(n):
        k = int(rng.integers(0, 30))
        order = rng.permutation(cells)[:k]
        moves = [(1 + (i % 2)) | (int(c) // 15 << 2) | (int(c) % 15 << 9) for i, c in enumerate(order)]
        b = np.zeros(cells, np.int8)
        for m in moves:
            b[((m >> 2) & 127) * 15 + ((m >> 9) & 127)] = m & 3
        games.append({"board": b, "sign_to_move": 1 if k % 2 == 0 else 2, "moves": moves, "samples": n_samples if g % 3 else 0,
                      "record": rec[:end] if g % 3 else struct.pack("<I", 0)})
    return games


def test_saved_state_files_load_in_the_reference_and_come_back(ref, tmp_path):
    import refapi
    from alphagomoku_b200 import saved_state
    rng = np.random.default_rng(21)
    games = _in_flight_games(11, rng)
    blob = saved_state.build_engine_blob(1, 15, 15, games)
    files = saved_state.write_reference_state(str(tmp_path), blob, games_per_thread=4)
    assert [os.path.basename(f) for f in files] == ["thread_0.bin", "thread_1.bin", "thread_2.bin"]
    ref.lib.agref_saved_state_roundtrip.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]
    back = tmp_path / "back" / "saved_state"
    os.makedirs(back)
    seen = 0
    for t, f in enumerate(files):
        info = np.zeros(3 * 8, np.int32)
        n = ref.lib.agref_saved_state_roundtrip(f.encode(), str(back / f"thread_{t}.bin").encode(), refapi._p(info), 8)
        assert n == min(4, 11 - 4 * t)
        for i in range(n):  # the reference's Game and GameDataStorage see what the engine held
            g = games[seen + i]
            assert info[3 * i] == len(g["moves"]) and info[3 * i + 1] == g["samples"] and info[3 * i + 2] == g["sign_to_move"]
        seen += n
    # what the reference wrote (its own Game::serialize / GameDataStorage::serialize / FileSaver) converts back to the same engine state
    again = saved_state.read_reference_state(str(tmp_path / "back"), 1, 15, 15)
    assert again == blob
    header, parsed = saved_state.parse_engine_blob(again)
    assert header["games"] == 11 and all((p["board"] == g["board"]).all() for p, g in zip(parsed, games))
    assert saved_state.move_text(1 | (7 << 2) | (7 << 9)) == "Xh7" and saved_state.move_from_text("Oa14") == (2 | (14 << 2))
