"""Network files in the reference's envelope (FileSaver / FileLoader, src/utils/file_util.cpp:42-118; AGNetwork::saveToFile / loadFrom,
src/networks/AGNetwork.cpp:167-192)."""
import ctypes
import json

import numpy as np
import pytest


def test_envelope_is_the_references(ref, tmp_path):
    """Files written here are split by the reference's FileLoader exactly where we put the binary; files written by the reference's FileSaver
    (plain and zlib-compressed) read back here."""
    import refapi
    from alphagomoku_b200 import netfile
    rng = np.random.default_rng(3)
    binary = rng.integers(0, 256, 5000, dtype=np.uint8).tobytes() + b"{}[]\n{"  # braces in the binary part must not confuse the split
    obj = {"architecture": "ResnetPV", "config": {"rules": "STANDARD", "rows": 15, "cols": 15, "draw_after": 225}, "model": {"nodes": [1, 2, {"a": [3]}]}}
    mine = str(tmp_path / "mine.bin")
    netfile.write_envelope(mine, obj, binary)
    ref.lib.agref_file_load.restype = ctypes.c_long
    ref.lib.agref_file_load.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
    text, out = ctypes.create_string_buffer(1 << 16), np.zeros(1 << 16, np.uint8)
    n = ref.lib.agref_file_load(mine.encode(), 0, text, len(text), refapi._p(out), out.size)
    assert n == len(binary) and out[:n].tobytes() == binary and json.loads(text.value.decode()) == obj
    for compress in (0, 1):
        theirs = str(tmp_path / f"theirs{compress}.bin")
        ref.lib.agref_file_save.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
        assert ref.lib.agref_file_save(theirs.encode(), json.dumps(obj).encode(), binary, len(binary), 2, compress) == 0
        got, blob = netfile.read_envelope(theirs, uncompress=bool(compress))
        assert got == obj and blob == binary


def test_network_file_round_trip_and_architecture_check(tmp_path):
    from alphagomoku_b200 import netblob, netfile
    tensors = netblob.random_tensors(15, 15, 3, 64, True, seed=9)
    path = str(tmp_path / "network.bin")
    netfile.save_network_file(path, tensors, "RENJU", 15, 15, 3, 64, True)
    net = netfile.load_network_file(path)
    assert net["architecture"] == "ResnetPVQ" and net["blocks"] == 3 and net["filters"] == 64 and net["q_head"] and net["game_config"]["rules"] == "RENJU"
    assert (net["blob"] == netblob.pack(tensors, 15, 15, 3, 64, True)).all()
    obj, binary = netfile.read_envelope(path)
    obj["architecture"] = "ConvNextPVQMraw"  # what the author trains now (TrainingConfig default, configs.hpp:169)
    netfile.write_envelope(path, obj, binary)
    with pytest.raises(ValueError, match="ConvNextPVQMraw"):
        netfile.load_network_file(path)
    obj["architecture"], obj["model"] = "ResnetPV", {"nodes": []}  # a MinML graph: the one missing function says so
    netfile.write_envelope(path, obj, binary)
    with pytest.raises(NotImplementedError, match="MinML"):
        netfile.load_network_file(path)
