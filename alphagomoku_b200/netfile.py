"""Network files in the reference's envelope: `<json>\\n<binary>` as written by FileSaver::save and read by FileLoader
(src/utils/file_util.cpp:42-52, 63-118), with the keys AGNetwork::saveToFile / loadFrom use (src/networks/AGNetwork.cpp:167-192):
"architecture" (AGNetwork::name()), "config" (GameConfig::toJson) and "model" (MinML's ml::Graph::save) over a binary blob of tensors.

What is here: the envelope reader / writer (bit-compatible with the reference's, checked against it in tests/test_netfile_cpu.py), the
architecture / GameConfig checks of AGNetwork::loadFrom, and a tensor table of our own ("agb200-tensors/1") inside the "model" key so that
networks round-trip through files of the reference's shape today.
What is not: MinML's graph schema. The reference's tree does not contain MinML (CMakeLists.txt:8, 56-79), so how ml::Graph::save lays out
its nodes and weights cannot be read from it; `minml_model_to_tensors` is the ONE function a maintainer with MinML's source fills in -- it
receives the "model" JSON and the binary blob and returns the tensor dict `netblob.pack` consumes (names and shapes: INTEGRATION.md)."""
import json
import zlib

import numpy as np

from . import netblob

SUPPORTED = {"ResnetPV": False, "ResnetPVQ": True}  # architecture -> has a Q head (src/networks/networks.cpp:63-66, 135-138)
TENSOR_FORMAT = "agb200-tensors/1"


def find_split_point(data):
    """FileLoader::find_split_point (file_util.cpp:102-118): the JSON ends where the brace count returns to zero, plus the newline."""
    if data[:4] == b"null":
        return 5
    opened = 0
    for i, ch in enumerate(data):
        if ch in (0x7B, 0x5B):  # { [
            opened += 1
        if ch in (0x7D, 0x5D):  # } ]
            opened -= 1
        if opened == 0:
            return i + 2
    return len(data)


def read_envelope(path, uncompress=False):
    """FileLoader(path, uncompress): -> (json object, binary bytes)."""
    with open(path, "rb") as f:
        data = f.read()
    if uncompress:
        data = zlib.decompress(data)
    split = min(len(data), find_split_point(data))
    return json.loads(data[:split].decode()), data[split:]


def write_envelope(path, obj, binary=b"", indent=2, compress=False):
    """FileSaver::save(json, binary_data, indent, compress)."""
    data = json.dumps(obj, indent=indent).encode() + b"\n" + bytes(binary)
    if compress:
        data = zlib.compress(data)
    with open(path, "wb") as f:
        f.write(data)


def minml_model_to_tensors(model, binary, rows, cols):
    """The MinML-dependent part: ml::Graph::save's node list + weight blob -> {"stem.w": ..., "block0.w1": ..., ...} with batch norm folded
    (AGNetwork::optimize, AGNetwork.cpp:136-149). Not derivable from the reference's tree; see the module docstring."""
    raise NotImplementedError("this file holds a MinML graph (ml::Graph::save); converting it needs MinML's serialisation schema, which is not part of "
                              "the reference's tree. Implement alphagomoku_b200.netfile.minml_model_to_tensors (INTEGRATION.md, 'network files').")


def load_network_file(path):
    """AGNetwork::loadFromFile for the engine: -> dict(architecture, game_config, blocks, filters, q_head, tensors, blob)."""
    obj, binary = read_envelope(path)
    for key in ("architecture", "config", "model"):
        if key not in obj:
            raise ValueError(f"{path}: not a network file (no '{key}' key)")
    arch = obj["architecture"]
    if arch not in SUPPORTED:
        raise ValueError(f"{path}: saved model has architecture '{arch}'; the device engine runs {sorted(SUPPORTED)} (ConvNext / NNUE families are out of scope)")
    cfg = obj["config"]
    rows, cols = int(cfg["rows"]), int(cfg["cols"])
    model = obj["model"]
    if isinstance(model, dict) and model.get("format") == TENSOR_FORMAT:
        tensors = {}
        for t in model["tensors"]:
            count = int(np.prod(t["shape"]))
            tensors[t["name"]] = np.frombuffer(binary, np.float32, count, t["offset"]).reshape(t["shape"]).copy()
        blocks, filters = int(model["blocks"]), int(model["filters"])
    else:
        tensors = minml_model_to_tensors(model, binary, rows, cols)
        blocks = sum(1 for name in tensors if name.endswith(".w1") and name.startswith("block"))
        filters = int(tensors["stem.w"].shape[0])
    q_head = SUPPORTED[arch]
    return {"architecture": arch, "game_config": cfg, "blocks": blocks, "filters": filters, "q_head": q_head, "tensors": tensors,
            "blob": netblob.pack(tensors, rows, cols, blocks, filters, q_head)}


def save_network_file(path, tensors, rules, rows, cols, blocks, filters, q_head):
    """AGNetwork::saveToFile with the tensor table in place of MinML's graph."""
    table, chunks, offset = [], [], 0
    for name, value in tensors.items():
        a = np.ascontiguousarray(value, np.float32)
        table.append({"name": name, "shape": list(a.shape), "offset": offset})
        chunks.append(a.tobytes())
        offset += a.nbytes
    obj = {"architecture": "ResnetPVQ" if q_head else "ResnetPV",
           "config": {"rules": rules, "rows": rows, "cols": cols, "draw_after": rows * cols},
           "model": {"format": TENSOR_FORMAT, "blocks": blocks, "filters": filters, "tensors": table}}
    write_envelope(path, obj, b"".join(chunks), indent=2)
