"""Host side of the self-play output contract: GameDataBuffer files (format 201).

The per-game blobs are produced on the device (K8, csrc/records.cuh) exactly as GameDataStorage::serialize writes them
(src/dataset/GameDataStorage.cpp:217-251); this module only frames them like GameDataBuffer::save / FileSaver::save do
(src/dataset/GameDataBuffer.cpp:96-112, src/utils/file_util.cpp:42-52): one JSON header line
{"format": 201, "config": GameConfig, "offsets": [...]}, '\\n', the concatenated blobs, the whole file zlib-compressed."""
import json
import struct
import zlib

RULE_NAMES = ["FREESTYLE", "STANDARD", "RENJU", "CARO5", "CARO6"]


def split_records(blob, n_games=None):
    """Walks concatenated GameDataStorage blobs -> list of bytes (one per game)."""
    out, off = [], 0
    while off < len(blob):
        start = off
        (n_samples,) = struct.unpack_from("<I", blob, off)
        off += 4
        for _ in range(n_samples):
            off += 12  # 3 fp16 scales, minimax score, move number, flags
            (n_entries,) = struct.unpack_from("<I", blob, off)
            off += 4 + 6 * n_entries
        (n_moves,) = struct.unpack_from("<I", blob, off)
        off += 4 + 2 * n_moves + 12  # moves, outcome, rows, cols
        out.append(bytes(blob[start:off]))
    assert n_games is None or len(out) == n_games
    return out


def parse_record(rec):
    """One game blob -> dict (for tests and inspection; the trainer-side reader of the reference is out of scope)."""
    off = 0
    (n_samples,) = struct.unpack_from("<I", rec, off)
    off += 4
    samples = []
    for _ in range(n_samples):
        vs, ps, ns, score, move_number, flags, n_entries = struct.unpack_from("<HHHHHHI", rec, off)
        off += 16
        entries = [struct.unpack_from("<6B", rec, off + 6 * i) for i in range(n_entries)]
        off += 6 * n_entries
        samples.append({"value_scale": vs, "policy_scale": ps, "visit_scale": ns, "minimax_score": score, "move_number": move_number,
                        "flags": flags, "entries": entries})
    (n_moves,) = struct.unpack_from("<I", rec, off)
    off += 4
    moves = list(struct.unpack_from(f"<{n_moves}H", rec, off))
    off += 2 * n_moves
    outcome, rows, cols = struct.unpack_from("<iii", rec, off)
    return {"samples": samples, "moves": moves, "outcome": outcome, "rows": rows, "cols": cols}


class GameDataBuffer:
    def __init__(self, rules, rows, cols, draw_after=0):
        self.config = {"rules": RULE_NAMES[int(rules)], "rows": rows, "cols": cols, "draw_after": draw_after or rows * cols}
        self.games = []

    def add_records(self, blob, n_games=None):
        self.games += split_records(blob, n_games)

    def number_of_games(self):
        return len(self.games)

    def save(self, path):
        offsets, off = [], 0
        for g in self.games:
            offsets.append(off)
            off += len(g)
        header = json.dumps({"format": 201, "config": self.config, "offsets": offsets}, separators=(",", ":"))
        with open(path, "wb") as f:
            f.write(zlib.compress(header.encode() + b"\n" + b"".join(self.games)))
