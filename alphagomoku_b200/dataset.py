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


# ---- trainer-side reader: what GameDataStorage::getSample / SearchDataStorage_v201::storeTo (src/dataset/GameDataStorage.cpp:109-148,
# src/dataset/SearchDataStorage.cpp:237-274) and SamplerVisits::prepare_training_data (src/dataset/Sampler.cpp:96-131) make of a record.
# float32 arithmetic step by step like the reference, so that the decoded fields are bit-identical (tests/test_oracle_cpu.py).
def _lowfp_to_fp32(x, sign_bits, exp_bits, man_bits, bias):
    """LowFP<S,E,M,B>::convert_to_fp32 (include/alphagomoku/utils/low_precision.hpp:141-148)."""
    import numpy as np
    x = np.asarray(x, np.uint32)
    exponent = ((x >> man_bits) & ((1 << exp_bits) - 1)).astype(np.int32) + bias
    base = (x & ((1 << man_bits) - 1)).astype(np.float32) / np.float32(1 << man_bits)
    subnormal = (exponent == bias).astype(np.int32)
    value = (np.float32(1) - subnormal.astype(np.float32) + base) * np.ldexp(np.float32(1), exponent + subnormal).astype(np.float32)
    if sign_bits:
        value = np.where((x >> (exp_bits + man_bits)) & 1, -value, value)
    return value.astype(np.float32)


def _int8_to_score(x):
    """int8_to_score (SearchDataStorage.cpp:32-49) -> Score::to_short."""
    pv, ev = x >> 6, x & 63
    if pv == 0:
        return (0 << 13) | (4000 + ev)
    if pv == 1:
        return (1 << 13) | (4000 + ev)
    if pv == 3:
        return (3 << 13) | (4000 - ev)
    import numpy as np
    e = int(np.float32(1000.0) * _lowfp_to_fp32(ev, 1, 3, 2, -8) + np.float32(0.5))
    return (2 << 13) | (4000 + e)


def _score_is_proven(s):
    return ((s >> 13) & 3) != 2 and s not in (0x0000, 0xFFFF)


def _score_to_value(s):
    pv = (s >> 13) & 3
    return (1.0, 0.0) if pv == 3 else ((0.0, 1.0) if pv == 1 else (0.0, 0.0))


def decode_sample(rec, index, rows=None, cols=None):
    """One sample of a game record -> dict of numpy arrays / scalars with the fields of SearchDataPack. rows / cols: the GameConfig of the
    buffer the record sits in (the reference's generator leaves the record's own size fields at zero, GameGenerator.hpp:39)."""
    import numpy as np
    game = parse_record(rec) if isinstance(rec, (bytes, bytearray)) else rec
    rows, cols = rows or game["rows"], cols or game["cols"]
    cells = rows * cols
    s = game["samples"][index]
    value_scale, policy_scale, visit_scale = (_lowfp_to_fp32(s[k], 0, 5, 11, -16) for k in ("value_scale", "policy_scale", "visit_scale"))
    board = np.zeros(cells, np.int8)
    for mv in game["moves"][:s["move_number"]]:
        board[((mv >> 2) & 127) * cols + ((mv >> 9) & 127)] = mv & 3
    visits, prior = np.zeros(cells, np.int32), np.zeros(cells, np.float32)
    values, scores = np.zeros((cells, 2), np.float32), np.full(cells, (2 << 13) | 4000, np.uint16)
    idx, sum_visits = 0, 0
    win_rate, draw_rate = np.float32(0), np.float32(0)
    for delta, v, p, sc, w, d in s["entries"]:
        idx += delta
        vf = _lowfp_to_fp32(v, 0, 3, 5, -8) * visit_scale + np.float32(0.5)
        visits[idx] = int(vf)
        win, draw = _lowfp_to_fp32(w, 0, 4, 4, -16) * value_scale, _lowfp_to_fp32(d, 0, 4, 4, -16) * value_scale
        total = np.float32(win + draw)
        if total > np.float32(1.0):  # get_valid_value
            win, draw = np.float32(win / total), np.float32(draw / total)
        values[idx] = (win, draw)
        scores[idx] = _int8_to_score(sc)
        prior[idx] = _lowfp_to_fp32(p, 0, 4, 4, -16) * policy_scale
        sum_visits += int(vf)
        win_rate = np.float32(win_rate + np.float32(win * vf))
        draw_rate = np.float32(draw_rate + np.float32(draw * vf))
    minimax_score = s["minimax_score"]
    if sum_visits == 0:
        minimax_value = _score_to_value(minimax_score)
    else:
        w, d = np.float32(win_rate / np.float32(sum_visits)), np.float32(draw_rate / np.float32(sum_visits))
        total = np.float32(w + d)
        minimax_value = (np.float32(w / total), np.float32(d / total)) if total > np.float32(1.0) else (w, d)
    return {"board": board, "visit_count": visits, "policy_prior": prior, "action_values": values, "action_scores": scores,
            "minimax_value": (float(minimax_value[0]), float(minimax_value[1])), "minimax_score": minimax_score,
            "moves_left": len(game["moves"]) - s["move_number"], "game_outcome": game["outcome"], "played_move": game["moves"][s["move_number"]],
            "flags": s["flags"], "rows": rows, "cols": cols}


def training_targets_visits(sample):
    """SamplerVisits::prepare_training_data: policy target from the visit counts (proven wins / losses overridden), action value targets,
    the value target from the game outcome seen by the side to move."""
    import numpy as np
    cells = sample["rows"] * sample["cols"]
    policy = np.zeros(cells, np.float32)
    value_targets = sample["action_values"].copy()
    visits = sample["visit_count"].copy()
    for i in range(cells):
        sc = int(sample["action_scores"][i])
        if _score_is_proven(sc):
            value_targets[i] = _score_to_value(sc)
            visits[i] = max(1, visits[i])
        pv = (sc >> 13) & 3
        policy[i] = np.float32(1.0e-6) if pv == 0 else (np.float32(1.0e6) if pv == 3 else np.float32(sample["visit_count"][i]))
    policy = _normalize(policy)
    sign = sample["played_move"] & 3
    outcome = sample["game_outcome"]
    value = (0.0, 1.0) if outcome == 1 else ((1.0, 0.0) if (outcome == 2) == (sign == 1) else (0.0, 0.0))
    return {"policy_target": policy, "action_values_target": value_targets, "visit_count": visits, "value_target": value,
            "sign_to_move": sign, "moves_left": float(sample["moves_left"])}


_libm = None


def _expf_logf():
    """glibc's expf / logf: SamplerValues calls std::exp / std::log on floats, and bit-identical targets need the same routines."""
    global _libm
    if _libm is None:
        import ctypes
        import ctypes.util
        lib = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        lib.expf.restype = lib.logf.restype = ctypes.c_float
        lib.expf.argtypes = lib.logf.argtypes = [ctypes.c_float]
        _libm = (lib.expf, lib.logf)
    return _libm


def _score_distance(s):
    """Score::getDistance (search/Score.hpp:101-114)."""
    pv, ev = (s >> 13) & 3, (s & 8191) - 4000
    return ev if pv in (0, 1) else (-ev if pv == 3 else 0)


def training_targets_values(sample):
    """SamplerValues::prepare_training_data (src/dataset/Sampler.cpp:138-216): the policy target is a softmax over 50 * Q + log(prior) of the
    empty cells (proven moves get +-1 / distance terms, unvisited ones the prior-weighted mean), in float32 step by step."""
    import numpy as np
    f32 = np.float32
    expf, logf = _expf_logf()
    cells = sample["rows"] * sample["cols"]
    board, scores, prior = sample["board"], sample["action_scores"], sample["policy_prior"]
    visits = sample["visit_count"].copy()
    values = sample["action_values"]
    action_targets = np.zeros((cells, 2), np.float32)
    policy = np.zeros(cells, np.float32)

    def value_of(v, sc):  # get_value (Sampler.cpp:21-24)
        return _score_to_value(sc) if _score_is_proven(sc) else (v[0], v[1])

    def expectation(v):
        return f32(f32(v[0]) + f32(0.5) * f32(v[1]))

    unvisited, max_n = 0, 0
    sum_pq, sum_p = f32(0), f32(0)
    for i in range(cells):
        if board[i] == 0:
            sc = int(scores[i])
            if sample["visit_count"][i] > 0 or _score_is_proven(sc):
                q = value_of(values[i], sc)
                sum_pq = f32(sum_pq + f32(prior[i] * expectation(q)))
                sum_p = f32(sum_p + prior[i])
                max_n = max(max_n, int(sample["visit_count"][i]))
                visits[i] = max(1, visits[i])
            else:
                unvisited += 1
    minimax_v = expectation(value_of(sample["minimax_value"], sample["minimax_score"]))
    sum_pq = f32(f32(sum_p * sum_pq) + f32(f32(f32(1.0) - sum_p) * minimax_v))
    max_value = f32(np.finfo(np.float32).min)
    eps = f32(np.finfo(np.float32).eps)
    for i in range(cells):
        if board[i] == 0:
            q = sum_pq
            with np.errstate(divide="ignore", invalid="ignore"):
                p = max(f32(0.0), f32(f32(f32(1.0) - sum_p) / f32(unvisited)))
            if visits[i] > 0:
                sc = int(scores[i])
                pv = (sc >> 13) & 3  # Score::getProvenValue
                if pv == 0:
                    q = f32(f32(-1.0) / f32(f32(1.0) + f32(_score_distance(sc))))
                elif pv == 1:
                    q = f32(0.5)
                elif pv == 3:
                    q = f32(f32(1.0) + f32(f32(2.0) / f32(f32(1.0) + f32(_score_distance(sc)))))
                else:
                    q = expectation(values[i])
                action_targets[i] = value_of(values[i], sc)
                p = prior[i]
            policy[i] = f32(f32(f32(50.0) * q) + f32(logf(f32(eps + f32(p)))))
            max_value = max(max_value, policy[i])
    total = f32(0)
    for i in range(cells):
        if board[i] == 0:
            policy[i] = f32(expf(max(f32(-20.0), f32(policy[i] - max_value))))
            total = f32(total + policy[i])
    for i in range(cells):
        if board[i] == 0:
            policy[i] = f32(policy[i] / total)
    sign = sample["played_move"] & 3
    outcome = sample["game_outcome"]
    value = (0.0, 1.0) if outcome == 1 else ((1.0, 0.0) if (outcome == 2) == (sign == 1) else (0.0, 0.0))
    return {"policy_target": policy, "action_values_target": action_targets, "visit_count": visits, "value_target": value,
            "sign_to_move": sign, "moves_left": float(sample["moves_left"])}


def _normalize(x):
    """normalize(matrix<float>&) of the reference (utils/misc.cpp): divide by the sequential float sum."""
    import numpy as np
    total = np.float32(0)
    for v in x:
        total = np.float32(total + v)
    if total == 0:
        return np.full_like(x, np.float32(1.0) / np.float32(len(x)))
    return (x * (np.float32(1.0) / total)).astype(np.float32)
