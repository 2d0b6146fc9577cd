"""Games in flight in the reference's own on-disk layout: <working directory>/saved_state/thread_<i>.bin (+ buffer.bin).

GeneratorManager::saveState (src/selfplay/GeneratorManager.cpp:240-263) writes, per GeneratorThread, FileSaver::save(json, binary, 2, compress)
where json is an array with one entry per GameGenerator -- Game::serialize (src/game/Game.cpp:159-167: "game_config", "moves" as Move::text
strings like "Xh7" = sign, column letter, row number) plus "state" (GameGenerator::GameState) and "offset" -- and binary holds each game's
GameDataStorage::serialize (format 201) at that offset (GameGenerator.cpp:122-129). loadState reads it back with Game(json, binary) and
GameDataStorage(binary, offset, 201) (:131-141). These functions convert between that layout and the engine's agb_save_games blob, so a run
can be stopped under one implementation and resumed under the other; what neither format carries (search trees) is rebuilt by both."""
import json
import os
import struct
import zlib

import numpy as np

RULE_NAMES = ["FREESTYLE", "STANDARD", "RENJU", "CARO5", "CARO6"]
GAMEPLAY_SELECT_SOLVE_EVALUATE = 2  # GameGenerator::GameState (GameGenerator.hpp:28-34)


def move_text(short):
    """Move::text (src/game/Move.cpp:141-144): 'X' / 'O', column letter, row number."""
    sign, row, col = short & 3, (short >> 2) & 127, (short >> 9) & 127
    return "_XO|"[sign] + chr(ord("a") + col) + str(row)


def move_from_text(text):
    """Move(const std::string&) (Move.cpp:16-40, 131-136) -> Move::toShort."""
    sign = {"_": 0, "X": 1, "O": 2}[text[0]]
    return sign | (int(text[2:]) << 2) | ((ord(text[1]) - ord("a")) << 9)


def parse_engine_blob(blob):
    """agb_save_games blob (selfplay.cu) -> header dict and a list of games: board, sign_to_move, moves (toShort), samples, record bytes
    ([u32 samples][samples...], the head of a GameDataStorage::serialize)."""
    magic, version, games, rows, cols, rules = struct.unpack_from("<IIiiii", blob, 0)
    if magic != 0x53424741:
        raise ValueError("not an agb_save_games blob")
    cells, off, out = rows * cols, 24, []
    for _ in range(games):
        board = np.frombuffer(blob, np.int8, cells, off).copy()
        off += cells
        stm = blob[off]
        off += 1
        (n_moves,) = struct.unpack_from("<i", blob, off)
        moves = list(struct.unpack_from(f"<{n_moves}H", blob, off + 4))
        off += 4 + 2 * n_moves
        samples, rec_len = struct.unpack_from("<ii", blob, off)
        rec = bytes(blob[off + 8:off + 8 + rec_len])
        off += 8 + rec_len
        out.append({"board": board, "sign_to_move": stm, "moves": moves, "samples": samples, "record": rec})
    return {"version": version, "games": games, "rows": rows, "cols": cols, "rules": rules, "tail": bytes(blob[off:])}, out


def build_engine_blob(rules, rows, cols, games):
    """The reverse: a version-2 blob with fresh random streams and no openings pool (the reference's files carry neither)."""
    out = [struct.pack("<IIiiii", 0x53424741, 2, len(games), rows, cols, rules)]
    for g in games:
        out.append(np.ascontiguousarray(g["board"], np.int8).tobytes())
        out.append(struct.pack("<b", int(g["sign_to_move"])))
        out.append(struct.pack(f"<i{len(g['moves'])}H", len(g["moves"]), *g["moves"]))
        out.append(struct.pack("<ii", g["samples"], len(g["record"])))
        out.append(g["record"])
    n = len(games)
    out.append(b"\0" * (4 * n * 3))  # symmetry, noise and opening counters
    out.append(struct.pack("<i", 0))  # no openings pool
    return b"".join(out)


def write_reference_state(directory, engine_blob, games_per_thread, draw_after=0):
    """Writes <directory>/saved_state/thread_<i>.bin files from an agb_save_games blob; returns their paths."""
    header, games = parse_engine_blob(engine_blob)
    rows, cols = header["rows"], header["cols"]
    cfg = {"rules": RULE_NAMES[header["rules"]], "rows": rows, "cols": cols, "draw_after": draw_after or rows * cols}
    path = os.path.join(directory, "saved_state")
    os.makedirs(path, exist_ok=True)
    written = []
    for t in range((len(games) + games_per_thread - 1) // games_per_thread):
        entries, binary = [], b""
        for g in games[t * games_per_thread:(t + 1) * games_per_thread]:
            entries.append({"game_config": cfg, "moves": [move_text(m) for m in g["moves"]], "state": GAMEPLAY_SELECT_SOLVE_EVALUATE, "offset": len(binary)})
            # GameDataStorage::serialize of a game in flight: samples so far, no moves yet (they are added when the game ends), outcome UNKNOWN,
            # rows = cols = 0 (GameGenerator default-constructs its storage, GameGenerator.hpp:39)
            binary += g["record"] + struct.pack("<IIII", 0, 0, 0, 0)
        data = zlib.compress(json.dumps(entries, indent=2).encode() + b"\n" + binary)
        name = os.path.join(path, f"thread_{t}.bin")
        with open(name, "wb") as f:
            f.write(data)
        written.append(name)
    return written


def read_reference_state(directory, rules, rows, cols):
    """<directory>/saved_state/thread_*.bin -> agb_save_games blob for agb_load_games (games in file order, thread by thread)."""
    from .netfile import find_split_point
    path = os.path.join(directory, "saved_state")
    games, t = [], 0
    while os.path.exists(os.path.join(path, f"thread_{t}.bin")):
        with open(os.path.join(path, f"thread_{t}.bin"), "rb") as f:
            data = zlib.decompress(f.read())
        split = min(len(data), find_split_point(data))
        entries, binary = json.loads(data[:split].decode()), data[split:]
        for e in entries:
            if e["game_config"]["rows"] != rows or e["game_config"]["cols"] != cols or RULE_NAMES.index(e["game_config"]["rules"]) != rules:
                raise ValueError("saved state was written for another game configuration")
            moves = [move_from_text(m) for m in e["moves"]]
            board = np.zeros(rows * cols, np.int8)
            for m in moves:
                board[((m >> 2) & 127) * cols + ((m >> 9) & 127)] = m & 3
            stm = 1 if not moves else 3 - (moves[-1] & 3)
            off = e["offset"]  # one GameDataStorage::serialize: [u32 samples][samples...][u32 moves][moves][outcome][rows][cols]
            (n_samples,) = struct.unpack_from("<I", binary, off)
            end = off + 4
            for _ in range(n_samples):
                (n_entries,) = struct.unpack_from("<I", binary, end + 12)
                end += 16 + 6 * n_entries
            record = binary[off:end]  # the moves of a game in flight live in "moves" above, the storage's own list is still empty
            games.append({"board": board, "sign_to_move": stm, "moves": moves, "samples": n_samples, "record": record})
        t += 1
    if not games:
        raise FileNotFoundError(f"no saved_state/thread_*.bin under {directory}")
    return build_engine_blob(rules, rows, cols, games)
