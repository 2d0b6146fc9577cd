#!/usr/bin/env python
"""Generates tests/golden/*.json|npz from the reference's own unit tests and from oracle/_ref/libagref.so.

Run in the build container only (needs /root/reference and a built oracle/_ref); the outputs are committed so that
the GPU box, which has neither, can still check against the reference's known answers.

Two kinds of vectors:
 * known answers hand-written by the reference's authors, transcribed mechanically from test/game/test_renju.cpp
   (EXPECT_TRUE/FALSE(is_forbidden(...)), getOutcome expectations), test_freestyle/standard/caro.cpp (getOutcome)
   test/networks/test_NNInputFeatures.cpp (feature bits at named cells) and test/search/alpha_beta/test_move_generator.cpp
   (action-list size, flags, members and scores per generator mode);
 * outputs of the reference itself (libagref.so) on every board literal found in those files and in
   test/search/alpha_beta/test_move_generator.cpp + src/utils/selfcheck.cpp, plus seeded random boards.
"""
import ctypes
import json
import os
import re
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from refapi import RefOracle  # noqa: E402

ROW_RE = re.compile(r'"((?:\s[_XO!?|])+)\s*\\n"')
RULES = {"FREESTYLE": 0, "STANDARD": 1, "RENJU": 2, "CARO5": 3, "CARO6": 4}
OUTCOMES = {"UNKNOWN": 0, "DRAW": 1, "CROSS_WIN": 2, "CIRCLE_WIN": 3}


def parse_move(text):
    sign = {"_": 0, "X": 1, "O": 2}[text[0]]
    col = ord(text[1]) - ord("a")
    row = int(text[2:])
    return row, col, sign


def board_rows_to_array(rows):
    size = len(rows)
    b = np.zeros((size, size), np.int8)
    for r, line in enumerate(rows):
        toks = line.split()
        assert len(toks) == size, (len(toks), size)
        for c, t in enumerate(toks):
            b[r, c] = {"_": 0, "!": 0, "?": 0, "X": 1, "O": 2}[t]
    return b


def walk_source(path):
    """Yields ('board', array) / ('line', text) events in file order, grouping consecutive board-row literals."""
    rows = []
    with open(path) as f:
        for line in f:
            if line.strip().startswith("//") and not ROW_RE.search(line):
                continue
            m = ROW_RE.search(line)
            if m and not line.strip().startswith("//"):
                rows.append(m.group(1))
                continue
            if rows:
                if 5 <= len(rows) <= 20 and len(rows[0].split()) == len(rows):
                    yield "board", board_rows_to_array(rows)
                rows = []
            yield "line", line


def transcribe_renju(known, boards):
    board = None
    for kind, val in walk_source(f"{REF}/test/game/test_renju.cpp"):
        if kind == "board":
            board = val.copy()
            boards.append((2, board.copy()))
            continue
        line = val.strip()
        if line.startswith("//"):
            continue
        m = re.match(r'(add_move|undo_move)\(Move\("(\w+)"\)\)', line)
        if m and board is not None:
            r, c, s = parse_move(m.group(2))
            board[r, c] = s if m.group(1) == "add_move" else 0
            boards.append((2, board.copy()))
            continue
        m = re.match(r'EXPECT_(TRUE|FALSE)\(is_forbidden\(Move\("(\w+)"\)\)\)', line)
        if m and board is not None:
            r, c, s = parse_move(m.group(2))
            known.append({"kind": "forbidden", "rules": 2, "board": board.flatten().tolist(), "size": 15, "row": r, "col": c, "sign": s,
                          "expected": m.group(1) == "TRUE", "source": "test/game/test_renju.cpp"})
            continue
        m = re.match(r'EXPECT_EQ\(getOutcome\(GameRules::(\w+), board, Move\("(\w+)"\)\), GameOutcome::(\w+)\)', line)
        if m and board is not None:
            r, c, s = parse_move(m.group(2))
            known.append({"kind": "outcome", "rules": RULES[m.group(1)], "board": board.flatten().tolist(), "size": 15, "row": r, "col": c,
                          "sign": s, "expected": OUTCOMES[m.group(3)], "source": "test/game/test_renju.cpp"})


def transcribe_outcomes(fname, known, boards):
    board = None
    for kind, val in walk_source(f"{REF}/test/game/{fname}"):
        if kind == "board":
            board = val.copy()
            continue
        line = val.strip()
        if line.startswith("//"):
            continue
        m = re.match(r'EXPECT_EQ\(getOutcome\(GameRules::(\w+), board, Move\("(\w+)"\)\), GameOutcome::(\w+)\)', line)
        if m and board is not None:
            r, c, s = parse_move(m.group(2))
            boards.append((RULES[m.group(1)], board.copy()))
            known.append({"kind": "outcome", "rules": RULES[m.group(1)], "board": board.flatten().tolist(), "size": board.shape[0], "row": r,
                          "col": c, "sign": s, "expected": OUTCOMES[m.group(3)], "source": f"test/game/{fname}"})


def transcribe_feature_bits(known, boards):
    """test_NNInputFeatures.cpp: EXPECT_TRUE(is_set_bit<B>(features.at(r, c))) after a board literal; the rule and side to
    move come from the nearest preceding GameConfig(...) / setBoard(board, Sign::...) lines."""
    board, rules, stm = None, 0, 1
    for kind, val in walk_source(f"{REF}/test/networks/test_NNInputFeatures.cpp"):
        if kind == "board":
            board = val.copy()
            continue
        line = val.strip()
        m = re.search(r"GameConfig\w*\s*\w*\(GameRules::(\w+)", line)
        if m:
            rules = RULES[m.group(1)]
        m = re.search(r"sign_to_move = Sign::(\w+)", line)
        if m:
            stm = 1 if m.group(1) == "CROSS" else 2
            if board is not None:
                boards.append((rules, board.copy()))
        m = re.match(r"EXPECT_(TRUE|FALSE)\(is_set_bit<(\d+)>\(features\.at\((\d+), (\d+)\)\)\)", line)
        if m and board is not None:
            known.append({"kind": "feature_bit", "rules": rules, "board": board.flatten().tolist(), "size": board.shape[0], "stm": stm,
                          "bit": int(m.group(2)), "row": int(m.group(3)), "col": int(m.group(4)), "expected": m.group(1) == "TRUE",
                          "source": "test/networks/test_NNInputFeatures.cpp"})


SCORES = {"loss_in": 0, "draw_in": 1, "win_in": 3}


def transcribe_move_generator(known):
    """test/search/alpha_beta/test_move_generator.cpp: per TEST one board, MoveGenWrapper objects (rules, side to move), action lists
    generated in a MoveGeneratorMode, and EXPECT_* on size / must_defend / has_initiative / contains(Move) / getScoreOf(Move)."""
    board, wrappers, lists = None, {}, {}
    for kind, val in walk_source(f"{REF}/test/search/alpha_beta/test_move_generator.cpp"):
        if kind == "board":
            board, wrappers, lists = val.copy(), {}, {}
            continue
        line = val.strip()
        if line.startswith("//") or board is None:
            continue
        m = re.match(r"MoveGenWrapper (\w+)\(GameRules::(\w+), board, Sign::(\w+)\);", line)
        if m:
            wrappers[m.group(1)] = (RULES[m.group(2)], 1 if m.group(3) == "CROSS" else 2)
            continue
        m = re.match(r"(?:const )?ActionList (\w+) = (\w+)\(MoveGeneratorMode::(\w+)\);", line)
        if m and m.group(2) in wrappers:
            rules, stm = wrappers[m.group(2)]
            entry = {"kind": "movegen", "rules": rules, "board": board.flatten().tolist(), "size": board.shape[0], "stm": stm, "mode": m.group(3),
                     "contains": [], "scores": [], "source": "test/search/alpha_beta/test_move_generator.cpp"}
            lists[m.group(1)] = entry
            known.append(entry)
            continue
        m = re.match(r"EXPECT_(EQ|GE)\((\w+)\.size\(\), (\d+)\);", line)
        if m and m.group(2) in lists:
            lists[m.group(2)]["size_eq" if m.group(1) == "EQ" else "size_ge"] = int(m.group(3))
            continue
        m = re.match(r"EXPECT_(TRUE|FALSE)\((\w+)\.(must_defend|has_initiative)\);", line)
        if m and m.group(2) in lists:
            lists[m.group(2)][m.group(3)] = m.group(1) == "TRUE"
            continue
        m = re.match(r'EXPECT_TRUE\((\w+)\.contains\(Move\("(\w+)"\)\)\);', line)
        if m and m.group(1) in lists:
            r, c, _ = parse_move(m.group(2))
            lists[m.group(1)]["contains"].append([r, c])
            continue
        m = re.match(r'EXPECT_EQ\((\w+)\.getScoreOf\(Move\("(\w+)"\)\), Score::(\w+)\((\d+)\)\);', line)
        if m and m.group(1) in lists:
            r, c, _ = parse_move(m.group(2))
            pv, n = SCORES[m.group(3)], int(m.group(4))
            lists[m.group(1)]["scores"].append([r, c, (pv << 13) | (4000 + (-n if pv == 3 else n))])


def collect_boards(path, rules_list, boards):
    for kind, val in walk_source(path):
        if kind == "board":
            for rules in rules_list:
                boards.append((rules, val.copy()))


def random_boards(rng, size, count):
    out = []
    for _ in range(count):
        n = int(rng.integers(0, int(0.6 * size * size) + 1))
        b = np.zeros(size * size, np.int8)
        idx = rng.permutation(size * size)[:n]
        b[idx[0::2]] = 1
        b[idx[1::2]] = 2
        out.append(b.reshape(size, size))
    return out


def main():
    ref = RefOracle()
    known, boards = [], []
    transcribe_renju(known, boards)
    for fname in ("test_freestyle.cpp", "test_standard.cpp", "test_caro.cpp"):
        transcribe_outcomes(fname, known, boards)
    transcribe_feature_bits(known, boards)
    transcribe_move_generator(known)
    collect_boards(f"{REF}/test/search/alpha_beta/test_move_generator.cpp", [0, 1, 2, 3, 4], boards)
    collect_boards(f"{REF}/src/utils/selfcheck.cpp", [0, 2], boards)
    rng = np.random.default_rng(20261017)
    for rules in range(5):
        for size in (15, 20):
            boards += [(rules, b) for b in random_boards(rng, size, 12)]
    # de-duplicate
    seen, uniq = set(), []
    for rules, b in boards:
        key = (rules, b.shape[0], b.tobytes())
        if key not in seen:
            seen.add(key)
            uniq.append((rules, b))
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(known, f)
    # reference outputs for every (rules, board, side to move)
    recs = {"rules": [], "size": [], "stm": [], "board": [], "pattern_types": [], "threats": [], "forbidden": [], "features": [],
            "hist_counts": [], "hist_cells": []}
    for rules, b in uniq:
        size = b.shape[0]
        for stm in (1, 2):
            st = ref.set_board(rules, size, b.flatten(), stm)
            pad = lambda a, shape: np.pad(a, [(0, s - d) for d, s in zip(a.shape, shape)])  # noqa: E731
            recs["rules"].append(rules)
            recs["size"].append(size)
            recs["stm"].append(stm)
            recs["board"].append(pad(b.flatten(), (400,)))
            recs["pattern_types"].append(pad(st["pattern_types"], (400, 4)))
            recs["threats"].append(pad(st["threats"], (400, 2)))
            recs["forbidden"].append(pad(st["forbidden"], (400,)))
            recs["features"].append(pad(st["features"], (400,)))
            recs["hist_counts"].append(st["hist_counts"])
            recs["hist_cells"].append(pad(st["hist_cells"], (2, 10, 400)))
    np.savez_compressed(os.path.join(HERE, "reference_states.npz"), **{k: np.array(v) for k, v in recs.items()})
    print(f"{len(known)} known answers, {len(uniq)} boards, {len(recs['rules'])} reference states")


if __name__ == "__main__":
    main()
