"""CPU tests (-m "not gpu"): pin the oracle against the reference's own known answers, check the host-compiled kernel
logic against the oracle, and check that the C-ABI library loads, exports every declared symbol and refuses to run
without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, random_boards
from refapi import _p

P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731


def test_reference_known_answers_pin_the_oracle(ref, golden):
    """Every hand-written expectation transcribed from the reference's unit tests holds for oracle/_ref."""
    known, _ = golden
    assert len(known) >= 150
    for k in known:
        size, board = k["size"], np.array(k["board"], np.int8)
        if k["kind"] == "forbidden":
            assert ref.is_forbidden(size, board, k["row"], k["col"], k["sign"]) == k["expected"], k
            # incremental path (PatternCalculator::isForbidden) must agree, as test_renju.cpp:45-51 demands
            st = ref.set_board(k["rules"], size, board, 1)
            if board[k["row"] * size + k["col"]] == 0:
                assert bool(st["forbidden"][k["row"] * size + k["col"]]) == k["expected"], k
        elif k["kind"] == "outcome":
            assert ref.outcome(k["rules"], size, board, k["row"], k["col"], k["sign"], -1) == k["expected"], k
        elif k["kind"] == "feature_bit":
            st = ref.set_board(k["rules"], size, board, k["stm"])
            bit = (int(st["features"][k["row"] * size + k["col"]]) >> k["bit"]) & 1
            assert bool(bit) == k["expected"], k


def test_golden_states_match_the_reference_build(ref, golden):
    """The committed reference_states.npz is what oracle/_ref produces today (guards against a stale fixture)."""
    _, states = golden
    for i in range(0, len(states["rules"]), 7):
        rules, size, stm = int(states["rules"][i]), int(states["size"][i]), int(states["stm"][i])
        c = size * size
        st = ref.set_board(rules, size, states["board"][i][:c], stm)
        assert (st["features"] == states["features"][i][:c]).all()
        assert (st["pattern_types"] == states["pattern_types"][i][:c]).all()
        assert (st["hist_counts"] == states["hist_counts"][i]).all()


@pytest.mark.parametrize("rules", [0, 1, 2, 3, 4])
def test_table_logic_matches_reference_tables(ref, hostsim, rules):
    """tables_logic.cuh (the code the device kernel runs) reproduces PatternTable / ThreatTable bit for bit."""
    pt, ho, th, _ = ref.tables(rules)
    mine = np.zeros(1 << 20, np.uint8)
    hostsim.hostsim_pattern_table(rules, P(mine))
    expected = (pt & 7) | ((ho & 1) << 3) | (((pt >> 4) & 7) << 4) | (((ho >> 1) & 1) << 7)
    assert (mine == expected).all()
    mt = np.zeros(4096, np.uint8)
    hostsim.hostsim_threat_table(rules, P(mt))
    assert (mt == (th[:, 0] | (th[:, 1] << 4))).all()


def _hostsim_tables(hostsim, rules):
    pt = np.zeros(1 << 20, np.uint8)
    tt = np.zeros(4096, np.uint8)
    hostsim.hostsim_pattern_table(rules, P(pt))
    hostsim.hostsim_threat_table(rules, P(tt))
    return pt, tt


def _hostsim_set_board(hostsim, rules, size, board, stm, pt, tt):
    c = size * size
    mp, mt = np.zeros((c, 4), np.uint8), np.zeros((c, 2), np.uint8)
    mf, mforb = np.zeros(c, np.uint32), np.zeros(c, np.uint8)
    ov = ctypes.c_int(0)
    board = np.ascontiguousarray(board, np.int8)
    hostsim.hostsim_set_board(rules, size, P(board), stm, P(pt), P(tt), P(mp), P(mt), P(mf), P(mforb), ctypes.byref(ov))
    assert ov.value == 0
    return mp, mt, mf, mforb


def test_kernel_logic_on_golden_states(hostsim, golden):
    """patterns_logic.cuh against the committed reference outputs (no reference needed: runs on any box)."""
    _, states = golden
    tables = {}
    for i in range(len(states["rules"])):
        rules, size, stm = int(states["rules"][i]), int(states["size"][i]), int(states["stm"][i])
        if rules not in tables:
            tables[rules] = _hostsim_tables(hostsim, rules)
        c = size * size
        mp, mt, mf, mforb = _hostsim_set_board(hostsim, rules, size, states["board"][i][:c], stm, *tables[rules])
        assert (mp == states["pattern_types"][i][:c]).all(), i
        assert (mt == states["threats"][i][:c]).all(), i
        assert (mf == states["features"][i][:c]).all(), i
        assert (mforb == states["forbidden"][i][:c]).all(), i


def test_kernel_logic_on_known_answers(hostsim, golden):
    known, _ = golden
    tables = {}
    for k in known:
        rules, size, board = k["rules"], k["size"], np.array(k["board"], np.int8)
        if rules not in tables:
            tables[rules] = _hostsim_tables(hostsim, rules)
        pt, tt = tables[rules]
        if k["kind"] == "outcome":
            got = hostsim.hostsim_outcome(rules, size, P(board), k["row"], k["col"], k["sign"], 0, P(pt), P(tt))
            assert got == k["expected"], k
        elif k["kind"] == "forbidden" and board[k["row"] * size + k["col"]] == 0:
            _, _, _, mforb = _hostsim_set_board(hostsim, rules, size, board, 1, pt, tt)
            assert bool(mforb[k["row"] * size + k["col"]]) == k["expected"], k
        elif k["kind"] == "feature_bit":
            _, _, mf, _ = _hostsim_set_board(hostsim, rules, size, board, k["stm"], pt, tt)
            assert bool((int(mf[k["row"] * size + k["col"]]) >> k["bit"]) & 1) == k["expected"], k


@pytest.mark.parametrize("rules,size", [(0, 15), (2, 15), (4, 20)])
def test_kernel_logic_random_boards_vs_reference(ref, hostsim, rules, size):
    rng = np.random.default_rng(rules * 100 + size)
    pt, tt = _hostsim_tables(hostsim, rules)
    for board in random_boards(rng, size, 60):
        stm = int(rng.integers(1, 3))
        st = ref.set_board(rules, size, board, stm)
        mp, mt, mf, mforb = _hostsim_set_board(hostsim, rules, size, board, stm, pt, tt)
        assert (mp == st["pattern_types"]).all() and (mt == st["threats"]).all()
        assert (mf == st["features"]).all() and (mforb == st["forbidden"]).all()
        mode = int(rng.integers(0, 8))
        aug = np.zeros_like(mf)
        hostsim.hostsim_augment(P(aug), P(mf), size, mode)
        assert (aug == ref.augment(st["features"], size, mode)).all()


def test_open_three_promotions_match_reference(ref, hostsim):
    """All 11-cell windows with an empty centre that contain one of the open-three shapes."""
    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(200000):
        w = int(rng.integers(0, 1 << 22)) & ~(3 << 10)
        mine = hostsim.hostsim_open3_promotions(w)
        if mine:
            assert mine == ref.lib.agref_open3_promotion_moves(w)
            checked += 1
    assert checked > 100


def test_ctypes_structures_match_the_header(tmp_path):
    """The ctypes mirror of AgbConfig / AgbStats (alphagomoku_b200/_lib.py) has the size and the field offsets the C compiler gives the header."""
    import ctypes
    import subprocess
    from alphagomoku_b200 import _lib
    probes = [("AgbConfig", _lib.AgbConfig), ("AgbStats", _lib.AgbStats)]
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "agb200.h"', "int main(void) {"]
    for cname, cls in probes:
        src.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for field, _ in cls._fields_:
            src.append(f'printf("{cname}.{field} %zu\\n", offsetof({cname}, {field}));')
    src += ["return 0; }"]
    c_file = tmp_path / "abi.c"
    c_file.write_text("\n".join(src))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c_file), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in probes:
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for field, _ in cls._fields_:
            assert int(out[f"{cname}.{field}"]) == getattr(cls, field).offset, (cname, field)


def test_c_abi_exports_every_declared_symbol():
    from alphagomoku_b200 import _lib
    header = open(os.path.join(ROOT, "include", "agb200.h")).read()
    declared = set(re.findall(r"\b(agb_\w+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert b"sm_100a" in lib.agb_version()


def test_engine_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import alphagomoku_b200 as agb
    with pytest.raises(agb.AgbError) as err:
        agb.Engine(agb.GameConfig(), max_boards=4)
    assert err.value.code == -2


def test_record_codecs_match_reference(ref, hostsim):
    """K8 logic (records.cuh, compiled for the host) against SearchDataStorage_v201::loadFrom/serialize of the reference:
    quantised bytes identical on random root statistics, incl. proven scores and empty / dense visit patterns."""
    rng = np.random.default_rng(21)
    hostsim.hostsim_serialize_sample_v201.restype = ctypes.c_size_t
    ref.lib.agref_serialize_sample_v201.restype = ctypes.c_size_t
    for size in (15, 20):
        cells = size * size
        for trial in range(200):
            board = random_boards(rng, size, 1)[0]
            empty = board == 0
            visits = np.where(empty & (rng.random(cells) < rng.random()), rng.integers(0, 1 + int(rng.integers(1, 800)), cells), 0).astype(np.int32)
            prior = np.where(empty, rng.random(cells) ** 3, 0).astype(np.float32)
            prior /= max(prior.sum(), 1e-9) if trial % 7 else 1.0
            win = np.where(empty, rng.random(cells), 0).astype(np.float32)
            draw = np.where(empty, (1 - win) * rng.random(cells), 0).astype(np.float32)
            scores = np.full(cells, (2 << 13) | 4000, np.uint16)
            k = rng.random(cells)
            scores[empty & (k < 0.05)] = (3 << 13) | (4000 - int(rng.integers(1, 70)))  # win in n
            scores[empty & (k > 0.95)] = (0 << 13) | (4000 + int(rng.integers(1, 70)))  # loss in n
            scores[empty & (k > 0.45) & (k < 0.5)] = (2 << 13) | (4000 + int(rng.integers(-1000, 1001)))  # unproven eval
            if trial % 11 == 0:
                visits[:] = 0
            minimax = int(scores[rng.integers(cells)])
            flags = int(rng.integers(0, 8))
            mine = np.zeros(16 + 6 * cells, np.uint8)
            theirs = np.zeros(16 + 6 * cells, np.uint8)
            n1 = hostsim.hostsim_serialize_sample_v201(cells, P(board), P(visits), P(prior), P(win), P(draw), P(scores), minimax, flags, P(mine))
            n2 = ref.lib.agref_serialize_sample_v201(size, size, P(board), P(visits), P(prior), P(win), P(draw), P(scores), minimax, flags, P(theirs),
                                                     ctypes.c_size_t(theirs.size))
            assert n1 == n2, (size, trial, n1, n2)
            assert (mine[:n1] == theirs[:n2]).all(), (size, trial, np.nonzero(mine[:n1] != theirs[:n2])[0][:8])


def test_game_data_buffer_file_loads_in_the_reference(ref, tmp_path):
    """A format-201 file framed by alphagomoku_b200.dataset loads with the reference's GameDataBuffer::load."""
    from alphagomoku_b200 import dataset
    import struct
    games = []
    for g in range(3):
        samples = b""
        for s in range(g + 1):
            samples += struct.pack("<HHHHHHI", 1, 2, 3, (2 << 13) | 4000, s, 0, 2) + bytes([0, 10, 20, 128, 30, 40]) + bytes([5, 1, 2, 128, 3, 4])
        moves = [1 | (7 << 2) | (7 << 9), 2 | (7 << 2) | (8 << 9)]
        games.append(struct.pack("<I", g + 1) + samples + struct.pack("<I", len(moves)) + struct.pack(f"<{len(moves)}H", *moves) + struct.pack("<iii", 2, 0, 0))
    buf = dataset.GameDataBuffer(1, 15, 15)
    buf.add_records(b"".join(games), 3)
    path = str(tmp_path / "buffer.bin")
    buf.save(path)
    spg, mpg, out = np.zeros(8, np.int32), np.zeros(8, np.int32), np.zeros(8, np.int32)
    rows, cols, rules = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int32(0)
    n = ref.lib.agref_buffer_load(path.encode(), P(spg), P(mpg), P(out), 8, ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(rules))
    assert n == 3 and (rows.value, cols.value, rules.value) == (15, 15, 1)
    assert spg[:3].tolist() == [1, 2, 3] and mpg[:3].tolist() == [2, 2, 2] and out[:3].tolist() == [2, 2, 2]
    assert dataset.parse_record(games[2])["samples"][1]["move_number"] == 1


# ---- K5: static solver (MoveGenerator in OPTIMAL mode + static evaluation) ------------------------------------------------------
def _solver_tables(hostsim, rules):
    pattern = np.zeros(1 << 20, np.uint8)
    threat = np.zeros(4096, np.uint8)
    deftab = np.zeros(15 * 256 * 2, np.uint16)
    hostsim.hostsim_pattern_table(rules, _p(pattern))
    hostsim.hostsim_threat_table(rules, _p(threat))
    hostsim.hostsim_defensive_table(rules, _p(deftab))
    return pattern, threat, deftab


def _solve_both(ref, hostsim, tables, rules, size, board, stm, draw_after=0):
    cells = size * size
    out = []
    for which in range(2):
        moves, scores = np.zeros(cells, np.uint16), np.zeros(cells, np.uint16)
        result, flags = np.zeros(1, np.uint16), np.zeros(1, np.int32)
        if which == 0:
            n = ref.lib.agref_solve(rules, size, size, draw_after, _p(board), int(stm), 1, _p(moves), _p(scores), _p(result), _p(flags))
        else:
            n = hostsim.hostsim_solve_static(rules, size, draw_after, _p(board), int(stm), _p(tables[0]), _p(tables[1]), _p(tables[2]), _p(moves),
                                             _p(scores), _p(result), _p(flags))
        out.append((moves[:n].copy(), scores[:n].copy(), int(result[0]), int(flags[0]) & 1))
    return out


@pytest.mark.parametrize("rules", [0, 1, 2, 3, 4])
def test_defensive_move_table_matches_reference(ref, hostsim, rules):
    """DefensiveMoveTable::getMoves (DefensiveMoveTable.cpp:393-477) against solver_logic.cuh's table + lookup on random 13-cell
    windows whose centre is empty, for every threat type the table serves."""
    rng = np.random.default_rng(900 + rules)
    _, _, deftab = _solver_tables(hostsim, rules)
    hostsim.hostsim_defensive_mask.restype = ctypes.c_uint32
    ref.lib.agref_defensive_moves.restype = ctypes.c_uint16
    checked = 0
    for _ in range(6000):
        cellsv = rng.choice(4, size=13, p=[0.45, 0.25, 0.25, 0.05])
        cellsv[6] = 0
        window = int(sum(int(v) << (2 * i) for i, v in enumerate(cellsv)))
        for defender in (1, 2):
            for threat in (3, 4, 5, 6, 7):  # OPEN_3, HALF_OPEN_4, OPEN_4, DOUBLE_4, FIVE
                a = ref.lib.agref_defensive_moves(rules, window, defender, threat)
                b = hostsim.hostsim_defensive_mask(_p(deftab), rules, window, defender, threat)
                assert a == b, (rules, hex(window), defender, threat, bin(a), bin(b))
                checked += 1
    assert checked == 60000


@pytest.mark.parametrize("rules,size", [(0, 15), (1, 15), (3, 20), (4, 15), (0, 12)])
def test_static_solver_matches_reference(ref, hostsim, rules, size):
    """AlphaBetaSearch::solve with a node limit of 1 (AlphaBetaSearch.cpp:77-156 -> MoveGenerator.cpp, static evaluation) against the
    host-compiled K5 logic: same action list in the same order, same action scores, same position score and must-defend flag."""
    rng = np.random.default_rng(1000 + 10 * rules + size)
    tables = _solver_tables(hostsim, rules)
    boards = random_boards(rng, size, 300, max_fill=0.5)
    for i, board in enumerate(boards):
        stm = 1 if (np.count_nonzero(board) % 2 == 0) else 2
        if i % 5 == 0:
            stm = 3 - stm
        a, b = _solve_both(ref, hostsim, tables, rules, size, board, stm)
        assert np.array_equal(a[0], b[0]), (i, a[0], b[0])
        assert np.array_equal(a[1], b[1]), i
        assert a[2:] == b[2:], (i, a[2:], b[2:])


def test_static_solver_renju_matches_reference_up_to_order(ref_fast, hostsim):
    """RENJU: the reference's isForbidden() re-runs add/undo on its calculator and so reorders its threat lists mid-generation; the device
    lists keep K1's order. Scores, flags and the SET of (move, score) must still agree. Uses the reference's Release build: on random
    (unreachable) positions its debug build aborts in MoveGenerator.cpp:999."""
    ref = ref_fast
    rng = np.random.default_rng(1234)
    tables = _solver_tables(hostsim, 2)
    boards = random_boards(rng, 15, 300, max_fill=0.5)
    for i, board in enumerate(boards):
        stm = 1 if (np.count_nonzero(board) % 2 == 0) else 2
        a, b = _solve_both(ref, hostsim, tables, 2, 15, board, stm)
        assert a[2:] == b[2:], (i, a[2:], b[2:])
        assert sorted(zip(a[0].tolist(), a[1].tolist())) == sorted(zip(b[0].tolist(), b[1].tolist())), i


@pytest.mark.parametrize("rules,size,max_nodes,use_fast,max_fill", [
    (0, 15, 100, False, 0.45), (1, 15, 100, False, 0.45), (3, 20, 60, False, 0.45), (4, 15, 400, False, 0.45), (2, 15, 100, True, 0.45),
    (2, 15, 1, True, 0.45),
    # sparse boards: calm positions, where quiet root children are visited without add/undo (solver_search.cuh: quiet_child_visit)
    (0, 15, 100, False, 0.08), (1, 15, 100, False, 0.12), (3, 20, 100, False, 0.06), (2, 15, 100, True, 0.08)])
def test_alpha_beta_solver_matches_reference(ref, ref_fast, hostsim, rules, size, max_nodes, use_fast, max_fill):
    """AlphaBetaSearch::solve with its transposition table kept between positions and generations (AlphaBetaSearch.cpp:77-339,
    SharedHashTable.hpp:27-220) against the host-compiled K5 search (solver_search.cuh): the same hash keys and table size, then the same
    action order, action scores, position score, flags and node count for every position of a sequence. RENJU also pins the ORDER of the
    threat lists after forbidden-move checks; it runs on the reference's Release build (its debug build asserts on unreachable boards)."""
    oracle = ref_fast if use_fast else ref
    lib = oracle.lib
    lib.agref_solver_create.restype = ctypes.c_void_p
    hostsim.hostsim_solver_create.restype = ctypes.c_void_p
    tables = _solver_tables(hostsim, rules)
    cells = size * size
    rh = ctypes.c_void_p(lib.agref_solver_create(rules, size, size, 0))
    keys = np.zeros(4 * cells, np.uint64)
    lib.agref_solver_keys(rh, _p(keys))
    entries = 4 * 1024 * 1024  # AlphaBetaSearch.cpp:55
    hh = ctypes.c_void_p(hostsim.hostsim_solver_create(rules, size, 0, _p(tables[0]), _p(tables[1]), _p(tables[2]), _p(keys), ctypes.c_size_t(entries)))
    rng = np.random.default_rng(500 + 7 * rules + max_nodes)
    boards = random_boards(rng, size, 150, max_fill=max_fill)
    total_nodes = 0
    for i, board in enumerate(boards):
        stm = 1 if (np.count_nonzero(board) % 2 == 0) else 2
        res = []
        for which in range(2):
            moves, scores = np.zeros(cells, np.uint16), np.zeros(cells, np.uint16)
            result, flags = np.zeros(1, np.uint16), np.zeros(1, np.int32)
            if which == 0:
                n = lib.agref_solver_solve(rh, _p(board), stm, max_nodes, _p(moves), _p(scores), _p(result), _p(flags))
            else:
                n = hostsim.hostsim_solver_solve(hh, _p(board), stm, max_nodes, _p(moves), _p(scores), _p(result), _p(flags))
            res.append((moves[:n].copy(), scores[:n].copy(), int(result[0]), int(flags[0])))
        a, b = res
        assert np.array_equal(a[0], b[0]), (i, a[0], b[0])
        assert np.array_equal(a[1], b[1]), i
        assert a[2:] == b[2:], (i, hex(a[2]), hex(b[2]), hex(a[3]), hex(b[3]))
        total_nodes += a[3] >> 8
        if i % 7 == 6:  # a new search (Search::setBoard) starts a new table generation
            lib.agref_solver_next_generation(rh)
            hostsim.hostsim_solver_next_generation(hh)
    assert total_nodes > 150 or max_nodes == 1
    lib.agref_solver_destroy(rh)
    hostsim.hostsim_solver_destroy(hh)


# ---- the reference's own move generator tests (test/search/alpha_beta/test_move_generator.cpp) --------------------------------------
MODES = {"BASIC": 0, "THREATS": 1, "OPTIMAL": 2, "REDUCED": 3, "LEGAL": 4}


def _check_movegen_expectations(entry, moves, scores, flags):
    cells = {((int(m) >> 2) & 127, (int(m) >> 9) & 127): int(s) for m, s in zip(moves, scores)}
    if "size_eq" in entry:
        assert len(moves) == entry["size_eq"], (entry["size_eq"], len(moves))
    if "size_ge" in entry:
        assert len(moves) >= entry["size_ge"]
    if "must_defend" in entry:
        assert bool(flags & 1) == entry["must_defend"]
    if "has_initiative" in entry:
        assert bool(flags & 2) == entry["has_initiative"]
    for r, c in entry["contains"]:
        assert (r, c) in cells, (r, c, sorted(cells))
    for r, c, s in entry["scores"]:
        assert cells.get((r, c)) == s, (r, c, hex(s), cells.get((r, c)))


def _movegen_entries(golden, modes):
    return [e for e in golden[0] if e["kind"] == "movegen" and e["mode"] in modes]


def test_reference_move_generator_known_answers_pin_the_oracle(ref, golden):
    """The hand-written expectations of the reference's move generator tests (list size, must_defend, has_initiative, members, scores in
    every MoveGeneratorMode) hold for oracle/_ref: the build used as the K5 oracle behaves like the reference's authors say it must."""
    entries = _movegen_entries(golden, MODES)
    assert len(entries) >= 50
    for e in entries:
        cells = e["size"] ** 2
        moves, scores, flags = np.zeros(cells, np.uint16), np.zeros(cells, np.uint16), np.zeros(1, np.int32)
        board = np.array(e["board"], np.int8)
        n = ref.lib.agref_generate(e["rules"], e["size"], e["size"], _p(board), e["stm"], MODES[e["mode"]], _p(moves), _p(scores), _p(flags))
        _check_movegen_expectations(e, moves[:n], scores[:n], int(flags[0]))


def test_kernel_logic_move_generator_known_answers(ref, hostsim, golden):
    """The same expectations for the host-compiled K5 generator (THREATS and OPTIMAL, the modes the solver uses), plus equality with the
    reference's list, order included."""
    entries = _movegen_entries(golden, ("THREATS", "OPTIMAL"))
    assert len(entries) >= 45
    tables = {}
    for e in entries:
        if e["rules"] not in tables:
            tables[e["rules"]] = _solver_tables(hostsim, e["rules"])
        t = tables[e["rules"]]
        cells = e["size"] ** 2
        board = np.array(e["board"], np.int8)
        moves, scores, flags = np.zeros(cells, np.uint16), np.zeros(cells, np.uint16), np.zeros(1, np.int32)
        n = hostsim.hostsim_generate(e["rules"], e["size"], _p(board), e["stm"], MODES[e["mode"]], _p(t[0]), _p(t[1]), _p(t[2]), _p(moves), _p(scores),
                                     _p(flags))
        _check_movegen_expectations(e, moves[:n], scores[:n], int(flags[0]))
        rm, rs, rf = np.zeros(cells, np.uint16), np.zeros(cells, np.uint16), np.zeros(1, np.int32)
        rn = ref.lib.agref_generate(e["rules"], e["size"], e["size"], _p(board), e["stm"], MODES[e["mode"]], _p(rm), _p(rs), _p(rf))
        assert rn == n and (rm[:n] == moves[:n]).all() and (rs[:n] == scores[:n]).all() and rf[0] == flags[0]


def test_score_arithmetic_known_answers(hostsim):
    """test/search/test_Score.cpp:14-111 restated on the 16-bit wire form the device works with (3-bit proven value, 13-bit eval + 4000,
    Score.hpp:47-68): ordering is the ordering of the raw words, unary minus, invert_up / invert_down, the two infinities."""
    win_in = lambda n: (3 << 13) | (4000 - n)  # noqa: E731
    loss_in = lambda n: (0 << 13) | (4000 + n)  # noqa: E731
    draw_in = lambda n: (1 << 13) | (4000 + n)  # noqa: E731
    ev = lambda e: (2 << 13) | (4000 + e)  # noqa: E731
    minus_inf, plus_inf = 0x0000, 0xFFFF
    for f in (hostsim.hostsim_score_negate, hostsim.hostsim_score_invert):
        f.restype = ctypes.c_uint16
    neg = lambda s: hostsim.hostsim_score_negate(ctypes.c_uint16(s))  # noqa: E731
    inv = lambda s, d: hostsim.hostsim_score_invert(ctypes.c_uint16(s), d)  # noqa: E731
    # comparison (test_Score.cpp:16-28)
    assert win_in(4) > win_in(5) and win_in(4) > draw_in(5) and win_in(4) > ev(4000) and win_in(4) > loss_in(5)
    assert ev(123) > ev(-123) and ev(-4000) > draw_in(5) and ev(-4000) > loss_in(5)
    assert draw_in(4) > draw_in(3) and draw_in(4) > loss_in(5) and loss_in(4) > loss_in(3)
    # unary minus (:58-72), invert_up (:77-91), invert_down (:96-110) on s1 = win_in(5), s2 = loss_in(5), s3 = draw_in(5), s4 = 123, s5 = -123
    s = [win_in(5), loss_in(5), draw_in(5), ev(123), ev(-123), minus_inf, plus_inf]
    assert [neg(x) for x in s] == [loss_in(5), win_in(5), draw_in(5), ev(-123), ev(123), plus_inf, minus_inf]
    assert [inv(x, +1) for x in s] == [loss_in(6), win_in(6), draw_in(6), ev(-123), ev(123), plus_inf, minus_inf]
    assert [inv(x, -1) for x in s] == [loss_in(4), win_in(4), draw_in(4), ev(-123), ev(123), plus_inf, minus_inf]
    # infinities are not proven scores (:36-43)
    assert hostsim.hostsim_score_is_proven(ctypes.c_uint16(minus_inf)) == 0 and hostsim.hostsim_score_is_proven(ctypes.c_uint16(plus_inf)) == 0
    assert hostsim.hostsim_score_is_proven(ctypes.c_uint16(win_in(3))) == 1 and hostsim.hostsim_score_is_proven(ctypes.c_uint16(ev(5))) == 0


def test_reference_side_shims_compile_against_the_reference_headers(tmp_path):
    """integration/agb200_shims.hpp (the NNEvaluator drop-in and the batched AlphaBetaSearch::solve of INTEGRATION.md) must compile against the
    reference's own headers: SearchTask accessors, Score / Value / Move constructors and the C ABI agree."""
    if not os.path.isdir("/root/reference/include"):
        pytest.skip("/root/reference absent")
    import subprocess
    tu = tmp_path / "shim_tu.cpp"
    tu.write_text('#include "integration/agb200_shims.hpp"\nint main() { return 0; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I/root/reference/include", "-I" + os.path.join(ROOT, "oracle", "ref_stub"),
                           "-I" + os.path.join(ROOT, "include"), "-I" + ROOT, str(tu)])


def test_trainer_side_reader_matches_the_reference(ref, tmp_path):
    """Records (format 201) -> GameDataBuffer file -> what the reference's trainer reads: GameDataStorage::getSample and
    SamplerVisits::prepare_training_data through oracle/_ref, against alphagomoku_b200.dataset.decode_sample / training_targets_visits on
    the same bytes. The games are played by the reference's own search with a synthetic evaluator, so every record field is exercised."""
    import refapi
    from alphagomoku_b200 import dataset
    size, cells = 15, 225
    rng = np.random.default_rng(77)

    def evaluate(features):
        n = features.shape[0]
        policy = rng.random((n, cells)).astype(np.float32) ** 4
        policy /= policy.sum(1, keepdims=True)
        win = rng.random(n).astype(np.float32)
        draw = ((1 - win) * rng.random(n)).astype(np.float32)
        return policy, np.stack([win, draw, 1 - win - draw], 1), None

    records = []
    for g in range(3):
        sp = refapi.RefSelfplay(1, size, evaluate, max_batch_size=4, max_simulations=60, use_solver=True, solver_max_positions=50, draw_after=18)
        board = np.zeros(cells, np.int8)
        board[[7 * 15 + 7, 7 * 15 + 8 + g, 8 * 15 + 7]] = [1, 2, 1]
        sp.set_position(board, 2)
        for _ in range(4000):
            if sp.step() == 2:
                break
        records.append(sp.record())
        sp.close()
    buf = dataset.GameDataBuffer(1, size, size, 18)
    buf.add_records(b"".join(records), len(records))
    path = str(tmp_path / "buffer.bin")
    buf.save(path)
    lib = ref.lib
    checked = 0
    for g, rec in enumerate(records):
        game = dataset.parse_record(rec)
        assert len(game["samples"]) >= 3
        for k in range(len(game["samples"])):
            board, visits, prior = np.zeros(cells, np.int8), np.zeros(cells, np.int32), np.zeros(cells, np.float32)
            values, scores, scalars = np.zeros((cells, 2), np.float32), np.zeros(cells, np.uint16), np.zeros(8, np.float32)
            policy_t, value_t, visits_t, scalars_t = np.zeros(cells, np.float32), np.zeros((cells, 2), np.float32), np.zeros(cells, np.float32), np.zeros(6, np.float32)
            rc = lib.agref_buffer_sample(path.encode(), g, k, _p(board), _p(visits), _p(prior), _p(values), _p(scores), _p(scalars), _p(policy_t), _p(value_t),
                                         _p(visits_t), _p(scalars_t))
            assert rc == 0
            mine = dataset.decode_sample(game, k, size, size)
            assert (mine["board"] == board).all() and (mine["visit_count"] == visits).all() and (mine["action_scores"] == scores).all()
            assert (mine["policy_prior"].view(np.uint32) == prior.view(np.uint32)).all()
            assert (mine["action_values"].view(np.uint32) == values.view(np.uint32)).all()
            assert np.float32(mine["minimax_value"][0]) == scalars[0] and np.float32(mine["minimax_value"][1]) == scalars[1]
            assert (mine["minimax_score"], mine["moves_left"], mine["game_outcome"], mine["played_move"], mine["flags"]) == tuple(int(x) for x in scalars[2:7])
            targets = dataset.training_targets_visits(mine)
            assert (targets["policy_target"].view(np.uint32) == policy_t.view(np.uint32)).all()
            assert (targets["action_values_target"].view(np.uint32) == value_t.view(np.uint32)).all()
            assert (targets["visit_count"] == visits_t.astype(np.int32)).all()
            assert np.float32(targets["value_target"][0]) == scalars_t[0] and np.float32(targets["value_target"][1]) == scalars_t[1]
            assert targets["moves_left"] == scalars_t[4] and targets["sign_to_move"] == int(scalars_t[5])
            # SamplerValues (the default sampler_type of the reference's TrainingConfig, configs.hpp:170)
            rc = lib.agref_buffer_sample_with(path.encode(), g, k, _p(board), _p(visits), _p(prior), _p(values), _p(scores), _p(scalars), _p(policy_t), _p(value_t),
                                              _p(visits_t), _p(scalars_t), 1)
            assert rc == 0
            targets = dataset.training_targets_values(mine)
            assert (targets["policy_target"].view(np.uint32) == policy_t.view(np.uint32)).all(), np.abs(targets["policy_target"] - policy_t).max()
            assert (targets["action_values_target"].view(np.uint32) == value_t.view(np.uint32)).all()
            assert (targets["visit_count"] == visits_t.astype(np.int32)).all()
            assert np.float32(targets["value_target"][0]) == scalars_t[0] and np.float32(targets["value_target"][1]) == scalars_t[1]
            checked += 1
    assert checked >= 10
