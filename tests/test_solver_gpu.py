"""K5 as a service (agb_solve) and the opening generator on the GPU, against the reference's AlphaBetaSearch (oracle/_ref)."""
import ctypes

import numpy as np
import pytest

from conftest import random_boards
from refapi import _p

pytestmark = pytest.mark.gpu


def _ref_solver(lib, rules, size):
    lib.agref_solver_create.restype = ctypes.c_void_p
    h = ctypes.c_void_p(lib.agref_solver_create(rules, size, size, 0))
    keys = np.zeros((2 * size * size, 2), np.uint64)
    lib.agref_solver_keys(h, _p(keys))
    return h, keys


def _ref_solve(lib, h, board, stm, max_nodes, cells):
    moves, scores = np.zeros(cells, np.uint16), np.zeros(cells, np.uint16)
    result, flags = np.zeros(1, np.uint16), np.zeros(1, np.int32)
    lib.agref_solver_clear(h)
    n = lib.agref_solver_solve(h, _p(board), int(stm), max_nodes, _p(moves), _p(scores), _p(result), _p(flags))
    return n, moves[:n].copy(), scores[:n].copy(), int(result[0]), int(flags[0])


@pytest.mark.parametrize("rules,size,max_nodes,max_fill", [(0, 15, 1, 0.5), (0, 15, 100, 0.4), (1, 15, 1000, 0.1), (3, 20, 100, 0.3), (4, 15, 100, 0.4),
                                                             (2, 15, 100, 0.4), (0, 12, 200, 0.08)])
def test_agb_solve_matches_reference(ref, ref_fast, rules, size, max_nodes, max_fill):
    """AlphaBetaSearch::solve (cleared table per position) vs agb_solve: same action list in the same order, same action scores, position
    score, must-defend flag and node count. RENJU runs on the reference's Release build (its debug build asserts on unreachable boards)."""
    import alphagomoku_b200 as agb
    lib = (ref_fast if rules == 2 else ref).lib
    cells = size * size
    h, keys = _ref_solver(lib, rules, size)
    eng = agb.Engine(agb.GameConfig(agb.GameRules(rules), size, size), max_boards=256, solver_table_entries=4 * 1024 * 1024)
    eng.set_solver_keys(keys)
    rng = np.random.default_rng(40 + rules + max_nodes)
    boards = random_boards(rng, size, 160, max_fill=max_fill)
    stm = np.array([1 if (np.count_nonzero(b) % 2 == 0) else 2 for b in boards], np.int8)
    scores, n_actions, moves, action_scores, flags = eng.solve(boards, stm, max_nodes)
    total_nodes = 0
    for i in range(len(boards)):
        n, rm, rs, rscore, rflags = _ref_solve(lib, h, boards[i], stm[i], max_nodes, cells)
        assert n_actions[i] == n, i
        assert (moves[i, :n] == rm).all(), (i, moves[i, :n], rm)
        assert (action_scores[i, :n] == rs).all(), i
        assert scores[i] == rscore, (i, hex(scores[i]), hex(rscore))
        assert (flags[i] & 1) == (rflags & 1) and (flags[i] >> 8) == (rflags >> 8), (i, hex(flags[i]), hex(rflags))
        total_nodes += rflags >> 8
    assert total_nodes >= len(boards)
    lib.agref_solver_destroy(h)
    eng.close()


def test_generate_openings(ref):
    """OpeningGenerator::generate on the device solver + network: every opening is a legal, unfinished position that the reference's
    solver cannot prove within 1000 positions either."""
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob
    size, rules = 15, 1
    eng = agb.Engine(agb.GameConfig(agb.GameRules(rules), size, size), max_boards=256, blocks=2, filters=64, seed=3)
    eng.load_weights(netblob.pack(netblob.random_tensors(size, size, 2, 64, False, seed=5), size, size, 2, 64, False))
    boards, stm = eng.generate_openings(48)
    h, _ = _ref_solver(ref.lib, rules, size)
    distinct = set()
    for b, s in zip(boards, stm):
        n_cross, n_circle = int((b == 1).sum()), int((b == 2).sum())
        assert n_cross + n_circle >= 1 and n_cross - n_circle in (0, 1)
        assert s == (1 if n_cross == n_circle else 2)
        _, _, _, score, _ = _ref_solve(ref.lib, h, b, s, 1000, size * size)
        assert (score >> 13) & 3 == 2, hex(score)  # ProvenValue::UNKNOWN
        distinct.add(b.tobytes())
    assert len(distinct) > 40
    # the same seed gives the same openings again
    eng.seed_openings(3)
    again, _ = eng.generate_openings(48)
    assert (again == boards).all()
    eng.close()


def test_low_register_solver_build_matches_reference(ref, monkeypatch):
    """Launches with more than 28 games per SM use a build of the solver kernel capped at 32 registers (56 resident games per SM);
    AGB_SOLVER_DENSE forces it here so that it is checked against the reference as well."""
    monkeypatch.setenv("AGB_SOLVER_DENSE", "1")
    test_agb_solve_matches_reference(ref, None, 1, 15, 100, 0.3)


def test_move_generator_known_answers_on_the_device(golden):
    """The reference's own move generator tests (test/search/alpha_beta/test_move_generator.cpp, OPTIMAL mode) through agb_solve with a
    budget of one position, i.e. K5's root generation on the device: list size, must-defend flag, members and scores as the reference's
    authors wrote them down."""
    import alphagomoku_b200 as agb
    entries = [e for e in golden[0] if e["kind"] == "movegen" and e["mode"] == "OPTIMAL"]
    assert len(entries) >= 45
    engines = {}
    for e in entries:
        key = (e["rules"], e["size"])
        if key not in engines:
            engines[key] = agb.Engine(agb.GameConfig(agb.GameRules(e["rules"]), e["size"], e["size"]), max_boards=8)
        board = np.array(e["board"], np.int8)
        _, n_actions, moves, action_scores, flags = engines[key].solve(board[None], np.array([e["stm"]], np.int8), 1)
        n = int(n_actions[0])
        cells = {((int(m) >> 2) & 127, (int(m) >> 9) & 127): int(s) for m, s in zip(moves[0, :n], action_scores[0, :n])}
        if "size_eq" in e:
            assert n == e["size_eq"], (e["size_eq"], n)
        if "size_ge" in e:
            assert n >= e["size_ge"]
        if "must_defend" in e:
            assert bool(flags[0] & 1) == e["must_defend"]
        for r, c in e["contains"]:
            assert (r, c) in cells
        for r, c, s in e["scores"]:
            assert cells.get((r, c)) == s
    for eng in engines.values():
        eng.close()
