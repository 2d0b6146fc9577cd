"""GPU parity tests for K1/K2/K3 (+ augment, outcome) through the C ABI, against the reference's own classes
(oracle/_ref) and the committed golden vectors. Bit-exact: integer work."""
import ctypes

import numpy as np
import pytest

from conftest import random_boards

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def agb():
    import alphagomoku_b200
    return alphagomoku_b200


def make_engine(agb, rules, size, max_boards=4096):
    return agb.Engine(agb.GameConfig(agb.GameRules(rules), size, size), max_boards=max_boards)


@pytest.mark.parametrize("rules", [0, 1, 2, 3, 4])
def test_device_tables_match_reference(agb, ref, rules):
    eng = make_engine(agb, rules, 15, 16)
    pt, ho, th = eng.get_tables()
    rpt, rho, rth, _ = ref.tables(rules)
    assert (pt == rpt).all() and (ho == rho).all() and (th == rth).all()
    eng.close()


def _compare_state(state, i, st, cells, hist_only=False):
    if not hist_only:
        assert (state["pattern_types"][i] == st["pattern_types"]).all()
        assert (state["threats"][i] == st["threats"]).all()
        assert (state["legal"][i] == st["legal"]).all()
        assert (state["forbidden"][i] == st["forbidden"]).all()
    assert (state["hist_counts"][i] == st["hist_counts"]).all()
    for colour in range(2):
        for t in range(1, 10):
            n = st["hist_counts"][colour][t]
            assert (state["hist_cells"][i][colour][t][:n] == st["hist_cells"][colour][t][:n]).all(), (colour, t)


def test_golden_states_through_cuda(agb, golden):
    """Committed reference outputs (boards of the reference's unit tests + seeded random boards), all rules and sizes."""
    _, states = golden
    keys = sorted(set(zip(states["rules"].tolist(), states["size"].tolist())))
    for rules, size in keys:
        sel = np.nonzero((states["rules"] == rules) & (states["size"] == size))[0]
        c = size * size
        eng = make_engine(agb, rules, size, len(sel))
        feats = eng.set_boards(states["board"][sel][:, :c], states["stm"][sel])
        st = eng.get_state(len(sel))
        assert (feats == states["features"][sel][:, :c]).all()
        assert (st["pattern_types"] == states["pattern_types"][sel][:, :c]).all()
        assert (st["threats"] == states["threats"][sel][:, :c]).all()
        assert (st["forbidden"] == states["forbidden"][sel][:, :c]).all()
        assert (st["hist_counts"] == states["hist_counts"][sel]).all()
        for j, i in enumerate(sel):
            for colour in range(2):
                for t in range(1, 10):
                    n = states["hist_counts"][i][colour][t]
                    assert (st["hist_cells"][j][colour][t][:n] == states["hist_cells"][i][colour][t][:n]).all()
        eng.close()


def test_known_answers_through_cuda(agb, golden):
    """The reference authors' hand-written expectations, evaluated by the CUDA kernels."""
    known, _ = golden
    engines = {}
    for k in known:
        key = (k["rules"], k["size"])
        if key not in engines:
            engines[key] = make_engine(agb, k["rules"], k["size"], 8)
        eng = engines[key]
        board = np.array(k["board"], np.int8)
        size = k["size"]
        if k["kind"] == "outcome":
            got = eng.get_outcomes(board[None], [agb.move_to_short(k["row"], k["col"], k["sign"])])[0]
            assert got == k["expected"], k
        elif k["kind"] == "forbidden" and board[k["row"] * size + k["col"]] == 0:
            eng.set_boards(board[None], [1])
            assert bool(eng.get_state(1, histograms=False)["forbidden"][0][k["row"] * size + k["col"]]) == k["expected"], k
        elif k["kind"] == "feature_bit":
            f = eng.set_boards(board[None], [k["stm"]])[0]
            assert bool((int(f[k["row"] * size + k["col"]]) >> k["bit"]) & 1) == k["expected"], k
    for eng in engines.values():
        eng.close()


@pytest.mark.parametrize("rules,size", [(0, 15), (1, 15), (2, 15), (3, 20), (4, 20), (2, 20)])
def test_set_boards_random_vs_reference(agb, ref, rules, size):
    rng = np.random.default_rng(1000 + 10 * rules + size)
    n = 256
    boards = random_boards(rng, size, n)
    stm = rng.integers(1, 3, n).astype(np.int8)
    eng = make_engine(agb, rules, size, n)
    feats = eng.set_boards(boards, stm)
    state = eng.get_state(n)
    for i in range(n):
        st = ref.set_board(rules, size, boards[i], stm[i])
        assert (feats[i] == st["features"]).all(), i
        _compare_state(state, i, st, size * size)
    # encode from the persistent state equals the fused set+encode
    assert (eng.encode(n) == feats).all()
    eng.close()


@pytest.mark.parametrize("rules,size", [(0, 15), (2, 15), (3, 20)])
def test_add_undo_sequences_vs_reference(agb, ref, rules, size):
    """K2: random games played forward and partly backward; state (incl. order-sensitive threat lists) after every ply."""
    rng = np.random.default_rng(77 + rules)
    n = 24
    boards = random_boards(rng, size, n, max_fill=0.3)
    stm = np.ones(n, np.int8)
    # legal alternation: make stone counts consistent with the side to move
    for i in range(n):
        stm[i] = 1 if (boards[i] == 1).sum() == (boards[i] == 2).sum() else 2
    eng = make_engine(agb, rules, size, n)
    eng.set_boards(boards, stm)
    cur = boards.copy()
    history = [[] for _ in range(n)]
    # Two reference calculators per board: `refs` is dumped in full every ply; `clean` only ever has its threat lists read,
    # because PatternCalculator::isForbidden (renju) runs addMove/undoMove internally and thereby reorders the lists.
    refs, clean = [], []
    import refapi
    for i in range(n):
        for group in (refs, clean):
            r = refapi.RefOracle()
            r.lib.agref_calc_set_board(r.calc(rules, size), ctypes.c_void_p(cur[i].ctypes.data), int(stm[i]))
            group.append(r)
    for ply in range(30):
        moves = np.zeros(n, np.uint16)
        undo = (ply % 7 == 6)
        for i in range(n):
            if undo and history[i]:
                row, col, sign = history[i].pop()
                cur[i][row * size + col] = 0
                refs[i].undo_move(rules, size, row, col, sign)
                clean[i].undo_move(rules, size, row, col, sign)
                moves[i] = agb.move_to_short(row, col, sign)
            elif not undo:
                empty = np.nonzero(cur[i] == 0)[0]
                if len(empty) == 0:
                    continue
                cell = int(empty[rng.integers(len(empty))])
                row, col = divmod(cell, size)
                sign = 1 if (cur[i] == 1).sum() == (cur[i] == 2).sum() else 2
                cur[i][cell] = sign
                history[i].append((row, col, sign))
                refs[i].add_move(rules, size, row, col, sign)
                clean[i].add_move(rules, size, row, col, sign)
                moves[i] = agb.move_to_short(row, col, sign)
        if undo:
            eng.undo_moves(moves)
        else:
            eng.add_moves(moves)
        feats = eng.encode(n)
        state = eng.get_state(n)
        for i in range(n):
            st = refs[i].dump(rules, size)
            assert (feats[i] == st["features"]).all(), (ply, i)
            st["hist_counts"], st["hist_cells"] = (clean[i].dump(rules, size, hist_only=True)[k] for k in ("hist_counts", "hist_cells"))
            _compare_state(state, i, st, size * size)
    eng.close()


def test_augment_and_outcomes_vs_reference(agb, ref):
    rng = np.random.default_rng(3)
    for rules, size in [(0, 15), (2, 15), (4, 20)]:
        n = 128
        boards = random_boards(rng, size, n)
        stm = rng.integers(1, 3, n).astype(np.int8)
        eng = make_engine(agb, rules, size, n)
        feats = eng.set_boards(boards, stm)
        sym = rng.integers(0, 8, n).astype(np.int8)
        aug = eng.augment(feats, sym)
        moves = np.zeros(n, np.uint16)
        expected = np.zeros(n, np.int8)
        for i in range(n):
            assert (aug[i] == ref.augment(feats[i], size, sym[i])).all()
            occ = np.nonzero(boards[i])[0]
            cell = int(occ[rng.integers(len(occ))]) if len(occ) else 0
            row, col = divmod(cell, size)
            sign = int(boards[i][cell]) if len(occ) else 1
            moves[i] = agb.move_to_short(row, col, sign)
            expected[i] = ref.outcome(rules, size, boards[i], row, col, sign, size * size)
        assert (eng.get_outcomes(boards, moves) == expected).all()
        eng.close()


def test_size_independent_properties_at_full_size(agb):
    """BASELINE micro-benchmark size (2^17 boards per call here): add then undo is the identity on the whole state, and
    the 8 symmetries commute with set+encode (augment(encode(b)) == encode(symmetry(b)))."""
    rng = np.random.default_rng(9)
    size, n = 15, 1 << 17
    eng = make_engine(agb, 1, size, n)
    fill = rng.random((n, size * size))
    boards = np.zeros((n, size * size), np.int8)
    boards[fill < 0.15] = 1
    boards[(fill >= 0.15) & (fill < 0.30)] = 2
    stm = rng.integers(1, 3, n).astype(np.int8)
    feats = eng.set_boards(boards, stm)
    before = eng.get_state(n, histograms=False)
    empty_choice = np.argmax(boards == 0, axis=1)
    moves = np.array([agb.move_to_short(c // size, c % size, s) for c, s in zip(empty_choice, stm)], np.uint16)
    eng.add_moves(moves)
    eng.undo_moves(moves)
    after = eng.get_state(n, histograms=False)
    for key in ("pattern_types", "threats", "legal", "forbidden"):
        assert (before[key] == after[key]).all(), key
    assert (eng.encode(n) == feats).all()
    # symmetry property on a slice
    m = 4096
    for mode in range(8):
        sym = np.full(m, mode, np.int8)
        aug = eng.augment(feats[:m], sym)
        b2 = boards[:m].reshape(m, size, size)
        if mode == 1:
            b2 = b2[:, ::-1, :]
        elif mode == 2:
            b2 = b2[:, :, ::-1]
        elif mode == 3:
            b2 = b2[:, ::-1, ::-1]
        elif mode == 4:
            b2 = b2.transpose(0, 2, 1)
        elif mode == 5:
            b2 = b2[:, ::-1, ::-1].transpose(0, 2, 1)
        elif mode == 6:
            b2 = np.rot90(boards[:m].reshape(m, size, size), k=1, axes=(1, 2))
        elif mode == 7:
            b2 = np.rot90(boards[:m].reshape(m, size, size), k=-1, axes=(1, 2))
        f2 = eng.set_boards(np.ascontiguousarray(b2).reshape(m, -1), stm[:m])
        assert (f2 == aug).all(), mode
    eng.close()
