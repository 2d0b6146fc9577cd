"""Evaluation games between two engines (alphagomoku_b200/arena.py on top of agb_think): the data-parallel core of the reference's arena
(src/evaluation). Checked for what must hold whatever the networks are: legal alternating moves, outcomes that the reference's getOutcome
confirms, determinism, and a 50 % score when a network plays itself with swapped colours."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(agb, netblob, seed, games, rules=1, size=15, sims=60):
    eng = agb.Engine(agb.GameConfig(agb.GameRules(rules), size, size), max_boards=games * 4, blocks=2, filters=64, games=games, max_batch_size=4,
                     max_simulations=sims, solver_max_positions=20)
    eng.load_weights(netblob.pack(netblob.random_tensors(size, size, 2, 64, False, seed=seed), size, size, 2, 64, False))
    return eng


def test_think_returns_one_legal_move_per_active_game(ref):
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob
    size, games = 15, 12
    eng = _engine(agb, netblob, 1, games)
    rng = np.random.default_rng(3)
    boards = np.zeros((games, size * size), np.int8)
    for b in boards:
        idx = rng.permutation(size * size)[:8]
        b[idx[0::2]], b[idx[1::2]] = 1, 2
    stm = np.ones(games, np.int8)
    active = (np.arange(games) % 3 != 0).astype(np.int8)
    moves, values = eng.think(boards, stm, active)
    for g in range(games):
        if not active[g]:
            assert moves[g] == 0
            continue
        mv = int(moves[g])
        assert (mv & 3) == 1 and boards[g, ((mv >> 2) & 127) * size + ((mv >> 9) & 127)] == 0
        assert 0.0 <= values[g, 0] <= 1.0 and 0.0 <= values[g, 1] <= 1.0
    again, _ = eng.think(boards, stm, active)
    assert (again == moves).all()  # fresh trees every time: same position, same decision
    eng.close()


def test_match_between_two_engines(ref):
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob, arena
    size, n_openings = 15, 6
    a = _engine(agb, netblob, 11, 2 * n_openings)
    b = _engine(agb, netblob, 22, 2 * n_openings)
    rng = np.random.default_rng(5)
    openings = np.zeros((n_openings, size * size), np.int8)
    for o in openings:
        idx = rng.permutation(size * size)[:4]
        o[idx[0::2]], o[idx[1::2]] = 1, 2
    stm = np.ones(n_openings, np.int8)
    match = arena.play_match(a, b, openings, stm, swap_colours=True, max_plies=60)
    assert match["n_games"] == 2 * n_openings and 0.0 <= match["score_a"] <= 1.0
    for g, game in enumerate(match["games"]):
        board = openings[g % n_openings].copy()
        sign = 1
        for k, mv in enumerate(game["moves"]):
            row, col = (mv >> 2) & 127, (mv >> 9) & 127
            assert (mv & 3) == sign and board[row * size + col] == 0
            board[row * size + col] = sign
            expected = ref.outcome(1, size, board, row, col, sign, 0)
            last = k == len(game["moves"]) - 1
            if not last:
                assert expected == 0, (g, k)
            elif game["outcome"] != "DRAW" or len(game["moves"]) < 60:
                assert arena.OUTCOME_NAMES[expected] == game["outcome"], (g, expected, game["outcome"])
            sign = 3 - sign
    # a network against itself, colours swapped: every opening is won once by each colour's owner or drawn twice -> exactly 50 %
    mirror = arena.play_match(a, a, openings, stm, swap_colours=True, max_plies=60)
    assert abs(mirror["score_a"] - 0.5) < 1e-9
    first, second = mirror["games"][:n_openings], mirror["games"][n_openings:]
    assert all(x["moves"] == y["moves"] for x, y in zip(first, second))
    a.close()
    b.close()


def test_players_reroot_their_trees_like_the_reference_player(ref):
    """Move for move against the reference's Player (src/evaluation/Player.cpp driven like EvaluationGame.cpp:95-150): two networks, every
    player keeps its tree between its moves (Player::setBoard -> Tree::setBoard -> NodeCache::cleanup re-roots it two plies further), the solver's
    table lives on across moves. Same moves, same root visit counts, same simulation counts for whole games."""
    import ctypes
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob
    import refapi
    size, games, sims, batch, solver = 15, 4, 100, 4, 50
    cells = size * size
    engines = []
    for seed in (11, 22):
        eng = agb.Engine(agb.GameConfig(agb.GameRules.STANDARD, size, size), max_boards=games * batch, blocks=2, filters=64, games=games, max_batch_size=batch,
                         max_simulations=sims, solver_max_positions=solver, solver_table_entries=4 * 1024 * 1024, max_nodes_per_game=2048)
        eng.load_weights(netblob.pack(netblob.random_tensors(size, size, 2, 64, False, seed=seed), size, size, 2, 64, False))
        engines.append(eng)
    lib = ctypes.CDLL(refapi.REF_LIB)
    lib.agref_player_create.restype = ctypes.c_void_p
    lib.agref_player_create.argtypes = [ctypes.c_int] * 6 + [ctypes.c_char_p, ctypes.c_float, refapi.EVAL_FN, ctypes.c_void_p]
    lib.agref_player_move.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.agref_player_solver_keys.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    callbacks, players = [], [[], []]
    for side, eng in enumerate(engines):
        def callback(ctx, features, n, rows, cols, policy, value, action_values, moves_left, eng=eng):
            f = np.ctypeslib.as_array(features, shape=(n, rows * cols)).copy()
            p, v, _ = eng.forward(f)
            np.ctypeslib.as_array(policy, shape=(n, rows * cols))[:] = p
            np.ctypeslib.as_array(value, shape=(n, 3))[:] = v
            np.ctypeslib.as_array(action_values, shape=(n, rows * cols, 3))[:] = 0.0
            np.ctypeslib.as_array(moves_left, shape=(n,))[:] = 0.0
        cb = refapi.EVAL_FN(callback)
        callbacks.append(cb)
        keys = np.zeros((games, 2 * cells, 2), np.uint64)
        for g in range(games):
            h = ctypes.c_void_p(lib.agref_player_create(1, size, size, batch, sims, solver, b"parent", 1.25, cb, None))
            players[side].append(h)
            lib.agref_player_solver_keys(h, refapi._p(keys[g]))
        eng.set_solver_keys(keys)
        eng.selfplay_reset()  # brand-new players
    rng = np.random.default_rng(8)
    boards = np.zeros((games, cells), np.int8)
    for g in range(games):
        idx = (7 + rng.integers(-3, 4, 4)) * size + 7 + rng.integers(-3, 4, 4)
        idx = np.unique(idx)[:2 * (len(np.unique(idx)) // 2)]
        boards[g, idx[0::2]], boards[g, idx[1::2]] = 1, 2
    ref_boards = boards.copy()
    outcome = np.zeros(games, np.int8)
    plies = 0
    for ply in range(40):
        side, stm = ply % 2, 1 + ply % 2
        active = (outcome == 0).astype(np.int8)
        if not active.any():
            break
        moves, _ = engines[side].think(boards, np.full(games, stm, np.int8), active)
        for g in np.flatnonzero(active):
            visits, count = np.zeros(cells, np.int32), ctypes.c_int32(0)
            ref_move = lib.agref_player_move(players[side][g], refapi._p(ref_boards[g]), stm, refapi._p(visits), ctypes.byref(count))
            dv, _, _, _, dn = engines[side].get_root(int(g))
            assert int(moves[g]) == ref_move, (ply, g, int(moves[g]), ref_move)
            assert dn == count.value and (dv == visits).all(), (ply, g, dn, count.value)
            row, col = (ref_move >> 2) & 127, (ref_move >> 9) & 127
            boards[g, row * size + col] = stm
            ref_boards[g, row * size + col] = stm
            plies += 1
        idx = np.flatnonzero(active)
        outcome[idx] = engines[side].get_outcomes(boards[idx], moves[idx])
    assert plies >= 60
    for side in (0, 1):
        for h in players[side]:
            lib.agref_player_destroy(h)
        engines[side].close()
