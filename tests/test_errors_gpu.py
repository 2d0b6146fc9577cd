"""Error behaviour of the C ABI on the GPU: the reference throws std::logic_error / std::runtime_error for misuse and asserts its
invariants (NNEvaluator.cpp:149-150, 186-187; AGNetwork.cpp:186-189); here every misuse is a negative status with a message, and a
bounded device structure that overflows is reported, never silently truncated."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_misuse_is_reported():
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob
    size = 15
    eng = agb.Engine(agb.GameConfig(agb.GameRules.STANDARD, size, size), max_boards=8, blocks=1, filters=64)
    boards = np.zeros((16, size * size), np.int8)
    stm = np.ones(16, np.int8)
    with pytest.raises(agb.AgbError, match="max_boards"):
        eng.set_boards(boards, stm)  # more boards than slots
    with pytest.raises(agb.AgbError, match="no weights"):
        eng.evaluate(boards[:4], stm[:4])  # forward before loadGraph
    with pytest.raises(agb.AgbError, match="wrong size"):
        eng.load_weights(np.zeros(10, np.float32))
    with pytest.raises(agb.AgbError, match="without games"):
        eng.step(1)
    eng.load_weights(netblob.pack(netblob.random_tensors(size, size, 1, 64, False), size, size, 1, 64, False))
    with pytest.raises(agb.AgbError, match="symmetry"):
        eng.evaluate(boards[:2], stm[:2], symmetry=[0, 9])
    policy, value, _ = eng.evaluate(boards[:2], stm[:2])
    assert np.isfinite(policy).all() and abs(float(value[0].sum()) - 1.0) < 1e-5
    eng.close()
    for bad in (dict(filters=48, blocks=1), dict(games=4, max_batch_size=2, blocks=1, filters=64, max_boards=8, max_simulations=40), dict(games=4, max_batch_size=4, blocks=1, filters=64, max_boards=8), dict(pipeline_groups=9, games=16, blocks=1, filters=64, max_boards=128),
                dict(solver_table_entries=1000, solver_max_positions=10, games=2, blocks=1, filters=64, max_boards=64)):
        kwargs = dict(max_boards=8)
        kwargs.update(bad)
        with pytest.raises(agb.AgbError):
            agb.Engine(agb.GameConfig(agb.GameRules.STANDARD, size, size), **kwargs)


def test_arena_overflow_is_reported():
    """A node arena that is too small for the search must surface as AGB_EOVERFLOW with the flag word, not as a corrupted tree."""
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob
    size, games = 15, 4
    eng = agb.Engine(agb.GameConfig(agb.GameRules.FREESTYLE, size, size), max_boards=games * 8, blocks=1, filters=64, games=games, max_batch_size=8,
                     max_simulations=400, max_nodes_per_game=24, max_edges_per_game=24 * 230)
    eng.load_weights(netblob.pack(netblob.random_tensors(size, size, 1, 64, False), size, size, 1, 64, False))
    eng.selfplay_reset()
    with pytest.raises(agb.AgbError, match="overflow"):
        eng.step(40)
    assert eng.stats()["overflow_flags"] & 6  # nodes (2) or edges (4)
    eng.close()


def test_overflow_is_reported_once_and_the_engine_stays_usable():
    """The device status word is cleared when it is reported: one overflow fails one call, not every later one; the flags stay visible
    in AgbStats until the next reset."""
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob
    size, games = 15, 4
    eng = agb.Engine(agb.GameConfig(agb.GameRules.FREESTYLE, size, size), max_boards=games * 8, blocks=1, filters=64, games=games, max_batch_size=8,
                     max_simulations=400, max_nodes_per_game=24, max_edges_per_game=24 * 230)
    eng.load_weights(netblob.pack(netblob.random_tensors(size, size, 1, 64, False), size, size, 1, 64, False))
    eng.selfplay_reset()
    with pytest.raises(agb.AgbError, match="overflow"):
        eng.step(40)
    assert eng.stats()["overflow_flags"] & 6
    boards = np.zeros((4, size * size), np.int8)
    features = eng.set_boards(boards, np.ones(4, np.int8))  # not poisoned by the earlier overflow
    assert features.shape == (4, size * size)
    eng.selfplay_reset()
    assert eng.stats()["overflow_flags"] == 0
    eng.close()


def test_bad_moves_and_cells_are_rejected():
    """PatternCalculator::addMove / undoMove assert on off-board or occupied cells (PatternCalculator.cpp:70, 89); here they are AGB_EINVAL and
    the slot keeps its state."""
    import alphagomoku_b200 as agb
    size = 15
    eng = agb.Engine(agb.GameConfig(agb.GameRules.STANDARD, size, size), max_boards=4)
    boards = np.zeros((2, size * size), np.int8)
    boards[0, 7 * size + 7] = 1
    stm = np.array([2, 1], np.int8)
    eng.set_boards(boards, stm)
    before = eng.get_state(2)
    with pytest.raises(agb.AgbError, match="not on the board"):
        eng.add_moves([agb.engine.move_to_short(15, 3, 1), 0])  # row 15 of a 15 x 15 board
    with pytest.raises(agb.AgbError, match="not on the board"):
        eng.add_moves([3 | (2 << 2) | (2 << 9), 0])  # Sign::ILLEGAL
    with pytest.raises(agb.AgbError, match="invalid input"):
        eng.add_moves([agb.engine.move_to_short(7, 7, 2), 0])  # occupied cell
    with pytest.raises(agb.AgbError, match="invalid input"):
        eng.undo_moves([0, agb.engine.move_to_short(3, 3, 1)])  # nothing to undo there
    after = eng.get_state(2)
    for key in before:
        assert (before[key] == after[key]).all(), key
    eng.add_moves([agb.engine.move_to_short(7, 8, 2), 0])  # a legal move still works afterwards
    bad = boards.copy()
    bad[1, 5] = 3
    with pytest.raises(agb.AgbError, match="board cells"):
        eng.set_boards(bad, stm)
    with pytest.raises(agb.AgbError, match="sign_to_move"):
        eng.set_boards(boards, np.array([0, 1], np.int8))
    eng.close()


def test_full_finished_queue_keeps_the_records(monkeypatch):
    """A finished-game queue the host does not drain pauses the finished games instead of dropping their records: every game that ended is
    popped exactly once, before or after the overflow report."""
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob, dataset
    monkeypatch.setenv("AGB_FINISHED_QUEUE_BYTES", "6000")  # two or three short games
    size, games = 9, 16
    eng = agb.Engine(agb.GameConfig(agb.GameRules.FREESTYLE, size, size), max_boards=games * 4, blocks=1, filters=64, games=games, max_batch_size=4,
                     max_simulations=50, max_nodes_per_game=400, max_edges_per_game=400 * 81)
    eng.load_weights(netblob.pack(netblob.random_tensors(size, size, 1, 64, False), size, size, 1, 64, False))
    eng.selfplay_reset()
    popped, reports = 0, 0
    for _ in range(60):
        try:
            eng.step(10)
        except agb.AgbError as err:
            assert "finished-game queue is full" in str(err)
            reports += 1
        blob, n = eng.pop_finished()
        assert len(dataset.split_records(blob, n)) == n
        popped += n
    for _ in range(3):  # let the waiting games publish
        try:
            eng.step(1)
        except agb.AgbError:
            pass
        popped += eng.pop_finished()[1]
    st = eng.stats()
    assert reports > 0, "the queue never filled: the test does not exercise the path"
    assert st["nb_games_finished"] > 0 and abs(int(st["nb_games_finished"]) - popped) <= games  # at most the games still waiting
    eng.close()
