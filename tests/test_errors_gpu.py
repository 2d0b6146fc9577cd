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
    for bad in (dict(filters=48, blocks=1), dict(games=4, max_batch_size=4, blocks=1, filters=64, max_boards=8), dict(pipeline_groups=5, games=8, blocks=1, filters=64, max_boards=64),
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
