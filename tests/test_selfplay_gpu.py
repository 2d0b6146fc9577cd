"""GPU parity tests for K6/K7 and the per-ply driver: the device lockstep engine against the reference's own
Tree / Search / EdgeSelector / EdgeGenerator / NodeCache classes (oracle/_ref), game by game and step by step.

The reference's evaluator is replaced by a callback that runs the SAME device network (agb_forward) on the features the
reference computed, so both sides see bit-identical policy/value inputs and every tree quantity must match exactly:
visit counts, priors, edge and root values (float bits), the moves played and the game outcomes."""
import numpy as np
import pytest

from conftest import random_boards

pytestmark = pytest.mark.gpu


def _openings(rng, size, n):
    boards = np.zeros((n, size * size), np.int8)
    stm = np.ones(n, np.int8)
    centre = size // 2
    for g in range(1, n):
        k = int(rng.integers(1, 7))
        cells = set()
        while len(cells) < k:
            r, c = centre + int(rng.integers(-3, 4)), centre + int(rng.integers(-3, 4))
            cells.add(r * size + c)
        for j, cell in enumerate(sorted(cells, key=lambda x: rng.random())):
            boards[g, cell] = 1 + (j % 2)
        stm[g] = 1 if k % 2 == 0 else 2
    return boards, stm


@pytest.mark.parametrize("rules,q_head,init_to,batch,sims,solver", [
    (0, False, "parent", 4, 60, 0), (1, True, "q_head", 8, 80, 0), (2, False, "parent", 3, 50, 0), (0, False, "loss", 1, 55, 0),
    # K5 on: Search::solve with tss max_positions = 1 (static solver) on every leaf, proven leaves skip the network
    (0, False, "parent", 4, 60, 1), (1, True, "q_head", 8, 80, 1), (4, False, "parent", 5, 70, 1), (3, True, "q_head", 2, 55, 1),
    # alpha-beta search with the per-game transposition table (tss max_positions > 1), renju with forbidden moves in the solver
    (0, False, "parent", 4, 60, 100), (1, True, "q_head", 8, 80, 100), (4, False, "parent", 3, 55, 30), (2, False, "parent", 4, 60, 1),
    (2, True, "q_head", 8, 80, 100)])
def test_lockstep_engine_matches_reference_search(ref, rules, q_head, init_to, batch, sims, solver):
    _run_case(ref, rules, q_head, init_to, batch, sims, solver)


@pytest.mark.parametrize("rules,q_head,init_to,batch,sims,solver", [(3, False, "parent", 4, 60, 0), (3, True, "q_head", 6, 70, 50)])
def test_lockstep_engine_caro_20x20(ref, rules, q_head, init_to, batch, sims, solver):
    """BASELINE configs[3]: caro on a 20x20 board (K4 runs one board per CTA pair, 7-word bitboards in the tree)."""
    _run_case(ref, rules, q_head, init_to, batch, sims, solver, size=20)


@pytest.mark.parametrize("rules,q_head,solver,max_children,threshold", [(0, False, 0, 12, 1.0e-4), (1, True, 50, 20, 0.02), (4, False, 1, 6, 0.1)])
def test_lockstep_engine_with_policy_pruning(ref, rules, q_head, solver, max_children, threshold):
    """MCTSConfig::max_children / policy_expansion_threshold: prune_weak_moves on unproven positions (std::partial_sort by prior, then the
    threshold on the kept priors), EdgeGenerator.cpp:69-84."""
    _run_case(ref, rules, q_head, "parent", 4, 60, solver, max_children=max_children, threshold=threshold)


@pytest.mark.parametrize("selector,rules,solver", [("best", 0, 0), ("max_value", 1, 20), ("lcb", 0, 1), ("max_policy", 4, 0), ("min_visit", 0, 0)])
def test_lockstep_engine_final_selectors(ref, selector, rules, solver):
    """SelfplayConfig::final_selector: the move actually played is chosen by another EdgeSelector (EdgeSelector.cpp:426-536, 680-711)."""
    _run_case(ref, rules, True, "parent", 4, 60, solver, final_selector=selector)


@pytest.mark.parametrize("rules,solver", [(0, 0), (1, 30)])
def test_lockstep_engine_policy_temperature_zero(ref, rules, solver):
    """MCTSConfig::policy_temperature = 0: initialize_edges gives prior 1 to the cells with the largest policy value and 0 to the rest
    (EdgeGenerator.cpp:90-101); positions that skipped the network have an all-zero policy, hence prior 1 everywhere."""
    _run_case(ref, rules, False, "parent", 4, 60, solver, temperature=0.0)


def _run_case(ref, rules, q_head, init_to, batch, sims, solver, size=15, max_children=0, threshold=1.0e-4, final_selector="max_visit", temperature=1.0):
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob
    import refapi
    games = 6
    blocks, filters = 2, 64
    # short games, so that game ends, restarts and finished-game records are exercised too (longer with the solver, so that real threats appear)
    draw_after = 14 if solver == 0 else 28
    eng = agb.Engine(agb.GameConfig(agb.GameRules(rules), size, size, draw_after), max_boards=256, blocks=blocks, filters=filters, q_head=q_head,
                     games=games, max_batch_size=batch, max_simulations=sims, init_to=init_to, max_nodes_per_game=1024, solver_max_positions=solver,
                     solver_table_entries=4 * 1024 * 1024 if solver > 1 else 0, pipeline_groups=2 if solver > 0 else 1, max_children=max_children,
                     policy_expansion_threshold=threshold, final_selector=final_selector, policy_temperature=temperature)  # the reference's table size (AlphaBetaSearch.cpp:55)
    eng.load_weights(netblob.pack(netblob.random_tensors(size, size, blocks, filters, q_head, seed=5), size, size, blocks, filters, q_head))

    def evaluate(features):
        return eng.forward(features, want_q=q_head)

    rng = np.random.default_rng(rules + 17)
    boards, stm = _openings(rng, size, games)
    refs = []
    # renju with the solver: the reference's debug build asserts in MoveGenerator.cpp:999 whenever black can win in one move, so that
    # case runs its Release build (nothing on this path draws random numbers once the hash keys are shared)
    fast = solver > 0 and rules == 2
    for g in range(games):
        r = refapi.RefSelfplay(rules, size, evaluate, max_batch_size=batch, max_simulations=sims, init_to=init_to, use_solver=solver > 0,
                                solver_max_positions=max(solver, 1), draw_after=draw_after, fast=fast, max_children=max_children,
                                policy_expansion_threshold=threshold, final_selector=final_selector, policy_temperature=temperature)
        refs.append(r)
    if solver > 0:
        eng.set_solver_keys(np.stack([r.solver_keys() for r in refs]))
    eng.selfplay_reset(boards, stm)
    for g in range(games):
        refs[g].set_position(boards[g], stm[g])
    active = [True] * games
    moves_checked = 0
    ref_records = []
    for step in range(400 if solver == 0 else 1200):
        eng.step(1)
        for g in range(games):
            if not active[g]:
                continue
            status = refs[g].step()
            rb, rstm, routcome, rlast = refs[g].board()
            if status == 2:
                active[g] = False  # the device engine restarts the game; the reference instance stops here
                ref_records.append(refs[g].record())
                moves_checked += 1
                continue
            db, dstm, _ = eng.get_board(g)
            assert (db == rb).all() and dstm == rstm, (step, g)
            rv, rp, rq, rval, rn = refs[g].root()
            dv, dp, dq, dval, dn = eng.get_root(g)
            assert dn == rn, (step, g, dn, rn)
            assert (dv == rv).all(), (step, g)
            assert (dp.view(np.uint32) == rp.view(np.uint32)).all(), (step, g)
            assert (dq.view(np.uint32) == rq.view(np.uint32)).all(), (step, g)
            assert (dval.view(np.uint32) == rval.view(np.uint32)).all(), (step, g, dval, rval)
            moves_checked += status
        if not any(active):
            break
    st = eng.stats()
    assert st["overflow_flags"] == 0
    # K8: every finished reference game must appear byte for byte among the device's format-201 records
    from alphagomoku_b200 import dataset
    blob, n_finished = eng.pop_finished()
    device_records = dataset.split_records(blob, n_finished)
    assert len(ref_records) >= 1
    for rec in ref_records:
        assert rec in device_records, dataset.parse_record(rec)["moves"]
    assert moves_checked >= 10
    assert st["nb_games_finished"] >= games - sum(active)
    for r in refs:
        r.close()
    eng.close()


def test_random_evaluation_symmetries():
    """SelfplayConfig::use_symmetries: every leaf is evaluated through a random board symmetry (NNEvaluator.cpp:134-146, 244-286). The
    root's priors after the first step must equal the NNEvaluator drop-in (agb_evaluate) for one of the 8 symmetries, all 8 must occur
    over the games, and the stream is keyed by (seed, global game id): a shard of the games reproduces the same trees."""
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob
    size, games, blocks, filters = 15, 64, 2, 64
    blob = netblob.pack(netblob.random_tensors(size, size, blocks, filters, True, seed=9), size, size, blocks, filters, True)
    rng = np.random.default_rng(123)
    boards, stm = _openings(rng, size, games)

    def make(n_games, first):
        eng = agb.Engine(agb.GameConfig(agb.GameRules.STANDARD, size, size), max_boards=n_games * 4, blocks=blocks, filters=filters, q_head=True,
                         games=n_games, max_batch_size=4, max_simulations=60, use_symmetries=True, seed=77, first_game_id=first)
        eng.load_weights(blob)
        eng.selfplay_reset(boards[first:first + n_games], stm[first:first + n_games])
        return eng

    eng = make(games, 0)
    eng.step(1)
    used = set()
    for g in range(games):
        _, priors, _, _, visits = eng.get_root(g)
        assert visits == 1
        empty = boards[g] == 0
        match = None
        for k in range(8):
            policy, _, _ = eng.evaluate(boards[g:g + 1], stm[g:g + 1], symmetry=[k])
            expect = np.where(empty, policy[0], 0.0)
            expect = expect / expect.sum()
            if np.abs(expect - priors).max() < 2e-6:
                match = k
                break
        assert match is not None, g
        used.add(match)
    assert len(used) == 8
    eng.step(40)
    shard = make(16, 32)
    shard.step(41)
    for g in range(16):
        a, b = eng.get_root(32 + g), shard.get_root(g)
        assert (a[0] == b[0]).all() and (a[1].view(np.uint32) == b[1].view(np.uint32)).all() and a[4] == b[4], g
    assert eng.stats()["overflow_flags"] == 0
    eng.close()
    shard.close()


@pytest.mark.parametrize("games", [96, 50])
def test_scheduling_of_solver_and_network_is_invisible(games):
    """AgbConfig::pipeline_groups / solver_sms only decide where and when K5 and K4 run (groups of games on their own streams, the solver on
    its own SMs beside the network kernel): roots, finished-game records and counters must not depend on them."""
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob, dataset
    size, blocks, filters = 15, 2, 64  # 50 games: groups of unequal size
    blob = netblob.pack(netblob.random_tensors(size, size, blocks, filters, False, seed=21), size, size, blocks, filters, False)
    rng = np.random.default_rng(321)
    boards, stm = _openings(rng, size, games)
    results = []
    for kwargs in (dict(pipeline_groups=1), dict(pipeline_groups=2), dict(pipeline_groups=3, solver_sms=8), dict(pipeline_groups=2, solver_sms=-1),
                   dict(pipeline_groups=4, solver_sms=146)):
        eng = agb.Engine(agb.GameConfig(agb.GameRules.RENJU, size, size, 40), max_boards=games * 4, blocks=blocks, filters=filters, games=games,
                         max_batch_size=4, max_simulations=50, solver_max_positions=60, use_symmetries=True, seed=3, **kwargs)
        eng.load_weights(blob)
        eng.selfplay_reset(boards, stm)
        eng.step(60)
        eng.step(45)
        records, n = eng.pop_finished()
        roots = [eng.get_root(g) for g in range(games)]
        st = eng.stats()
        assert st["overflow_flags"] == 0
        records = sorted(dataset.split_records(records, n))  # games of different groups finish in any order
        results.append((records, n, roots, {k: st[k] for k in ("nb_network_evaluations", "nb_node_count", "nb_proven_states", "nb_moves_played", "nb_games_finished")}))
        eng.close()
    base = results[0]
    assert (base[1] >= 1 or games < 96) and base[3]["nb_proven_states"] > 0 and base[3]["nb_moves_played"] > games
    for other in results[1:]:
        assert other[0] == base[0] and other[1] == base[1] and other[3] == base[3]
        for a, b in zip(base[2], other[2]):
            assert (a[0] == b[0]).all() and (a[1].view(np.uint32) == b[1].view(np.uint32)).all() and a[4] == b[4]


def test_save_and_resume_games_in_flight():
    """GeneratorManager::saveState / loadState: positions, move lists and recorded samples survive; trees are rebuilt. A resumed engine
    writes the same blob back, continues, and its finished records start with the samples recorded before the save."""
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob, dataset
    size, games, blocks, filters = 15, 12, 2, 64
    blob = netblob.pack(netblob.random_tensors(size, size, blocks, filters, False, seed=4), size, size, blocks, filters, False)
    rng = np.random.default_rng(5)
    boards, stm = _openings(rng, size, games)

    def make():
        eng = agb.Engine(agb.GameConfig(agb.GameRules.FREESTYLE, size, size, 20), max_boards=games * 4, blocks=blocks, filters=filters, games=games,
                         max_batch_size=4, max_simulations=50, solver_max_positions=20)
        eng.load_weights(blob)
        return eng

    a = make()
    a.selfplay_reset(boards, stm)
    a.step(70)
    a.pop_finished()
    saved = a.save_games()
    positions = [a.get_board(g) for g in range(games)]
    b = make()
    b.selfplay_reset(boards, stm)
    b.load_games(saved)
    assert b.save_games() == saved
    for g in range(games):
        board, to_move, _ = b.get_board(g)
        assert (board == positions[g][0]).all() and to_move == positions[g][1]
        assert b.get_root(g)[4] == 0  # empty tree
    with pytest.raises(agb.AgbError):
        b.load_games(saved[:len(saved) // 2])
    b.step(250)
    records, n = b.pop_finished()
    assert n >= 1 and b.stats()["overflow_flags"] == 0
    for rec in dataset.split_records(records, n):
        assert len(dataset.parse_record(rec)["moves"]) >= 1
    a.close()
    b.close()


@pytest.mark.parametrize("noise_type", ["custom", "dirichlet", "gumbel"])
def test_root_noise(noise_type):
    """EdgeSelectorConfig::noise_type / noise_weight (EdgeSelector.cpp:602-623, 1127-1137; random.cpp:89-123): the root's priors are
    mixed with noise drawn once per search. The random stream differs from the reference's thread-local mt19937 by design (keyed by seed
    and global game id), so the checks are on what the mixture must look like, on determinism and on shard invariance."""
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob
    size, games, blocks, filters, w = 15, 32, 2, 64, 0.25
    blob = netblob.pack(netblob.random_tensors(size, size, blocks, filters, False, seed=2), size, size, blocks, filters, False)
    rng = np.random.default_rng(8)
    boards, stm = _openings(rng, size, games)

    def make(n_games, first, weight=w):
        eng = agb.Engine(agb.GameConfig(agb.GameRules.STANDARD, size, size), max_boards=n_games * 4, blocks=blocks, filters=filters, games=n_games,
                         max_batch_size=4, max_simulations=80, noise_type=noise_type, noise_weight=weight, seed=31, first_game_id=first)
        eng.load_weights(blob)
        eng.selfplay_reset(boards[first:first + n_games], stm[first:first + n_games])
        return eng

    eng = make(games, 0)
    eng.step(3)  # the root is expanded in step 1; step 2 selects at the root for the first time and draws the noise
    mixtures, top5, sizeable = [], [], []
    for g in range(games):
        _, priors, _, _, visits = eng.get_root(g)
        noisy = eng.get_root_noise(g)
        legal = boards[g] == 0
        assert visits > 1 and (noisy[~legal] == 0).all() and (noisy[legal] >= 0).all()
        if noise_type == "gumbel":
            assert abs(float(noisy.sum()) - 1.0) < 1e-4
        else:
            noise = (noisy - (1.0 - w) * priors) / w  # what was mixed in
            assert (noise[legal] > -1e-6).all()
            total = float(noise[legal].sum())
            assert (abs(total - 1.0) < 1e-3) if noise_type == "dirichlet" else (0.3 < total <= 1.0 + 1e-4), total
            if noise_type == "dirichlet":
                top5.append(float(np.sort(noise[legal])[-5:].sum()))
                sizeable.append(int((noise[legal] > 0.01).sum()))
        assert np.abs(noisy - priors).max() > 1e-4
        mixtures.append(noisy)
    assert len({m.tobytes() for m in mixtures}) == games  # every game has its own draw
    if noise_type == "dirichlet":  # Dirichlet(0.05) over ~215 moves (numpy: top-5 share median 0.55, ~19 components above 0.01)
        assert 0.42 < np.median(top5) < 0.68 and 12 <= np.median(sizeable) <= 27, (np.median(top5), np.median(sizeable))
    eng.step(30)
    shard = make(8, 16)
    shard.step(33)
    for g in range(8):
        a, b = eng.get_root(16 + g), shard.get_root(g)
        assert (a[0] == b[0]).all() and a[4] == b[4], g
        assert (eng.get_root_noise(16 + g) == shard.get_root_noise(g)).all()
    eng.close()
    shard.close()


# ---- the configuration bench.py times: 20 x 128 network, batch 8, 400 simulations, solver 100, evaluation symmetries on ------------------
class _OwnThread:
    """Runs every call of one reference game on its own OS thread: the reference's random generators are thread_local (utils/random.cpp:17-23),
    so each game's evaluator then draws its symmetries from a fresh mt19937(0) of its own, independent of how the games are interleaved."""

    def __init__(self):
        from concurrent.futures import ThreadPoolExecutor
        self.pool = ThreadPoolExecutor(max_workers=1)

    def __call__(self, fn, *args):
        return self.pool.submit(fn, *args).result()

    def close(self):
        self.pool.shutdown()


def _reference_symmetry_stream(n):
    """The first n symmetries a reference NNEvaluator draws on a fresh thread: randInt(8) of mt19937(0) (NNEvaluator.cpp:134-139)."""
    import ctypes
    import refapi
    lib = ctypes.CDLL(refapi.REF_LIB)
    out = np.zeros(n, np.int32)
    lib.agref_rand_ints(8, n, refapi._p(out))
    assert ((np.random.RandomState(0).randint(0, 2 ** 32, n, dtype=np.uint64) >> 29) == out).all()  # it IS std::mt19937(0) >> 29
    return out.astype(np.int8)


def _benched_engine(games, table_entries, sym_table, **kw):
    import alphagomoku_b200 as agb
    from alphagomoku_b200 import netblob
    size, blocks, filters = 15, 20, 128
    eng = agb.Engine(agb.GameConfig(agb.GameRules.FREESTYLE, size, size), max_boards=games * 8, blocks=blocks, filters=filters, games=games, max_batch_size=8,
                     max_simulations=400, init_to="parent", max_nodes_per_game=1536, max_edges_per_game=1536 * 200, solver_max_positions=100,
                     solver_table_entries=table_entries, use_symmetries=True, seed=1234, **kw)
    eng.load_weights(netblob.pack(netblob.random_tensors(size, size, blocks, filters, False), size, size, blocks, filters, False))
    eng.set_symmetry_table(sym_table)
    return eng


def test_benched_configuration_matches_reference(ref):
    """Eight full-length freestyle games at the bench's settings (ResNet 20 x 128, 8 leaves per step, 400 simulations, solver at 100 positions
    with the reference's 4 Mi-entry table, evaluation symmetries ON and replayed from the reference's own generator) against the reference's
    Tree / Search / AlphaBetaSearch / NNEvaluator, step by step: boards, visit counts, priors, edge and root values to the bit, and the records."""
    import refapi
    from alphagomoku_b200 import dataset
    import bench
    games, size = 8, 15
    sym = _reference_symmetry_stream(1 << 17)
    eng = _benched_engine(games, 4 * 1024 * 1024, sym)

    def evaluate(features):
        return eng.forward(features)

    bench.select_workload("freestyle15")
    boards, stm = bench.random_openings(np.random.default_rng(5), games)
    threads = [_OwnThread() for _ in range(games)]
    refs = []
    for g in range(games):
        r = threads[g](lambda: refapi.RefSelfplay(0, size, evaluate, max_batch_size=8, max_simulations=400, init_to="parent", use_solver=True, solver_max_positions=100))
        threads[g](r.lib.agref_sp_use_symmetries, r.h, 1)
        refs.append(r)
    eng.set_solver_keys(np.stack([r.solver_keys() for r in refs]))
    eng.selfplay_reset(boards, stm)
    for g in range(games):
        threads[g](refs[g].set_position, boards[g], stm[g])
    active, plies, ref_records = [True] * games, 0, []
    for step in range(4000):
        eng.step(1)
        for g in range(games):
            if not active[g]:
                continue
            status = threads[g](refs[g].step)
            if status == 2:
                active[g] = False
                ref_records.append(refs[g].record())
                plies += 1
                continue
            rb, rstm, _, _ = refs[g].board()
            db, dstm, _ = eng.get_board(g)
            assert (db == rb).all() and dstm == rstm, (step, g)
            rv, rp, rq, rval, rn = refs[g].root()
            dv, dp, dq, dval, dn = eng.get_root(g)
            assert dn == rn and (dv == rv).all(), (step, g, dn, rn)
            assert (dp.view(np.uint32) == rp.view(np.uint32)).all(), (step, g)
            assert (dq.view(np.uint32) == rq.view(np.uint32)).all(), (step, g)
            assert (dval.view(np.uint32) == rval.view(np.uint32)).all(), (step, g)
            plies += status
        if not any(active):
            break
    assert not any(active), "a game did not finish"
    assert eng.stats()["overflow_flags"] == 0
    blob, n = eng.pop_finished()
    device_records = dataset.split_records(blob, n)
    for rec in ref_records:
        assert rec in device_records
    print(f"benched configuration: {games} games, {plies} plies, {step + 1} lockstep steps identical to the reference")
    for r, t in zip(refs, threads):
        r.close()
        t.close()
    eng.close()


def test_small_solver_table_is_sound_and_rarely_changes_moves():
    """The bench runs 65 536-entry solver tables per game (the reference: 4 Mi). Same games, same seeds, both sizes side by side: as long as two
    copies of a game are in the same position, nothing one table proves may contradict the other (a win is never a loss or a draw elsewhere),
    and the share of plies where the played move differs is reported."""
    games = 8
    import bench
    bench.select_workload("freestyle15")
    sym = _reference_symmetry_stream(1 << 17)
    big, small = _benched_engine(games, 4 * 1024 * 1024, sym), _benched_engine(games, 65536, sym)
    boards, stm = bench.random_openings(np.random.default_rng(5), games)
    keys = np.random.default_rng(7).integers(0, 2 ** 63, (games, 2 * 225, 2), dtype=np.int64).astype(np.uint64)
    for eng in (big, small):
        eng.set_solver_keys(keys)
        eng.selfplay_reset(boards, stm)
    same = [True] * games
    plies, differing, proven_compared, contradictions = 0, 0, 0, []
    stones = [int((boards[g] != 0).sum()) for g in range(games)]
    for step in range(3000):
        big.step(1)
        small.step(1)
        for g in range(games):
            if not same[g]:
                continue
            b1, s1, n1 = big.get_board(g)
            b2, s2, n2 = small.get_board(g)
            if n1 != stones[g] or n2 != stones[g]:  # a move was played (or the game ended and restarted) in at least one copy
                if n1 == stones[g] or n2 == stones[g]:
                    continue  # only one copy has moved so far (its root was proven a step earlier): wait for the other
                if n1 < stones[g] or n2 < stones[g]:
                    same[g] = False  # the game is over in at least one copy: stop following it
                    differing += int(n1 >= stones[g] or n2 >= stones[g])
                    plies += 1
                elif (b1 == b2).all():
                    plies += 1
                    stones[g] = n1
                else:
                    differing += 1
                    plies += 1
                    same[g] = False
                continue
            e1, r1 = big.get_root_scores(g)
            e2, r2 = small.get_root_scores(g)
            for a, b in list(zip(e1.tolist(), e2.tolist())) + [(r1, r2)]:
                pa, pb = (a >> 13) & 3, (b >> 13) & 3  # ProvenValue: 0 loss, 1 draw, 2 unknown, 3 win (Score.hpp)
                if pa != 2 and pb != 2 and a not in (0, 0xFFFF) and b not in (0, 0xFFFF):
                    proven_compared += 1
                    if pa != pb:
                        contradictions.append((step, g, a, b))
        if not any(same):
            break
    print(f"small solver table: {plies} plies followed, {differing} played differently ({100.0 * differing / max(plies, 1):.1f} %), "
          f"{proven_compared} proven scores compared, {len(contradictions)} contradictions")
    assert not contradictions, contradictions[:5]
    assert plies >= 40 and proven_compared > 0
    assert big.stats()["overflow_flags"] == 0 and small.stats()["overflow_flags"] == 0
    big.close()
    small.close()
