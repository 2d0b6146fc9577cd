"""Evaluation games between two engines (two networks) on the device: the arena of the reference
(src/evaluation/EvaluationManager.cpp, EvaluationGame.cpp, Player.cpp), reduced to its data-parallel core.

Every game has two players, each with its own engine (its own network, search settings and tree). Per ply, the engine whose colour is
to move searches all of its games at once (Engine.think = Player::setBoard ... getMove) and the host plays the chosen moves, asks the
device for the outcomes (getOutcome) and flips the side to move. Colours alternate between games like in EvaluationGame (each opening is
played twice with swapped colours when `swap_colours` is set). Like the reference's Player, a player keeps its search tree from move to move:
Engine.think re-roots it on the new position (Player::setBoard -> Tree::setBoard -> NodeCache::cleanup), move for move identical to the
reference's Player on the same network (tests/test_arena_gpu.py). There is no time control: the search budget is the engine's max_simulations.

Returns per-game records (moves, outcome, who played cross) and the score of engine A, ready for an Elo fit or a PGN dump."""
import numpy as np

OUTCOME_NAMES = {0: "UNKNOWN", 1: "DRAW", 2: "CROSS_WIN", 3: "CIRCLE_WIN"}


def play_match(engine_a, engine_b, openings, sign_to_move, swap_colours=True, max_plies=None, max_steps_per_move=4096):
    """openings: int8 [n, cells]; sign_to_move: int8 [n]. Both engines must have been created with `games` >= the number of games played
    (n, or 2 n with swap_colours) and the same GameConfig. Returns a dict with `games` (list of dicts) and `score_a` in [0, 1]."""
    openings = np.ascontiguousarray(openings, np.int8).reshape(len(openings), -1)
    stm0 = np.ascontiguousarray(sign_to_move, np.int8)
    if swap_colours:
        boards = np.concatenate([openings, openings])
        stm = np.concatenate([stm0, stm0])
        a_is_cross = np.concatenate([np.ones(len(openings), bool), np.zeros(len(openings), bool)])
    else:
        boards, stm = openings.copy(), stm0.copy()
        a_is_cross = (np.arange(len(openings)) % 2) == 0
    n, cells = boards.shape
    size = engine_a.game.rows
    assert engine_a.config.games >= n and engine_b.config.games >= n, "engines need at least as many games as the match has"

    def padded(engine, x, fill=0):
        out = np.full((engine.config.games,) + x.shape[1:], fill, x.dtype)
        out[:n] = x
        return out

    outcome = np.zeros(n, np.int8)
    moves = [[] for _ in range(n)]
    max_plies = max_plies or cells
    for engine in (engine_a, engine_b):  # new Player objects for this match: empty trees, cleared solver tables
        engine.selfplay_reset()
    for _ in range(max_plies):
        running = outcome == 0
        if not running.any():
            break
        for engine, mine in ((engine_a, a_is_cross == (stm == 1)), (engine_b, a_is_cross != (stm == 1))):
            active = running & mine
            if not active.any():
                continue
            chosen, _ = engine.think(padded(engine, boards), padded(engine, stm, 1), padded(engine, active.astype(np.int8)), max_steps=max_steps_per_move)
            idx = np.flatnonzero(active)
            for g in idx:
                mv = int(chosen[g])
                row, col = (mv >> 2) & 127, (mv >> 9) & 127
                assert (mv & 3) == stm[g] and boards[g, row * size + col] == 0, (g, mv)
                boards[g, row * size + col] = stm[g]
                moves[g].append(mv)
            last = np.zeros(n, np.uint16)
            last[idx] = chosen[idx]
            result = engine.get_outcomes(boards[idx], last[idx])
            outcome[idx] = result
            stm[idx] = 3 - stm[idx]
    outcome[outcome == 0] = 1  # unfinished after max_plies: a draw
    games = []
    score_a = 0.0
    for g in range(n):
        a_won = (outcome[g] == 2) == bool(a_is_cross[g]) and outcome[g] in (2, 3)
        score_a += 0.5 if outcome[g] == 1 else (1.0 if a_won else 0.0)
        games.append({"moves": moves[g], "outcome": OUTCOME_NAMES[int(outcome[g])], "a_plays_cross": bool(a_is_cross[g]), "opening": openings[g % len(openings)].copy()})
    return {"games": games, "score_a": score_a / n, "n_games": n}


def generate_pgn(game, rows, cols, rules, first_name="A", second_name="B"):
    """Game::generatePGN (src/game/Game.cpp): the tags EvaluationManager's pgn files carry, moves as Move::text without the sign."""
    cross, circle = (first_name, second_name) if game["a_plays_cross"] else (second_name, first_name)
    result = {"CROSS_WIN": "1-0", "CIRCLE_WIN": "0-1", "DRAW": "1/2-1/2"}.get(game["outcome"], "*")
    text = [f'[White "{cross}"]', f'[Black "{circle}"]', f'[Result "{result}"]', f'[Variant "{rules}:{rows}x{cols}"]', ""]
    plies = [chr(ord("a") + ((mv >> 9) & 127)) + str((mv >> 2) & 127) for mv in game["moves"]]
    text.append(" ".join(f"{i // 2 + 1}. {plies[i]}" + (f" {plies[i + 1]}" if i + 1 < len(plies) else "") for i in range(0, len(plies), 2)) + f" {result}")
    return "\n".join(text) + "\n"
