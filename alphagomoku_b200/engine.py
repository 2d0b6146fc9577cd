"""Host-side mirror of the reference interfaces on top of the C ABI.

Names follow the reference: GameRules / Sign (include/alphagomoku/game/rules.hpp:18-35, game/Move.hpp:17-23),
GameConfig (utils/configs.hpp:23-44); Engine methods are named after the PatternCalculator / NNInputFeatures /
AGNetwork / GameGenerator members they stand in for, batched over many positions."""
import ctypes
import enum
from dataclasses import dataclass

import numpy as np

from . import _lib


class GameRules(enum.IntEnum):
    FREESTYLE = 0
    STANDARD = 1
    RENJU = 2
    CARO5 = 3
    CARO6 = 4


class Sign(enum.IntEnum):
    NONE = 0
    CROSS = 1
    CIRCLE = 2
    ILLEGAL = 3


@dataclass
class GameConfig:
    rules: GameRules = GameRules.FREESTYLE
    rows: int = 15
    cols: int = 15
    draw_after: int = 0  # 0 -> rows * cols (configs.hpp:32)


class AgbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"agb200 error {code}: {message}")
        self.code = code


def move_to_short(row, col, sign):
    """Move::toShort (game/Move.hpp:144-147)."""
    return int(sign) | (int(row) << 2) | (int(col) << 9)


def short_to_move(s):
    return (s >> 2) & 127, (s >> 9) & 127, s & 3


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def config_from_json(text):
    """The reference's config.json text -> AgbConfig (agb_config_from_json; src/utils/configs.cpp:33-306)."""
    cfg = _lib.AgbConfig()
    err = ctypes.create_string_buffer(512)
    rc = _lib.load().agb_config_from_json(text.encode(), ctypes.byref(cfg), err, len(err))
    if rc != 0:
        raise AgbError(rc, err.value.decode())
    return cfg


class Engine:
    """One engine per GPU (one GeneratorThread per DeviceConfig in the reference)."""

    def __init__(self, game: GameConfig, max_boards=1024, device=0, blocks=0, filters=0, q_head=False, games=0, max_batch_size=1,
                 max_simulations=400, max_nodes_per_game=0, max_edges_per_game=0, init_to="parent", exploration_constant=1.25,
                 information_leak_threshold=0.01, policy_expansion_threshold=1.0e-4, max_children=0, solver_max_positions=0,
                 use_symmetries=False, seed=0, first_game_id=0, solver_table_entries=0, pipeline_groups=0, solver_sms=0, final_selector="max_visit",
                 final_exploration_constant=1.25, noise_type="none", noise_weight=0.0, policy_temperature=1.0):
        self._lib = _lib.load()
        self.game = game
        self.cells = game.rows * game.cols
        cfg = _lib.AgbConfig()
        cfg.rules, cfg.rows, cfg.cols, cfg.draw_after = int(game.rules), game.rows, game.cols, game.draw_after
        cfg.device, cfg.max_boards = device, max_boards
        cfg.blocks, cfg.filters, cfg.q_head = blocks, filters, int(q_head)
        cfg.games, cfg.max_batch_size, cfg.max_simulations = games, max_batch_size, max_simulations
        cfg.max_nodes_per_game, cfg.max_edges_per_game = max_nodes_per_game, max_edges_per_game
        cfg.init_to = {"loss": 0, "parent": 1, "draw": 2, "q_head": 3}[init_to]
        cfg.exploration_constant = exploration_constant
        cfg.information_leak_threshold = information_leak_threshold
        cfg.policy_expansion_threshold = policy_expansion_threshold
        cfg.max_children, cfg.solver_max_positions = max_children, solver_max_positions
        cfg.use_symmetries, cfg.seed, cfg.first_game_id = int(use_symmetries), seed, first_game_id
        cfg.solver_table_entries = solver_table_entries
        cfg.pipeline_groups = pipeline_groups
        cfg.solver_sms = solver_sms
        cfg.final_selector = {"max_visit": 0, "best": 1, "max_value": 2, "max_policy": 3, "min_visit": 4, "lcb": 5}[final_selector]
        cfg.final_exploration_constant = final_exploration_constant
        cfg.noise_type = {"none": 0, "custom": 1, "dirichlet": 2, "gumbel": 3}[noise_type]
        cfg.noise_weight = noise_weight
        cfg.policy_temperature = -1.0 if policy_temperature == 0 else policy_temperature  # 0 in the C struct means "default"
        self.config = cfg
        self.max_boards = max_boards
        handle = ctypes.c_void_p()
        rc = self._lib.agb_create(ctypes.byref(cfg), ctypes.byref(handle))
        if rc != 0:
            raise AgbError(rc, self._lib.agb_last_error(None).decode())
        self._h = handle

    def close(self):
        if getattr(self, "_h", None):
            self._lib.agb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise AgbError(rc, self._lib.agb_last_error(self._h).decode())

    # ---- tables ----------------------------------------------------------------------------------------------
    def get_tables(self):
        pt = np.zeros(1 << 20, np.uint8)
        ho = np.zeros(1 << 20, np.uint8)
        th = np.zeros((4096, 2), np.uint8)
        self._check(self._lib.agb_get_tables(self._h, _ptr(pt), _ptr(ho), _ptr(th)))
        return pt, ho, th

    # ---- PatternCalculator / NNInputFeatures -------------------------------------------------------------------
    def set_boards(self, boards, sign_to_move):
        """PatternCalculator::setBoard + NNInputFeatures::encode for n positions -> uint32 [n, cells]."""
        boards = np.ascontiguousarray(boards, np.int8).reshape(-1, self.cells)
        stm = np.ascontiguousarray(sign_to_move, np.int8).reshape(-1)
        n = boards.shape[0]
        assert stm.shape[0] == n
        features = np.zeros((n, self.cells), np.uint32)
        self._check(self._lib.agb_set_boards(self._h, _ptr(boards), _ptr(stm), n, _ptr(features)))
        return features

    def add_moves(self, moves):
        moves = np.ascontiguousarray(moves, np.uint16)
        self._check(self._lib.agb_add_moves(self._h, _ptr(moves), moves.shape[0]))

    def undo_moves(self, moves):
        moves = np.ascontiguousarray(moves, np.uint16)
        self._check(self._lib.agb_undo_moves(self._h, _ptr(moves), moves.shape[0]))

    def encode(self, n):
        features = np.zeros((n, self.cells), np.uint32)
        self._check(self._lib.agb_encode(self._h, n, _ptr(features)))
        return features

    def get_state(self, n, histograms=True):
        c = self.cells
        out = {
            "pattern_types": np.zeros((n, c, 4), np.uint8), "threats": np.zeros((n, c, 2), np.uint8),
            "legal": np.zeros((n, c), np.uint8), "forbidden": np.zeros((n, c), np.uint8),
            "hist_counts": np.zeros((n, 2, 10), np.int32) if histograms else None,
            "hist_cells": np.zeros((n, 2, 10, c), np.uint16) if histograms else None,
        }
        self._check(self._lib.agb_get_state(self._h, n, _ptr(out["pattern_types"]), _ptr(out["threats"]), _ptr(out["legal"]),
                                            _ptr(out["forbidden"]), _ptr(out["hist_counts"]), _ptr(out["hist_cells"])))
        return out

    def augment(self, features, symmetry):
        features = np.ascontiguousarray(features, np.uint32).reshape(-1, self.cells).copy()
        sym = np.ascontiguousarray(symmetry, np.int8).reshape(-1)
        self._check(self._lib.agb_augment(self._h, _ptr(features), _ptr(sym), features.shape[0]))
        return features

    def get_outcomes(self, boards, last_moves):
        boards = np.ascontiguousarray(boards, np.int8).reshape(-1, self.cells)
        moves = np.ascontiguousarray(last_moves, np.uint16).reshape(-1)
        out = np.zeros(boards.shape[0], np.int8)
        self._check(self._lib.agb_get_outcomes(self._h, _ptr(boards), _ptr(moves), boards.shape[0], _ptr(out)))
        return out

    # ---- AGNetwork -----------------------------------------------------------------------------------------------
    def weights_size(self):
        return self._lib.agb_weights_size(self._h)

    def load_weights(self, blob):
        blob = np.ascontiguousarray(blob)
        self._check(self._lib.agb_load_weights(self._h, _ptr(blob), blob.nbytes))

    def forward(self, features, want_q=False):
        features = np.ascontiguousarray(features, np.uint32).reshape(-1, self.cells)
        n = features.shape[0]
        policy = np.zeros((n, self.cells), np.float32)
        value = np.zeros((n, 3), np.float32)
        q = np.zeros((n, self.cells, 3), np.float32) if want_q else None
        self._check(self._lib.agb_forward(self._h, _ptr(features), n, _ptr(policy), _ptr(value), _ptr(q)))
        return policy, value, q

    def evaluate(self, boards, sign_to_move, symmetry=None, want_q=False):
        """NNEvaluator::evaluateGraph for n positions given as host boards: pack (K1+K3) -> forward (K4) -> unpack."""
        boards = np.ascontiguousarray(boards, np.int8).reshape(-1, self.cells)
        stm = np.ascontiguousarray(sign_to_move, np.int8).reshape(-1)
        sym = None if symmetry is None else np.ascontiguousarray(symmetry, np.int8).reshape(-1)
        n = boards.shape[0]
        policy = np.zeros((n, self.cells), np.float32)
        value = np.zeros((n, 3), np.float32)
        q = np.zeros((n, self.cells, 3), np.float32) if want_q else None
        self._check(self._lib.agb_evaluate(self._h, _ptr(boards), _ptr(stm), _ptr(sym), n, _ptr(policy), _ptr(value), _ptr(q)))
        return policy, value, q

    def evaluate_features(self, features, symmetry=None, want_q=False):
        """NNEvaluator::evaluateGraph for tasks that carry their (already augmented) feature words: forward (K4) -> inverse symmetry."""
        features = np.ascontiguousarray(features, np.uint32).reshape(-1, self.cells)
        sym = None if symmetry is None else np.ascontiguousarray(symmetry, np.int8).reshape(-1)
        n = features.shape[0]
        policy = np.zeros((n, self.cells), np.float32)
        value = np.zeros((n, 3), np.float32)
        q = np.zeros((n, self.cells, 3), np.float32) if want_q else None
        self._check(self._lib.agb_evaluate_features(self._h, _ptr(features), _ptr(sym), n, _ptr(policy), _ptr(value), _ptr(q)))
        return policy, value, q

    # ---- openings (OpeningGenerator) ------------------------------------------------------------------------------------
    def seed_openings(self, seed):
        self._check(self._lib.agb_seed_openings(self._h, seed))

    def prepare_opening(self, min_moves=1):
        """prepareOpening(config, min_moves): list of Move::toShort words."""
        moves = np.zeros(self.cells, np.uint16)
        n = ctypes.c_int32()
        self._check(self._lib.agb_prepare_opening(self._h, min_moves, _ptr(moves), ctypes.byref(n)))
        return moves[:n.value].copy()

    def generate_openings(self, count):
        """OpeningGenerator::generate: `count` unproven, balanced openings -> boards int8 [count, cells], sign_to_move int8 [count]."""
        boards, stm = np.zeros((count, self.cells), np.int8), np.zeros(count, np.int8)
        self._check(self._lib.agb_generate_openings(self._h, count, _ptr(boards), _ptr(stm)))
        return boards, stm

    # ---- lockstep self-play ------------------------------------------------------------------------------------------
    def selfplay_reset(self, boards=None, sign_to_move=None):
        b = None if boards is None else np.ascontiguousarray(boards, np.int8)
        s = None if sign_to_move is None else np.ascontiguousarray(sign_to_move, np.int8)
        self._check(self._lib.agb_selfplay_reset(self._h, _ptr(b), _ptr(s)))

    def solve(self, boards, sign_to_move, max_positions=100):
        """AlphaBetaSearch::solve on n positions (each from a cleared table). Returns scores uint16 [n], n_actions int32 [n],
        moves uint16 [n, cells], action_scores uint16 [n, cells], flags int32 [n] (bit 0 must-defend, bits 8.. positions visited)."""
        b = np.ascontiguousarray(boards, np.int8).reshape(-1, self.cells)
        s = np.ascontiguousarray(sign_to_move, np.int8)
        n = b.shape[0]
        scores, n_actions, flags = np.zeros(n, np.uint16), np.zeros(n, np.int32), np.zeros(n, np.int32)
        moves, action_scores = np.zeros((n, self.cells), np.uint16), np.zeros((n, self.cells), np.uint16)
        self._check(self._lib.agb_solve(self._h, _ptr(b), _ptr(s), n, max_positions, _ptr(scores), _ptr(n_actions), _ptr(moves), _ptr(action_scores),
                                        _ptr(flags)))
        return scores, n_actions, moves, action_scores, flags

    def think(self, boards, sign_to_move, active=None, max_steps=0):
        """Player::getMove for every active game: search boards[g] from a fresh tree, return (moves uint16 [games], root values [games, 2])."""
        b = np.ascontiguousarray(boards, np.int8).reshape(-1, self.cells)
        s = np.ascontiguousarray(sign_to_move, np.int8)
        a = np.ones(b.shape[0], np.int8) if active is None else np.ascontiguousarray(active, np.int8)
        moves, values = np.zeros(b.shape[0], np.uint16), np.zeros((b.shape[0], 2), np.float32)
        self._check(self._lib.agb_think(self._h, _ptr(b), _ptr(s), _ptr(a), _ptr(moves), _ptr(values), max_steps))
        return moves, values

    def save_games(self):
        """GeneratorManager::saveState: positions, move lists and samples of the games in flight, as bytes."""
        used = ctypes.c_size_t()
        self._lib.agb_save_games(self._h, None, 0, ctypes.byref(used))
        buf = np.zeros(used.value, np.uint8)
        self._check(self._lib.agb_save_games(self._h, _ptr(buf), buf.size, ctypes.byref(used)))
        return buf.tobytes()

    def load_games(self, blob):
        buf = np.frombuffer(blob, np.uint8)
        self._check(self._lib.agb_load_games(self._h, _ptr(buf), buf.size))

    def set_solver_keys(self, keys):
        """Zobrist words of the solver's transposition tables: uint64 [2 * cells, 2] (FastZobristHashing::m_keys)."""
        k = np.ascontiguousarray(keys, np.uint64).reshape(-1)
        self._check(self._lib.agb_set_solver_keys(self._h, _ptr(k), k.size))

    def set_symmetry_table(self, table):
        """Replay hook: the i-th network evaluation of every game uses symmetry table[i % len(table)] (empty: back to the random stream)."""
        t = np.ascontiguousarray(table, np.int8).reshape(-1)
        self._check(self._lib.agb_set_symmetry_table(self._h, _ptr(t) if t.size else None, int(t.size)))

    # ---- the trainer's batch loader (torch_api.h) ---------------------------------------------------------------------
    def load_dataset_fragment(self, index, path):
        self._check(self._lib.agb_dataset_load_fragment(self._h, index, path.encode()))

    def unload_dataset_fragment(self, index):
        self._check(self._lib.agb_dataset_unload_fragment(self._h, index))

    def dataset_size(self):
        """[(fragment, game, samples, symmetries)] for every loaded game (get_dataset_size)."""
        n = ctypes.c_int(0)
        self._check(self._lib.agb_dataset_size(self._h, ctypes.byref(n), None))
        sizes = np.zeros((n.value, 4), np.int32)
        self._check(self._lib.agb_dataset_size(self._h, ctypes.byref(n), _ptr(sizes)))
        return sizes

    def load_batch(self, samples):
        """load_batch: samples int32 [batch, 4] = (fragment, game, sample, augmentation) -> input [batch, rows, cols, 32], policy target
        [batch, rows, cols], value target [batch, 3], moves-left target [batch], action-value target [rows, cols, 3] (see agb200.h)."""
        s = np.ascontiguousarray(samples, np.int32).reshape(-1, 4)
        n, g = s.shape[0], self.game
        out = (np.zeros((n, g.rows, g.cols, 32), np.float32), np.zeros((n, g.rows, g.cols), np.float32), np.zeros((n, 3), np.float32),
               np.zeros(n, np.float32), np.zeros((g.rows, g.cols, 3), np.float32))
        self._check(self._lib.agb_load_batch(self._h, n, _ptr(s), *[_ptr(a) for a in out]))
        return out

    def step(self, n_steps=1):
        self._check(self._lib.agb_step(self._h, n_steps))

    def stats(self):
        st = _lib.AgbStats()
        self._check(self._lib.agb_get_stats(self._h, ctypes.byref(st)))
        return {name: getattr(st, name) for name, _ in st._fields_}

    def get_root(self, game):
        visits = np.zeros(self.cells, np.int32)
        priors = np.zeros(self.cells, np.float32)
        q = np.zeros(self.cells, np.float32)
        value = np.zeros(3, np.float32)
        rv = ctypes.c_int32(0)
        self._check(self._lib.agb_get_root(self._h, game, _ptr(visits), _ptr(priors), _ptr(q), _ptr(value), ctypes.byref(rv)))
        return visits, priors, q, value, rv.value

    def get_root_scores(self, game):
        """Score::to_short of every root edge (per cell) and of the root node."""
        scores = np.zeros(self.cells, np.uint16)
        root = ctypes.c_uint16(0)
        self._check(self._lib.agb_get_root_scores(self._h, game, _ptr(scores), ctypes.byref(root)))
        return scores, root.value

    def get_root_noise(self, game):
        """PUCTSelector::noisy_policy of game's current search, per cell (zeros before it is drawn)."""
        out = np.zeros(self.cells, np.float32)
        self._check(self._lib.agb_get_root_noise(self._h, game, _ptr(out)))
        return out

    def get_board(self, game):
        board = np.zeros(self.cells, np.int8)
        stm = ctypes.c_int8(0)
        mv = ctypes.c_int32(0)
        self._check(self._lib.agb_get_board(self._h, game, _ptr(board), ctypes.byref(stm), ctypes.byref(mv)))
        return board, stm.value, mv.value

    def pop_finished(self, capacity=1 << 24):
        buf = np.zeros(capacity, np.uint8)
        used = ctypes.c_size_t(0)
        n = ctypes.c_int(0)
        self._check(self._lib.agb_pop_finished(self._h, _ptr(buf), capacity, ctypes.byref(used), ctypes.byref(n)))
        return bytes(buf[:used.value]), n.value

    def synchronize(self):
        self._check(self._lib.agb_synchronize(self._h))

    def stream(self):
        return self._lib.agb_stream(self._h)
