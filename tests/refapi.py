"""TEST INFRASTRUCTURE: ctypes view of oracle/_ref/libagref.so (the reference's own classes behind oracle/ref_shim.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libagref.so")
REF_LIB_FAST = os.path.join(ROOT, "oracle", "_ref", "libagref_fast.so")


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def available():
    return os.path.exists(REF_LIB)


class RefOracle:
    def __init__(self, fast=False):
        path = REF_LIB_FAST if fast else REF_LIB
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = ctypes.CDLL(path)
        self.lib.agref_calc_create.restype = ctypes.c_void_p
        self.lib.agref_calc_create.argtypes = [ctypes.c_int] * 3
        for name in ("agref_calc_destroy", "agref_calc_set_board", "agref_calc_add_move", "agref_calc_undo_move", "agref_calc_dump",
                     "agref_calc_histogram", "agref_calc_encode"):
            getattr(self.lib, name).argtypes = None
        self.lib.agref_defensive_moves.restype = ctypes.c_uint16
        self.lib.agref_open3_promotion_moves.restype = ctypes.c_uint16
        self._calcs = {}

    def tables(self, rules, with_update_mask=False):
        pt = np.zeros(1 << 20, np.uint8)
        ho = np.zeros(1 << 20, np.uint8)
        um = np.zeros((1 << 20, 2), np.uint32) if with_update_mask else None
        th = np.zeros((4096, 2), np.uint8)
        self.lib.agref_dump_tables(rules, _p(pt), _p(ho), _p(um), _p(th))
        return pt, ho, th, um

    def calc(self, rules, size):
        key = (rules, size)
        if key not in self._calcs:
            self._calcs[key] = ctypes.c_void_p(self.lib.agref_calc_create(rules, size, size))
        return self._calcs[key]

    def dump(self, rules, size, hist_only=False):
        h = self.calc(rules, size)
        c = size * size
        out = {"pattern_types": np.zeros((c, 4), np.uint8), "threats": np.zeros((c, 2), np.uint8), "legal": np.zeros(c, np.uint8),
               "forbidden": np.zeros(c, np.uint8), "features": np.zeros(c, np.uint32), "hist_counts": np.zeros((2, 10), np.int32),
               "hist_cells": np.zeros((2, 10, c), np.uint16)}
        # threat lists first: PatternCalculator::isForbidden (used by the forbidden dump and by encode in renju) runs
        # addMove/undoMove internally and thereby reorders the lists (swap-with-last removal, ThreatHistogram.hpp:61-66)
        for colour in (0, 1):
            self.lib.agref_calc_histogram(h, colour + 1, _p(out["hist_counts"][colour]), _p(out["hist_cells"][colour]), c)
        if hist_only:
            return out
        self.lib.agref_calc_dump(h, _p(out["pattern_types"]), _p(out["threats"]), _p(out["legal"]), _p(out["forbidden"]), None)
        self.lib.agref_calc_encode(h, _p(out["features"]))
        return out

    def set_board(self, rules, size, board, stm):
        board = np.ascontiguousarray(board, np.int8)
        self.lib.agref_calc_set_board(self.calc(rules, size), _p(board), int(stm))
        return self.dump(rules, size)

    def add_move(self, rules, size, row, col, sign):
        self.lib.agref_calc_add_move(self.calc(rules, size), int(row), int(col), int(sign))

    def undo_move(self, rules, size, row, col, sign):
        self.lib.agref_calc_undo_move(self.calc(rules, size), int(row), int(col), int(sign))

    def augment(self, features, size, mode):
        f = np.ascontiguousarray(features, np.uint32).copy()
        self.lib.agref_augment(_p(f), size, size, int(mode))
        return f

    def outcome(self, rules, size, board, row, col, sign, draw_after):
        board = np.ascontiguousarray(board, np.int8)
        return self.lib.agref_get_outcome(rules, size, size, _p(board), int(row), int(col), int(sign), int(draw_after))

    def is_forbidden(self, size, board, row, col, sign):
        board = np.ascontiguousarray(board, np.int8)
        return bool(self.lib.agref_is_forbidden(size, size, _p(board), int(row), int(col), int(sign)))


EVAL_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                           ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
                           ctypes.POINTER(ctypes.c_float))


class RefSelfplay:
    """One game driven by the reference's own Tree / Search / NNEvaluator (oracle/ref_shim_search.cpp). `evaluate` is a
    Python callable: features uint32 [n, cells] -> (policy [n, cells], value [n, 3], q [n, cells, 3] or None)."""

    def __init__(self, rules, size, evaluate, max_batch_size=8, max_simulations=100, init_to="parent", exploration_constant=1.25,
                 information_leak_threshold=0.01, use_solver=False, solver_max_positions=100, draw_after=0, fast=False, max_children=0,
                 policy_expansion_threshold=1.0e-4, final_selector="max_visit", final_exploration_constant=1.25,
                 policy_temperature=1.0):
        # fast=True: the reference's Release flags (-O3 -DNDEBUG), for timing only (its RNG is then seeded from the clock)
        self.lib = ctypes.CDLL(REF_LIB_FAST if fast and os.path.exists(REF_LIB_FAST) else REF_LIB)
        self.size, self.cells = size, size * size
        self.evaluations = 0

        def callback(ctx, features, batch, rows, cols, policy, value, action_values, moves_left):
            f = np.ctypeslib.as_array(features, shape=(batch, rows * cols)).copy()
            p, v, q = evaluate(f)
            self.evaluations += batch
            np.ctypeslib.as_array(policy, shape=(batch, rows * cols))[:] = p
            np.ctypeslib.as_array(value, shape=(batch, 3))[:] = v
            av = np.ctypeslib.as_array(action_values, shape=(batch, rows * cols, 3))
            av[:] = 0.0 if q is None else q
            np.ctypeslib.as_array(moves_left, shape=(batch,))[:] = 0.0

        self._cb = EVAL_FN(callback)
        self.lib.agref_sp_create.restype = ctypes.c_void_p
        self.lib.agref_sp_create.argtypes = [ctypes.c_int] * 6 + [ctypes.c_char_p, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                             EVAL_FN, ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_char_p, ctypes.c_float, ctypes.c_float]
        self.h = ctypes.c_void_p(self.lib.agref_sp_create(rules, size, size, draw_after, max_batch_size, max_simulations, init_to.encode(),
                                                          exploration_constant, information_leak_threshold, int(use_solver),
                                                          solver_max_positions, self._cb, None, max_children, policy_expansion_threshold,
                                                          final_selector.encode(), final_exploration_constant, policy_temperature))

    def set_position(self, board, stm):
        board = np.ascontiguousarray(board, np.int8)
        self.lib.agref_sp_set_position(self.h, _p(board), int(stm))

    def step(self):
        return self.lib.agref_sp_step(self.h)

    def root(self):
        visits = np.zeros(self.cells, np.int32)
        priors = np.zeros(self.cells, np.float32)
        q = np.zeros(self.cells, np.float32)
        value = np.zeros(3, np.float32)
        rv, ne = ctypes.c_int32(0), ctypes.c_int32(0)
        self.lib.agref_sp_root(self.h, _p(visits), _p(priors), _p(q), _p(value), ctypes.byref(rv), ctypes.byref(ne))
        return visits, priors, q, value, rv.value

    def board(self):
        b = np.zeros(self.cells, np.int8)
        stm, outcome, last = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int32(0)
        self.lib.agref_sp_board(self.h, _p(b), ctypes.byref(stm), ctypes.byref(outcome), ctypes.byref(last))
        return b, stm.value, outcome.value, last.value

    def stats(self):
        out = np.zeros(6, np.uint64)
        self.lib.agref_sp_stats(self.h, _p(out))
        return dict(zip(["nb_network_evaluations", "nb_node_count", "nb_duplicate_nodes", "nb_information_leaks", "nb_proven_states",
                         "nb_wasted_expansions"], out.tolist()))

    def solver_keys(self):
        """Zobrist words of this instance's AlphaBetaSearch table: uint64 [2 * cells, 2] (low, high)."""
        keys = np.zeros((2 * self.cells, 2), np.uint64)
        self.lib.agref_sp_solver_keys(self.h, _p(keys))
        return keys

    def record(self):
        """GameDataStorage::serialize of the game (complete after step() returned 2)."""
        self.lib.agref_sp_record.restype = ctypes.c_size_t
        buf = np.zeros(1 << 20, np.uint8)
        n = self.lib.agref_sp_record(self.h, _p(buf), ctypes.c_size_t(buf.size))
        return bytes(buf[:n])

    def close(self):
        if self.h:
            self.lib.agref_sp_destroy(self.h)
            self.h = None


def run_generator_threads(rules, size, evaluate, seconds, threads=1, games_per_thread=8, max_batch_size=8, evaluator_batch=64, max_simulations=400,
                          solver_max_positions=100, use_opening=True, use_symmetries=True, init_to="parent", exploration_constant=1.25,
                          start_boards=None, fast=True):
    """The reference's own GeneratorManager / GeneratorThread::run loop (oracle/ref_shim_manager.cpp) for `seconds`; `evaluate` as in
    RefSelfplay. Returns a dict: network evaluations (SearchStats), evaluator batches, finished games, seconds run, evaluator positions."""
    lib = ctypes.CDLL(REF_LIB_FAST if fast and os.path.exists(REF_LIB_FAST) else REF_LIB)

    def callback(ctx, features, batch, rows, cols, policy, value, action_values, moves_left):
        f = np.ctypeslib.as_array(features, shape=(batch, rows * cols)).copy()
        p, v, q = evaluate(f)
        np.ctypeslib.as_array(policy, shape=(batch, rows * cols))[:] = p
        np.ctypeslib.as_array(value, shape=(batch, 3))[:] = v
        av = np.ctypeslib.as_array(action_values, shape=(batch, rows * cols, 3))
        av[:] = 0.0 if q is None else q
        np.ctypeslib.as_array(moves_left, shape=(batch,))[:] = 0.0

    cb = EVAL_FN(callback)
    lib.agref_manager_run.argtypes = [ctypes.c_int] * 11 + [ctypes.c_char_p, ctypes.c_float, ctypes.c_double, EVAL_FN, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p]
    out = np.zeros(8, np.float64)
    sb = None
    if start_boards is not None:
        sb = np.ascontiguousarray(start_boards, np.int8).reshape(threads * games_per_thread, size * size)
    rc = lib.agref_manager_run(rules, size, size, threads, games_per_thread, max_batch_size, evaluator_batch, max_simulations, solver_max_positions,
                               int(use_opening), int(use_symmetries), init_to.encode(), exploration_constant, float(seconds), cb, None, _p(sb), _p(out))
    if rc != 0:
        raise RuntimeError(f"agref_manager_run failed ({rc})")
    return {"nb_network_evaluations": out[0], "evaluator_batches": out[1], "games_finished": out[2], "seconds": out[3], "evaluator_positions": out[4]}
