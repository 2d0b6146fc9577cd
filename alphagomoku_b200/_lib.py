"""ctypes binding of libagb200.so (include/agb200.h). Importing this module never falls back to a CPU path:
if the CUDA library is missing and cannot be built, it raises."""
import ctypes
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libagb200.so")


class AgbConfig(ctypes.Structure):
    _fields_ = [
        ("rules", ctypes.c_int32), ("rows", ctypes.c_int32), ("cols", ctypes.c_int32), ("draw_after", ctypes.c_int32),
        ("device", ctypes.c_int32), ("max_boards", ctypes.c_int32),
        ("blocks", ctypes.c_int32), ("filters", ctypes.c_int32), ("q_head", ctypes.c_int32),
        ("games", ctypes.c_int32), ("max_batch_size", ctypes.c_int32), ("max_simulations", ctypes.c_int32),
        ("max_nodes_per_game", ctypes.c_int32), ("max_edges_per_game", ctypes.c_int32), ("init_to", ctypes.c_int32),
        ("exploration_constant", ctypes.c_float), ("information_leak_threshold", ctypes.c_float),
        ("policy_expansion_threshold", ctypes.c_float), ("max_children", ctypes.c_int32),
        ("solver_max_positions", ctypes.c_int32), ("use_symmetries", ctypes.c_int32),
        ("seed", ctypes.c_uint64), ("first_game_id", ctypes.c_int32), ("solver_table_entries", ctypes.c_int32),
        ("pipeline_groups", ctypes.c_int32), ("final_selector", ctypes.c_int32), ("final_exploration_constant", ctypes.c_float),
        ("noise_type", ctypes.c_int32), ("noise_weight", ctypes.c_float), ("policy_temperature", ctypes.c_float),
        ("solver_sms", ctypes.c_int32),
    ]


class AgbStats(ctypes.Structure):
    _fields_ = [
        ("nb_network_evaluations", ctypes.c_uint64), ("nb_node_count", ctypes.c_uint64), ("nb_duplicate_nodes", ctypes.c_uint64),
        ("nb_information_leaks", ctypes.c_uint64), ("nb_proven_states", ctypes.c_uint64), ("nb_wasted_expansions", ctypes.c_uint64),
        ("nb_moves_played", ctypes.c_uint64), ("nb_games_finished", ctypes.c_uint64), ("nb_kernel_launches", ctypes.c_uint64),
        ("overflow_flags", ctypes.c_uint64), ("nn_kernel_ns", ctypes.c_uint64), ("nn_kernel_launches", ctypes.c_uint64),
        ("nn_positions", ctypes.c_uint64), ("solver_kernel_ns", ctypes.c_uint64), ("solver_sms", ctypes.c_uint64), ("pipeline_groups", ctypes.c_uint64),
    ]


# every symbol include/agb200.h declares: name -> (restype, argtypes)
_VP = ctypes.c_void_p
_I = ctypes.c_int
SYMBOLS = {
    "agb_create": (_I, [ctypes.POINTER(AgbConfig), ctypes.POINTER(_VP)]),
    "agb_destroy": (None, [_VP]),
    "agb_last_error": (ctypes.c_char_p, [_VP]),
    "agb_get_config": (_I, [_VP, ctypes.POINTER(AgbConfig)]),
    "agb_version": (ctypes.c_char_p, []),
    "agb_config_from_json": (_I, [ctypes.c_char_p, ctypes.POINTER(AgbConfig), ctypes.c_char_p, ctypes.c_size_t]),
    "agb_get_tables": (_I, [_VP, _VP, _VP, _VP]),
    "agb_set_boards": (_I, [_VP, _VP, _VP, _I, _VP]),
    "agb_set_boards_dev": (_I, [_VP, _VP, _VP, _I, _VP]),
    "agb_add_moves": (_I, [_VP, _VP, _I]),
    "agb_undo_moves": (_I, [_VP, _VP, _I]),
    "agb_encode": (_I, [_VP, _I, _VP]),
    "agb_get_state": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
    "agb_augment": (_I, [_VP, _VP, _VP, _I]),
    "agb_get_outcomes": (_I, [_VP, _VP, _VP, _I, _VP]),
    "agb_load_weights": (_I, [_VP, _VP, ctypes.c_size_t]),
    "agb_weights_size": (ctypes.c_size_t, [_VP]),
    "agb_forward": (_I, [_VP, _VP, _I, _VP, _VP, _VP]),
    "agb_forward_dev": (_I, [_VP, _VP, _I, _VP, _VP, _VP]),
    "agb_evaluate": (_I, [_VP, _VP, _VP, _VP, _I, _VP, _VP, _VP]),
    "agb_evaluate_features": (_I, [_VP, _VP, _VP, _I, _VP, _VP, _VP]),
    "agb_seed_openings": (_I, [_VP, ctypes.c_uint32]),
    "agb_prepare_opening": (_I, [_VP, _I, _VP, _VP]),
    "agb_generate_openings": (_I, [_VP, _I, _VP, _VP]),
    "agb_selfplay_reset": (_I, [_VP, _VP, _VP]),
    "agb_think": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I]),
    "agb_save_games": (_I, [_VP, _VP, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "agb_load_games": (_I, [_VP, _VP, ctypes.c_size_t]),
    "agb_set_solver_keys": (_I, [_VP, _VP, ctypes.c_size_t]),
    "agb_set_symmetry_table": (_I, [_VP, _VP, _I]),
    "agb_get_root_scores": (_I, [_VP, _I, _VP, _VP]),
    "agb_dataset_load_fragment": (_I, [_VP, _I, ctypes.c_char_p]),
    "agb_dataset_unload_fragment": (_I, [_VP, _I]),
    "agb_dataset_size": (_I, [_VP, ctypes.POINTER(_I), _VP]),
    "agb_load_batch": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
    "agb_solve": (_I, [_VP, _VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _VP]),
    "agb_step": (_I, [_VP, _I]),
    "agb_pop_finished": (_I, [_VP, _VP, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(_I)]),
    "agb_get_stats": (_I, [_VP, ctypes.POINTER(AgbStats)]),
    "agb_get_root": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP]),
    "agb_get_root_noise": (_I, [_VP, _I, _VP]),
    "agb_get_board": (_I, [_VP, _I, _VP, _VP, _VP]),
    "agb_synchronize": (_I, [_VP]),
    "agb_stream": (_VP, [_VP]),
}

_lib = None


def load():
    """Loads (building first if the sources are newer) the CUDA library. Raises if that is impossible."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or (os.path.exists("/usr/local/cuda/bin/nvcc") and _build.needs_build()):
            _build.build()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib
