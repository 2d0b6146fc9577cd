"""agb_config_from_json against the reference's own (const Json&) constructors (src/utils/configs.cpp:33-306) on the same config.json texts."""
import ctypes
import json

import numpy as np
import pytest


def _config(**overrides):
    cfg = {
        "description": "test", "data_type": "games",
        "game_config": {"rules": "FREESTYLE", "rows": 15, "cols": 15},
        "training_config": {"network_arch": "ResnetPV", "blocks": 10, "filters": 64},
        "generation_config": {
            "use_opening": True, "use_symmetries": True, "keep_loaded": False, "games_per_iteration": 1000, "games_per_thread": 8, "simulations": 400,
            "final_selector": {"policy": "max_visit"},
            "device_config": [{"device": "CPU", "batch_size": 64}, {"device": "CPU", "batch_size": 64}],
            "search_config": {"max_batch_size": 8, "tree_config": {}, "mcts_config": {"edge_selector_config": {"policy": "puct", "exploration_constant": 1.25}},
                              "tss_config": {"mode": 2, "max_positions": 100, "hash_table_size": 1048576}}}}
    for path, value in overrides.items():
        node = cfg
        keys = path.split("/")
        for k in keys[:-1]:
            node = node[k]
        if value is None:
            node.pop(keys[-1], None)
        else:
            node[keys[-1]] = value
    return cfg


CASES = [
    {},
    {"game_config/rules": "RENJU", "game_config/draw_after": 200},
    {"game_config/rules": "CARO5", "game_config/rows": 20, "game_config/cols": 20, "generation_config/simulations": 800},
    {"generation_config/search_config/mcts_config/edge_selector_config": {"policy": "puct", "init_to": "q_head", "noise_type": "dirichlet", "noise_weight": 0.25,
                                                                            "exploration_constant": 1.1}},
    {"generation_config/search_config/mcts_config/edge_selector_config": None},  # the struct default applies: init_to "q_head" (configs.hpp:78)
    {"generation_config/search_config/mcts_config": {"edge_selector_config": {"policy": "puct"}, "max_children": 32, "policy_expansion_threshold": 0.01,
                                                     "policy_temperature": 0.5}},
    {"generation_config/simulations": None, "generation_config/constraints": {"type": "simulations", "max_simulations": 250}},
    {"generation_config/final_selector": {"policy": "lcb", "exploration_constant": 0.7}},
    {"generation_config/search_config/tree_config": {"information_leak_threshold": 0.05}, "generation_config/search_config/tss_config": {"max_positions": 30}},
    {"generation_config/use_symmetries": False, "generation_config/games_per_thread": 16},
]


@pytest.mark.parametrize("overrides", CASES)
def test_config_from_json_matches_reference_parser(ref, overrides):
    import alphagomoku_b200 as agb
    import refapi
    text = json.dumps(_config(**overrides))
    ints, floats = np.zeros(16, np.int32), np.zeros(8, np.float32)
    ref.lib.agref_parse_config.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p]
    assert ref.lib.agref_parse_config(text.encode(), refapi._p(ints), refapi._p(floats)) == 0
    c = agb.engine.config_from_json(text)
    mine = [c.rules, c.rows, c.cols, c.draw_after, c.use_symmetries, c.games, c.max_simulations, c.max_batch_size,
            c.max_children if c.max_children > 0 else 2 ** 31 - 1, c.solver_max_positions, c.init_to, c.noise_type, c.final_selector]
    assert mine == ints[:13].tolist(), (mine, ints[:13].tolist())
    temperature = 0.0 if c.policy_temperature < 0 else c.policy_temperature
    mine_f = np.array([c.information_leak_threshold, c.exploration_constant, c.noise_weight, c.policy_expansion_threshold, temperature, c.final_exploration_constant], np.float32)
    assert (mine_f.view(np.uint32) == floats[:6].view(np.uint32)).all(), (mine_f, floats[:6])
    assert c.blocks == 10 and c.filters == 64 and c.q_head == 0 and c.max_boards == c.games * c.max_batch_size


@pytest.mark.parametrize("overrides,message", [
    ({"game_config/rows": None}, "rows"), ({"generation_config/use_symmetries": None}, "use_symmetries"),
    ({"generation_config/search_config/mcts_config/edge_selector_config": {"policy": "uct"}}, "puct"),
    ({"generation_config/simulations": None, "generation_config/constraints": {"type": "time", "time_for_turn": 5.0}}, "time"),
    ({"training_config/network_arch": "ConvNextPVQMraw"}, "network_arch"), ({"game_config/rules": "GO"}, "rules")])
def test_config_from_json_rejects_what_the_reference_or_the_engine_cannot_take(ref, overrides, message):
    import alphagomoku_b200 as agb
    with pytest.raises(agb.AgbError, match=message):
        agb.engine.config_from_json(json.dumps(_config(**overrides)))
    with pytest.raises(agb.AgbError, match="JSON"):
        agb.engine.config_from_json('{"game_config": {"rules": "FREESTYLE", ')
