/*
 * agb200.h -- C ABI of the B200-native lockstep self-play engine (libagb200.so).
 *
 * This is the drop-in boundary for AlphaGomoku's data-parallel hot path. The reference has no FFI for the
 * search (its seams are C++ classes), so every entry point below names the reference interface it replaces
 * (paths relative to the reference tree). Conventions follow the reference's only existing C ABI,
 * include/alphagomoku/dataset/torch_api.h:14-41: extern "C", plain pointers and sizes, caller-owned buffers,
 * no exceptions across the boundary. Every function returns 0 on success or a negative AGB_E* code;
 * agb_last_error() returns the text for the calling engine.
 *
 * Threading: thread-compatible, not thread-safe -- one host thread per engine, one engine per GPU
 * (mirrors one GeneratorThread per DeviceConfig, src/selfplay/GeneratorManager.cpp:29-53).
 *
 * Pointers named *_host are host memory (pinned or pageable); pointers named *_dev are device memory on the
 * engine's GPU. Board cells are int8: 0 empty, 1 cross/black, 2 circle/white (Sign, include/alphagomoku/game/Move.hpp:17-23).
 */
#ifndef AGB200_H_
#define AGB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

/* GameRules, include/alphagomoku/game/rules.hpp:18-25 */
enum
{
	AGB_FREESTYLE = 0,
	AGB_STANDARD = 1,
	AGB_RENJU = 2,
	AGB_CARO5 = 3,
	AGB_CARO6 = 4
};
/* GameOutcome, include/alphagomoku/game/rules.hpp:29-35 */
enum
{
	AGB_OUTCOME_UNKNOWN = 0,
	AGB_OUTCOME_DRAW = 1,
	AGB_OUTCOME_CROSS_WIN = 2,
	AGB_OUTCOME_CIRCLE_WIN = 3
};
enum
{
	AGB_OK = 0,
	AGB_EINVAL = -1, /* bad argument */
	AGB_ECUDA = -2, /* CUDA runtime error (text in agb_last_error) */
	AGB_ENOMEM = -3, /* capacity exceeded (boards, nodes, edges, records) */
	AGB_ESTATE = -4, /* call made in the wrong state (e.g. forward before weights) */
	AGB_EOVERFLOW = -5 /* a device-side bounded structure overflowed (reported, never silent) */
};

typedef struct AgbEngine AgbEngine;

/* root noise of the tree selector */
enum
{
	AGB_NOISE_NONE = 0, AGB_NOISE_CUSTOM = 1, AGB_NOISE_DIRICHLET = 2, AGB_NOISE_GUMBEL = 3
};

/* final move selectors of the self-play loop */
enum
{
	AGB_FINAL_MAX_VISIT = 0, AGB_FINAL_BEST = 1, AGB_FINAL_MAX_VALUE = 2, AGB_FINAL_MAX_POLICY = 3, AGB_FINAL_MIN_VISIT = 4, AGB_FINAL_LCB = 5
};

/* GameConfig + the parts of SelfplayConfig / SearchConfig the device engine honours
 * (include/alphagomoku/utils/configs.hpp:23-255). */
typedef struct AgbConfig
{
	int32_t rules; /* AGB_FREESTYLE .. AGB_CARO6 */
	int32_t rows, cols; /* 15x15 or 20x20 (square, <= 20: RawPatternCalculator.hpp:48) */
	int32_t draw_after; /* GameConfig::draw_after; <= 0 means rows*cols */
	int32_t device; /* CUDA device ordinal */
	int32_t max_boards; /* capacity of the pattern store = max concurrent positions per call */
	/* network shape (ResnetPV / ResnetPVQ, src/networks/networks.cpp:71-93, 143-168) */
	int32_t blocks, filters; /* residual blocks, channels (64 or 128) */
	int32_t q_head; /* 0 = "pv", 1 = "pvq" */
	/* search (SearchConfig, MCTSConfig, EdgeSelectorConfig, TreeConfig) */
	int32_t games; /* concurrent self-play games on this engine */
	int32_t max_batch_size; /* SearchConfig::max_batch_size: leaves selected per game per step */
	int32_t max_simulations; /* Constraints::max_simulations */
	int32_t max_nodes_per_game; /* node arena capacity per game */
	int32_t max_edges_per_game; /* edge arena capacity per game */
	int32_t init_to; /* 0 loss, 1 parent, 2 draw, 3 q_head (EdgeSelectorConfig::init_to) */
	float exploration_constant; /* EdgeSelectorConfig::exploration_constant (1.25) */
	float information_leak_threshold; /* TreeConfig (0.01) */
	float policy_expansion_threshold; /* MCTSConfig (1e-4) */
	int32_t max_children; /* MCTSConfig::max_children; <= 0 means unlimited */
	int32_t solver_max_positions; /* TSSConfig::max_positions: 0 = no solver, 1 = static move generator + evaluation, n = alpha-beta search of n positions */
	int32_t use_symmetries; /* SelfplayConfig::use_symmetries */
	uint64_t seed; /* base seed; per-game streams are keyed by (seed, global game id) */
	int32_t first_game_id; /* global id of this engine's game 0 (rank * games when sharded) */
	int32_t solver_table_entries; /* entries of each game's solver transposition table (power of two; 0 = 65536; the reference uses 4 Mi,
	                                 AlphaBetaSearch.cpp:55) */
	int32_t pipeline_groups; /* 1: all games advance together; 2..8: that many groups of games on their own streams, so that the solver and tree
	                            kernels of some groups overlap another group's network kernel; 0 = automatic (with the alpha-beta solver on and at
	                            least 1024 games: 6 when the SM partition is made of green contexts, else 3 from 3072 games on and 2 below;
	                            otherwise 1). Per-game results do not
	                            depend on it */
	int32_t final_selector; /* SelfplayConfig::final_selector.policy: AGB_FINAL_* (EdgeSelector::create, EdgeSelector.cpp:680-711) */
	float final_exploration_constant; /* its exploration_constant (used by AGB_FINAL_LCB) */
	int32_t noise_type; /* EdgeSelectorConfig::noise_type of the tree selector: AGB_NOISE_* (applied at the root, EdgeSelector.cpp:1127-1137) */
	float noise_weight; /* EdgeSelectorConfig::noise_weight; 0 = no noise */
	float policy_temperature; /* MCTSConfig::policy_temperature: 0 or 1 = priors as the network gives them (the default), t > 0 = prior^(1/t),
	                             negative = the reference's temperature 0 (one-hot on the best move) */
	int32_t solver_sms; /* with 2..8 pipeline groups and the alpha-beta solver: SMs the solver side (K5 and the small tree kernels) runs on while the
	                       network kernel takes the others (side by side, not sharing SMs: the solver is bound by instruction supply, the network by
	                       the tensor pipe, and on a shared SM both lose). The partition is a pair of CUDA green contexts (sizes in multiples of 8
	                       SMs); where the driver has none, SM-filling solver blocks. 0 = automatic, -1 = no partition. Ignored with one group */
} AgbConfig;

/* ---- lifetime ---------------------------------------------------------------------------------------------- */
int agb_create(const AgbConfig *config, AgbEngine **engine);
void agb_destroy(AgbEngine *engine);
const char* agb_last_error(const AgbEngine *engine); /* engine may be NULL: error of the last failed agb_create */
int agb_get_config(const AgbEngine *engine, AgbConfig *config);
/* The reference's config.json (MasterLearningConfig, src/utils/configs.cpp:33-306) -> AgbConfig: "game_config" (GameConfig), "generation_config"
 * (SelfplayConfig with its search_config: tree / mcts / tss, edge selector, constraints or "simulations", final_selector) and, when present,
 * "training_config" (network_arch ResnetPV / ResnetPVQ, blocks, filters). Same key names, required keys and defaults as the reference's
 * (const Json&) constructors, incl. init_to = "parent" when absent (configs.cpp:71). games = games_per_thread x device_config entries,
 * max_boards = games x max_batch_size; device, seed, first_game_id, the table size, groups and SM split stay 0 (= engine defaults) for the
 * caller to set. What the engine cannot do (selectors other than puct, time constraints, other architectures) is an error, not ignored.
 * Returns AGB_OK or AGB_EINVAL with a message in `error` (may be NULL). */
int agb_config_from_json(const char *json_text, AgbConfig *config, char *error, size_t error_size);
const char* agb_version(void);

/* ---- static tables (replaces PatternTable::get / ThreatTable::get, src/patterns/PatternTable.cpp:110-142,
 *      src/patterns/ThreatTable.cpp:135-167). Built on the device at agb_create; these read them back. --------- */
/* pattern_types[1<<20]: low nibble cross PatternType, high nibble circle PatternType, indexed by the 20-bit
 * narrowed window (PatternTable.hpp:135-138); half_open_3[1<<20]: bit0 cross, bit1 circle (PatternTable.hpp:118-127);
 * threats[4096*2]: (cross, circle) ThreatType per 12-bit pattern-type group (ThreatTable.hpp:79-100). Any may be NULL. */
int agb_get_tables(AgbEngine *engine, uint8_t *pattern_types_host, uint8_t *half_open_3_host, uint8_t *threats_host);

/* ---- K1 + K3: PatternCalculator::setBoard (src/patterns/PatternCalculator.cpp:40-67) followed by
 *      NNInputFeatures::encode (src/networks/NNInputFeatures.cpp:59-113) for n independent positions ------------- */
/* boards[n][rows*cols] int8, sign_to_move[n] int8 (1 or 2) -> features[n][rows*cols] uint32. The persistent
 * state of board slot i (lines, pattern types, threats, threat lists) is kept for agb_add_move / agb_get_state. */
int agb_set_boards(AgbEngine *engine, const int8_t *boards_host, const int8_t *sign_to_move_host, int n, uint32_t *features_host);
/* same, device pointers, asynchronous on the engine's stream (use agb_synchronize) */
int agb_set_boards_dev(AgbEngine *engine, const int8_t *boards_dev, const int8_t *sign_to_move_dev, int n, uint32_t *features_dev);

/* ---- K2: PatternCalculator::addMove / undoMove (PatternCalculator.cpp:68-106, 278-366), one move per board slot.
 *      moves[n]: Move::toShort wire form sign | row<<2 | col<<9 (Move.hpp:144-147); sign 0 = skip this slot. ----- */
int agb_add_moves(AgbEngine *engine, const uint16_t *moves_host, int n);
int agb_undo_moves(AgbEngine *engine, const uint16_t *moves_host, int n);
/* re-encode features from the persistent state (NNInputFeatures::encode on the incremental calculator) */
int agb_encode(AgbEngine *engine, int n, uint32_t *features_host);

/* read back the persistent state of the first n slots; any pointer may be NULL.
 * pattern_types[n][cells][4]: PatternEncoding byte per direction (H,V,D,A); threats[n][cells][2] (cross, circle);
 * legal[n][cells]; forbidden[n][cells] (PatternCalculator::isForbidden(CROSS,..), renju only, else 0);
 * hist_counts[n][2][10] and hist_cells[n][2][10][cells] uint16 (row | col<<8, Location::toShort) in list order
 * (ThreatHistogram.hpp:21-140). */
int agb_get_state(AgbEngine *engine, int n, uint8_t *pattern_types_host, uint8_t *threats_host, uint8_t *legal_host,
		uint8_t *forbidden_host, int32_t *hist_counts_host, uint16_t *hist_cells_host);

/* NNInputFeatures::augment (NNInputFeatures.cpp:114-154) on n feature planes, in place; symmetry[n] in -7..7 */
int agb_augment(AgbEngine *engine, uint32_t *features_host, const int8_t *symmetry_host, int n);

/* getOutcome (src/game/rules.cpp:110-133) for n (board, last move) pairs; outcomes[n] int8 AGB_OUTCOME_* */
int agb_get_outcomes(AgbEngine *engine, const int8_t *boards_host, const uint16_t *last_moves_host, int n, int8_t *outcomes_host);

/* ---- K4: AGNetwork (src/networks/AGNetwork.cpp:61-87) -- ResNet policy/value forward ---------------------------- */
/* weight blob layout: see DESIGN.md "network blob"; fp32, BN already folded (AGNetwork::optimize, AGNetwork.cpp:136-149) */
int agb_load_weights(AgbEngine *engine, const void *blob_host, size_t bytes);
size_t agb_weights_size(const AgbEngine *engine); /* expected blob size for the configured network */
/* features[n][cells] uint32 -> policy[n][cells] f32 (softmax), value[n][3] f32 (win, draw, loss; softmax),
 * q[n][cells][3] f32 or NULL (NetworkDataPack.cpp:112-129, 200-235) */
int agb_forward(AgbEngine *engine, const uint32_t *features_host, int n, float *policy_host, float *value_host, float *q_host);
int agb_forward_dev(AgbEngine *engine, const uint32_t *features_dev, int n, float *policy_dev, float *value_dev, float *q_dev);
/* NNEvaluator drop-in (src/search/monte_carlo/NNEvaluator.cpp:147-181 evaluateGraph = pack_to_network -> forward ->
 * unpack_from_network): host boards in, host policy/value(/q) out; K1 + K3 + K4 run back to back on the device.
 * symmetry[n] (0..7, or NULL for identity) is applied to the features before and inverted on policy/q after the network,
 * like NNEvaluator::pack_to_network / unpack_from_network (NNEvaluator.cpp:244-286). */
int agb_evaluate(AgbEngine *engine, const int8_t *boards_host, const int8_t *sign_to_move_host, const int8_t *symmetry_host, int n,
		float *policy_host, float *value_host, float *q_host);
/* The same for positions whose feature words the caller already holds -- NNEvaluator::pack_to_network's first branch (NNEvaluator.cpp:246-251:
 * a task the solver has processed carries its NNInputFeatures, which the reference augments IN PLACE and packs): features_host[n][cells] go to
 * the network as they are (already augmented by the caller), symmetry[n] (or NULL) is only inverted on policy / q afterwards. A task that sits
 * in the queue twice is augmented twice by the reference; taking its features from the caller reproduces that. */
int agb_evaluate_features(AgbEngine *engine, const uint32_t *features_host, const int8_t *symmetry_host, int n, float *policy_host, float *value_host,
		float *q_host);

/* ---- opening generation (OpeningGenerator, src/selfplay/OpeningGenerator.cpp:21-78; prepareOpening, src/utils/misc.cpp:142-170) -----
 * The generator's std::mt19937 starts from the low 32 bits of AgbConfig::seed; agb_seed_openings restarts it. */
int agb_seed_openings(AgbEngine *engine, uint32_t seed);
/* one random opening like prepareOpening(config, min_moves): moves[<= rows*cols] Move::toShort words, *n_moves their number */
int agb_prepare_opening(AgbEngine *engine, int min_moves, uint16_t *moves_host, int32_t *n_moves);
/* `count` balanced openings: random openings that the solver (1000 positions) cannot prove and whose network expectation is within
 * 0.1 + 0.01 * trials of 0.5 (OpeningGenerator::generate). boards[count][rows*cols] int8, sign_to_move[count]; feed them to
 * agb_selfplay_reset. Needs loaded weights. */
int agb_generate_openings(AgbEngine *engine, int count, int8_t *boards_host, int8_t *sign_to_move_host);

/* ---- lockstep self-play (GameGenerator::generate, src/selfplay/GameGenerator.cpp:46-121, over all games) ------- */
/* start (or restart) all games from given positions: boards[games][cells], sign_to_move[games]; NULL = empty boards, cross to move */
int agb_selfplay_reset(AgbEngine *engine, const int8_t *boards_host, const int8_t *sign_to_move_host);
/* ---- solver as a service (AlphaBetaSearch::solve, src/search/alpha_beta/AlphaBetaSearch.cpp:77-156) -------------------------------
 * Solves n caller-supplied positions, each from a cleared transposition table, with a budget of max_positions search nodes
 * (1 = static move generator + evaluation only). scores[n]: Score::to_short of the position score. Optional outputs (NULL to skip):
 * n_actions[n]; moves[n][cells] / action_scores[n][cells]: the root action list in its final order (what Search::solve leaves in
 * SearchTask::getEdges / getActionScores); flags[n]: bit 0 must-defend, bits 8.. positions visited. */
int agb_solve(AgbEngine *engine, const int8_t *boards_host, const int8_t *sign_to_move_host, int n, int max_positions, uint16_t *scores_host,
		int32_t *n_actions_host, uint16_t *moves_host, uint16_t *action_scores_host, int32_t *flags_host);
/* ---- one move per game on request (Player::setBoard / selectSolveEvaluate / expandBackup / isSearchOver / getMove,
 * src/evaluation/Player.cpp:93-238): every game with active[g] != 0 is handed boards[g] like Player::setBoard -- its search tree keeps every
 * node the new position can still reach (Tree::setBoard -> NodeCache::cleanup: the subtree under the moves played since its last search), the
 * solver's table enters a new generation -- then searches with the engine's settings and stops at its decision. agb_selfplay_reset before the
 * first call of a match = brand-new Player objects (empty trees). moves[games]: the chosen Move::toShort (0 for inactive games);
 * root_values[games][2] (optional): win and draw rate of the root. max_steps > 0 bounds the lockstep iterations. The building block of
 * evaluation games between two engines (alphagomoku_b200/arena.py). */
int agb_think(AgbEngine *engine, const int8_t *boards_host, const int8_t *sign_to_move_host, const int8_t *active_host, uint16_t *moves_host,
		float *root_values_host, int max_steps);

/* games in flight (GeneratorManager::saveState / loadState, src/selfplay/GeneratorManager.cpp:240-290): every game's position, move
 * list and the samples recorded so far, plus the per-game random streams (evaluation symmetries, root noise, opening choice) and the openings
 * pool, so that a resumed run continues like an uninterrupted one. Finished games must have been popped first (AGB_ESTATE otherwise). *used receives the size; AGB_ENOMEM when capacity is too small (call once with NULL to size the
 * buffer). Loading needs an engine with the same games / board / rules; the search trees start empty, as after the reference's
 * GameGenerator::load -> prepare_search. */
int agb_save_games(AgbEngine *engine, void *blob_host, size_t capacity, size_t *used);
int agb_load_games(AgbEngine *engine, const void *blob_host, size_t bytes);
/* replace the Zobrist words of the solver's transposition tables: keys[2 * rows * cols][2] = (low, high) 64-bit word of
 * (cell, CROSS) then (cell, CIRCLE), i.e. FastZobristHashing::m_keys (include/alphagomoku/search/ZobristHashing.hpp:111-127).
 * The words only decide which bucket a position maps to; parity tests pass the reference's own so that bucket replacement
 * (SharedHashTable::insert, SharedHashTable.hpp:160-181) sees the same collisions. n_words = 4 * rows * cols for one set shared by all
 * games, or games times that for one set per game (the reference has one per GameGenerator); agb_solve uses the first set. Call before
 * agb_selfplay_reset. */
int agb_set_solver_keys(AgbEngine *engine, const uint64_t *keys_host, size_t n_words);
/* replay hook for SelfplayConfig::use_symmetries: the i-th position a game sends to the network uses symmetry table[i mod n] instead of the
 * engine's random stream (n = 0 restores the stream). With table[i] = the i-th randInt(8) of a fresh reference thread this is exactly what
 * NNEvaluator::addToQueue draws (src/search/monte_carlo/NNEvaluator.cpp:134-139, src/utils/random.cpp:17-23) when each game has its own
 * evaluator thread, so a lockstep run can be compared with the reference move for move with symmetries on. Call before agb_selfplay_reset. */
int agb_set_symmetry_table(AgbEngine *engine, const int8_t *table_host, int n);
/* advance every game by n_steps lockstep iterations of select -> solve/encode -> evaluate -> expand -> backup (-> move) */
int agb_step(AgbEngine *engine, int n_steps);
/* pop finished-game records (GameDataStorage::serialize, src/dataset/GameDataStorage.cpp:217-251, format 201) */
int agb_pop_finished(AgbEngine *engine, void *records_host, size_t capacity, size_t *used, int *n_games);

/* ---- the trainer's batch loader (include/alphagomoku/dataset/torch_api.h:14-41, src/dataset/torch_api.cpp:130-281) ------------------------
 * GameDataBuffer files (what GeneratorManager::getGameBuffer().save writes, format 201) are loaded into numbered fragments; a batch is a list
 * of (fragment, game, sample, augmentation) like the reference's Sample_t. agb_load_batch fills the same five float arrays as load_batch:
 * input[batch][rows][cols][32] (bit c of NNInputFeatures::encode's word as channel c; computed by K1 + K3 on the device for the whole batch),
 * policy_target[batch][rows][cols] (visit counts, proven wins / losses overridden, normalised), value_target[batch][3] (win, draw, loss of the
 * side to move), moves_left_target[batch], action_values_target[rows][cols][3] (the reference does not advance this pointer between samples,
 * torch_api.cpp:251-280: one board's worth, written by every sample in turn; kept as it is). */
typedef struct AgbSample
{
	int32_t buffer_index, game_index, sample_index, augmentation; /* Sample_t, torch_api.h:16-22 */
} AgbSample;
int agb_dataset_load_fragment(AgbEngine *engine, int index, const char *path); /* load_dataset_fragment */
int agb_dataset_unload_fragment(AgbEngine *engine, int index); /* unload_dataset_fragment */
/* get_dataset_size: *n_games = games over all fragments; sizes[n_games][4] = fragment, game, samples, available symmetries (may be NULL) */
int agb_dataset_size(AgbEngine *engine, int *n_games, int32_t *sizes_host);
int agb_load_batch(AgbEngine *engine, int batch_size, const AgbSample *samples, float *input_host, float *policy_target_host, float *value_target_host,
		float *moves_left_target_host, float *action_values_target_host);

typedef struct AgbStats
{
	/* SearchStats / NNEvaluatorStats (Search.hpp:33-54, NNEvaluator.hpp:28-40) */
	uint64_t nb_network_evaluations;
	uint64_t nb_node_count;
	uint64_t nb_duplicate_nodes;
	uint64_t nb_information_leaks;
	uint64_t nb_proven_states;
	uint64_t nb_wasted_expansions;
	uint64_t nb_moves_played;
	uint64_t nb_games_finished;
	uint64_t nb_kernel_launches; /* kernels this engine launched since creation */
	uint64_t overflow_flags; /* non-zero: a bounded device structure overflowed */
	/* PerfEstimator-style device timing of the network kernel inside agb_step (networks/perf_stats.hpp:22-46) */
	uint64_t nn_kernel_ns; /* total CUDA-event time of the network kernel launches issued by agb_step */
	uint64_t nn_kernel_launches;
	uint64_t nn_positions; /* positions those launches evaluated */
	uint64_t solver_kernel_ns; /* device time during which at least one solver kernel (K5) launch issued by agb_step was queued or running (CUDA events;
	                              with one or two pipeline groups that is the sum of the launches, with more their intervals overlap) */
	uint64_t solver_sms; /* SMs the solver kernel currently runs on (AgbConfig::solver_sms; 0 = no partition). In automatic mode the engine
	                        re-balances it after every agb_step call from the measured K5 and K4 launch times */
	uint64_t pipeline_groups; /* groups of games the engine advances on their own streams (AgbConfig::pipeline_groups after defaults) */
} AgbStats;
int agb_get_stats(AgbEngine *engine, AgbStats *stats);

/* read-only view of game g's root: visits[cells] int32, priors[cells] f32, q[cells] f32 (expectation) -- for parity tests
 * against Tree::getInfo({}) (src/search/monte_carlo/Tree.cpp:394-416) */
int agb_get_root(AgbEngine *engine, int game, int32_t *visits_host, float *priors_host, float *q_host, float *root_value3_host,
		int32_t *root_visits);
/* proven scores of game g's root: edge_scores[cells] = Score::to_short of every root edge (Edge::getScore; the default Score for cells without an
 * edge), *root_score = the root node's (Node::getScore, include/alphagomoku/search/monte_carlo/Node.hpp). For checks of what the solver proved. */
int agb_get_root_scores(AgbEngine *engine, int game, uint16_t *edge_scores_host, uint16_t *root_score);
int agb_get_board(AgbEngine *engine, int game, int8_t *board_host, int8_t *sign_to_move, int32_t *move_number);
/* the root priors as the tree selector currently sees them (PUCTSelector::noisy_policy), noisy_policy[cells]; zeros until the search
 * of the current move has drawn its noise */
int agb_get_root_noise(AgbEngine *engine, int game, float *noisy_policy_host);

int agb_synchronize(AgbEngine *engine);
/* CUDA stream (cudaStream_t) the engine launches on, for callers that time with events */
void* agb_stream(AgbEngine *engine);

#ifdef __cplusplus
}
#endif

#endif /* AGB200_H_ */
