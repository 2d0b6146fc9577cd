// TEST INFRASTRUCTURE -- C-ABI shim around the UNMODIFIED reference GeneratorManager / GeneratorThread / GameGenerator
// (src/selfplay/GeneratorManager.cpp, GameGenerator.cpp): the reference's own self-play loop, GeneratorThread::run
// (GeneratorManager.cpp:120-141), is started with GeneratorThread::start() and stopped with stop(), exactly as
// GeneratorManager::generate does (:177-218) minus its one-second polling loop. bench.py's reference arm times it.
//
// This translation unit is compiled with -fno-access-control (oracle/Makefile) because the reference keeps the pieces a
// timed run needs private: GeneratorManager::games_to_generate / network_loader / generators (set by generate(), which blocks
// until the games are played) and GameGenerator::game / state (to start the games from given positions the way
// GameGenerator::load does, GameGenerator.cpp:131-141). No reference code is modified or copied.
#include <alphagomoku/selfplay/GeneratorManager.hpp>
#include <alphagomoku/selfplay/GameGenerator.hpp>
#include <alphagomoku/selfplay/NetworkLoader.hpp>
#include <alphagomoku/networks/AGNetwork.hpp>
#include <alphagomoku/evaluation/Player.hpp>
#include <alphagomoku/search/monte_carlo/NNEvaluator.hpp>
#include <alphagomoku/utils/configs.hpp>

#include <minml/utils/json.hpp>

#include <chrono>
#include <climits>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

using namespace ag;

namespace agref
{
	extern GameConfig g_game_config;
	extern agref_eval_fn g_eval_fn;
	extern void *g_eval_ctx;
}

namespace
{
	// one evaluation-game player of the reference with its own evaluator (src/evaluation/Player.cpp; EvaluationGame.cpp:95-150 drives it)
	struct RefPlayer
	{
			NNEvaluator evaluator;
			Player player;
			RefPlayer(const GameConfig &gc, const SelfplayConfig &sc) :
					evaluator(sc.device_config.at(0)),
					player(gc, sc, evaluator, "player")
			{
				evaluator.useSymmetries(false);
				evaluator.loadGraph(NetworkLoader(""));
			}
	};
}

extern "C"
{
	// Player with the network replaced by a callback (the shadow AGNetwork reads the callback installed at loadGraph time)
	void* agref_player_create(int rules, int rows, int cols, int max_batch_size, int max_simulations, int solver_max_positions, const char *init_to,
			float exploration_constant, agref_eval_fn eval_fn, void *ctx)
	{
		GameConfig gc(static_cast<GameRules>(rules), rows, cols);
		agref::g_game_config = gc;
		agref::g_eval_fn = eval_fn;
		agref::g_eval_ctx = ctx;
		SelfplayConfig sc;
		sc.use_symmetries = false;
		sc.constraints = Constraints::simulations(max_simulations);
		sc.final_selector.policy = "max_visit";
		sc.device_config = { DeviceConfig() };
		sc.device_config[0].batch_size = max_batch_size;
		sc.search_config.max_batch_size = max_batch_size;
		sc.search_config.mcts_config.edge_selector_config.policy = "puct";
		sc.search_config.mcts_config.edge_selector_config.init_to = init_to;
		sc.search_config.mcts_config.edge_selector_config.exploration_constant = exploration_constant;
		sc.search_config.tss_config.max_positions = solver_max_positions;
		return new RefPlayer(gc, sc);
	}
	void agref_player_destroy(void *h)
	{
		delete static_cast<RefPlayer*>(h);
	}
	// the per-move loop of EvaluationGame::generate (EvaluationGame.cpp:95-150): setBoard, then select / solve / evaluate / expand / backup until
	// isSearchOver, then getMove. Returns Move::toShort; root visits per cell and the root's simulation count for the comparison.
	int agref_player_move(void *h, const int8_t *board, int sign_to_move, int32_t *root_visits_per_cell, int32_t *simulations)
	{
		RefPlayer *rp = static_cast<RefPlayer*>(h);
		const GameConfig gc = rp->player.game_config;
		matrix<Sign> b(gc.rows, gc.cols);
		for (int i = 0; i < b.size(); i++)
			b[i] = static_cast<Sign>(board[i]);
		rp->player.setBoard(b, static_cast<Sign>(sign_to_move));
		do
		{
			rp->player.selectSolveEvaluate();
			rp->evaluator.evaluateGraph();
			rp->player.expandBackup();
		} while (not rp->player.isSearchOver());
		const Node root = rp->player.tree.getInfo( { });
		for (int i = 0; i < b.size(); i++)
			root_visits_per_cell[i] = 0;
		for (Edge *edge = root.begin(); edge < root.end(); edge++)
			root_visits_per_cell[edge->getMove().row * gc.cols + edge->getMove().col] = edge->getVisits();
		*simulations = rp->player.tree.getSimulationCount();
		return rp->player.getMove().toShort();
	}
	// the solver's hash keys of this player (FastZobristHashing::m_keys), for the device table's bucket mapping
	void agref_player_solver_keys(void *h, uint64_t *keys)
	{
		RefPlayer *rp = static_cast<RefPlayer*>(h);
		const GameConfig gc = rp->player.game_config;
		for (int r = 0; r < gc.rows; r++)
			for (int c = 0; c < gc.cols; c++)
				for (int sgn = 1; sgn <= 2; sgn++)
				{ // like export_keys in ref_shim_search.cpp
					HashKey128 k;
					rp->player.search.getSolver().shared_table.getHashFunction().updateHash(k, Move(r, c, static_cast<Sign>(sgn)));
					keys[2 * (2 * (r * gc.cols + c) + sgn - 1) + 0] = k.getLow();
					keys[2 * (2 * (r * gc.cols + c) + sgn - 1) + 1] = k.getHigh();
				}
	}
	// the reference's own parse of a config.json text: GameConfig(json["game_config"]) and SelfplayConfig(json["generation_config"])
	// (src/utils/configs.cpp:44-50, 253-268), flattened for the comparison with agb_config_from_json. Returns 0, or -1 when the reference throws.
	int agref_parse_config(const char *json_text, int32_t *ints, float *floats)
	{
		try
		{
			const Json root = Json::load(json_text);
			const GameConfig gc(root["game_config"]);
			const SelfplayConfig sc(root["generation_config"]);
			const EdgeSelectorConfig &es = sc.search_config.mcts_config.edge_selector_config;
			const auto code = [](const std::string &v, std::initializer_list<const char*> names)
			{
				int i = 0;
				for (const char *n : names)
				{
					if (v == n)
						return i;
					i++;
				}
				return -1;
			};
			ints[0] = static_cast<int>(gc.rules);
			ints[1] = gc.rows;
			ints[2] = gc.cols;
			ints[3] = gc.draw_after;
			ints[4] = sc.use_symmetries ? 1 : 0;
			ints[5] = sc.games_per_thread * static_cast<int>(sc.device_config.size());
			ints[6] = sc.constraints.max_simulations;
			ints[7] = sc.search_config.max_batch_size;
			ints[8] = sc.search_config.mcts_config.max_children;
			ints[9] = sc.search_config.tss_config.max_positions;
			ints[10] = code(es.init_to, { "loss", "parent", "draw", "q_head" });
			ints[11] = code(es.noise_type, { "none", "custom", "dirichlet", "gumbel" });
			ints[12] = code(sc.final_selector.policy, { "max_visit", "best", "max_value", "max_policy", "min_visit", "lcb" });
			floats[0] = sc.search_config.tree_config.information_leak_threshold;
			floats[1] = es.exploration_constant;
			floats[2] = es.noise_weight;
			floats[3] = sc.search_config.mcts_config.policy_expansion_threshold;
			floats[4] = sc.search_config.mcts_config.policy_temperature;
			floats[5] = sc.final_selector.exploration_constant;
			return 0;
		}
		catch (std::exception &e)
		{
			return -1;
		}
	}

	// Runs `threads` GeneratorThreads (one NNEvaluator of batch `evaluator_batch` each, `games_per_thread` GameGenerators each) for `seconds`.
	// start_boards (optional): [threads * games_per_thread][rows * cols] int8 positions the games start from instead of their opening
	// generator (stones are replayed as alternating moves; games that finish continue the reference's normal way, with use_opening).
	// out[0] = network evaluations (SearchStats::nb_network_evaluations summed over the games), out[1] = NNEvaluator batches,
	// out[2] = games added to the buffer, out[3] = seconds the threads actually ran, out[4] = GameGenerator::generate calls are not counted
	// by the reference, so: nodes visited by the solver is not exposed either; out[4] = evaluated positions by NNEvaluatorStats (batch_sizes).
	int agref_manager_run(int rules, int rows, int cols, int threads, int games_per_thread, int max_batch_size, int evaluator_batch, int max_simulations,
			int solver_max_positions, int use_opening, int use_symmetries, const char *init_to, float exploration_constant, double seconds,
			agref_eval_fn eval_fn, void *ctx, const int8_t *start_boards, double *out)
	{
		try
		{
			GameConfig gc(static_cast<GameRules>(rules), rows, cols);
			agref::g_game_config = gc;
			agref::g_eval_fn = eval_fn;
			agref::g_eval_ctx = ctx;
			SelfplayConfig sc;
			sc.use_opening = use_opening != 0;
			sc.use_symmetries = use_symmetries != 0;
			sc.games_per_thread = games_per_thread;
			sc.constraints.max_simulations = max_simulations;
			sc.final_selector.policy = "max_visit";
			sc.device_config.assign(threads, DeviceConfig());
			for (DeviceConfig &dc : sc.device_config)
				dc.batch_size = evaluator_batch;
			sc.search_config.max_batch_size = max_batch_size;
			sc.search_config.mcts_config.edge_selector_config.policy = "puct";
			sc.search_config.mcts_config.edge_selector_config.init_to = init_to;
			sc.search_config.mcts_config.edge_selector_config.exploration_constant = exploration_constant;
			sc.search_config.tss_config.max_positions = solver_max_positions;

			GeneratorManager manager(gc, sc);
			manager.games_to_generate = INT_MAX; // what generate(loader, n) sets before starting the threads
			manager.network_loader = NetworkLoader("");
			if (start_boards != nullptr)
			{
				const int cells = rows * cols;
				int index = 0;
				for (auto &thread : manager.generators)
					for (auto &generator : thread->generators)
					{
						const int8_t *b = start_boards + static_cast<size_t>(index++) * cells;
						std::vector<Move> cross, circle, moves;
						for (int i = 0; i < cells; i++)
							if (b[i] == 1)
								cross.push_back(Move(i / cols, i % cols, Sign::CROSS));
							else if (b[i] == 2)
								circle.push_back(Move(i / cols, i % cols, Sign::CIRCLE));
						if (cross.size() != circle.size() and cross.size() != circle.size() + 1)
							return -2; // not a position of alternating play
						for (size_t i = 0; i < cross.size(); i++)
						{
							moves.push_back(cross[i]);
							if (i < circle.size())
								moves.push_back(circle[i]);
						}
						// GameGenerator::generate's GAME_NOT_STARTED branch followed by what it does with a popped opening (GameGenerator.cpp:48-78)
						generator->game.beginGame();
						generator->game_data_storage.clear();
						generator->tree.clear();
						generator->search.getSolver().clear();
						generator->game.loadOpening(moves);
						generator->state = GameGenerator::GAMEPLAY_SELECT_SOLVE_EVALUATE;
						generator->prepare_search();
					}
			}
			const auto t0 = std::chrono::steady_clock::now();
			for (auto &thread : manager.generators)
			{
				thread->clearStats();
				thread->start();
			}
			std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
			for (auto &thread : manager.generators)
				thread->stop();
			const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			double evals = 0.0, batches = 0.0, positions = 0.0;
			for (auto &thread : manager.generators)
			{
				evals += static_cast<double>(thread->getSearchStats().nb_network_evaluations);
				const NNEvaluatorStats es = thread->getEvaluatorStats();
				batches += static_cast<double>(es.compute.getTotalCount());
				positions += static_cast<double>(es.batch_sizes);
			}
			out[0] = evals;
			out[1] = batches;
			out[2] = static_cast<double>(manager.getGameBuffer().numberOfGames());
			out[3] = elapsed;
			out[4] = positions;
			return 0;
		}
		catch (std::exception &e)
		{
			std::fprintf(stderr, "agref_manager_run: %s\n", e.what());
			return -1;
		}
	}
}
