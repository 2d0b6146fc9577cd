// TEST INFRASTRUCTURE -- C-ABI shim that drives the UNMODIFIED reference search classes (Tree, Search, NNEvaluator,
// EdgeSelector, EdgeGenerator, AlphaBetaSearch) one self-play step at a time, with the network replaced by a callback
// (oracle/ref_shadow AGNetwork). The control flow below is the per-game state machine of GameGenerator::generate /
// make_move / prepare_search (src/selfplay/GameGenerator.cpp:46-185) made synchronous and started from a given position;
// `use_solver = 0` leaves out the Search::solve() call so that the tree kernels can be compared before the device solver
// exists (tasks then take the "not processed by solver" path of UnifiedGenerator, EdgeGenerator.cpp:269-303).
#include <array>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <thread>
#include <vector>
// the solver's hash keys are private to AlphaBetaSearch; tests need them to give the device table the same bucket mapping
#define private public
#include <alphagomoku/search/alpha_beta/AlphaBetaSearch.hpp>
#undef private
#include <alphagomoku/dataset/GameDataBuffer.hpp>
#include <alphagomoku/dataset/GameDataStorage.hpp>
#include <alphagomoku/dataset/data_packs.hpp>
#include <alphagomoku/game/Board.hpp>
#include <alphagomoku/game/rules.hpp>
#include <alphagomoku/networks/AGNetwork.hpp>
#include <alphagomoku/search/monte_carlo/EdgeGenerator.hpp>
#include <alphagomoku/search/monte_carlo/EdgeSelector.hpp>
#include <alphagomoku/search/monte_carlo/NNEvaluator.hpp>
#include <alphagomoku/search/monte_carlo/Search.hpp>
#include <alphagomoku/search/monte_carlo/Tree.hpp>
#include <alphagomoku/search/alpha_beta/AlphaBetaSearch.hpp>
#include <alphagomoku/selfplay/NetworkLoader.hpp>
#include <alphagomoku/utils/configs.hpp>
#include <alphagomoku/utils/misc.hpp>
#include <alphagomoku/utils/random.hpp>

#include <minml/utils/serialization.hpp>

#include <cstdio>
#include <cstring>
#include <memory>
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
	struct RefSelfplay
	{
			GameConfig game_config;
			SelfplayConfig selfplay_config;
			bool use_solver;
			matrix<Sign> board;
			Sign sign_to_move = Sign::CROSS;
			std::vector<Move> moves;
			Tree tree;
			Search search;
			NNEvaluator evaluator;
			GameOutcome outcome = GameOutcome::UNKNOWN;
			Move last_move;
			GameDataStorage game_data_storage; // default constructed like GameGenerator's (GameGenerator.hpp:39)

			RefSelfplay(const GameConfig &gc, const SelfplayConfig &sc, bool solver) :
					game_config(gc),
					selfplay_config(sc),
					use_solver(solver),
					board(gc.rows, gc.cols),
					tree(sc.search_config.tree_config),
					search(gc, sc.search_config),
					evaluator(sc.device_config.at(0))
			{
				search.setBatchSize(sc.search_config.max_batch_size);
				evaluator.useSymmetries(sc.use_symmetries);
				evaluator.loadGraph(NetworkLoader(""));
			}
			void prepare_search()
			{ // GameGenerator.cpp:174-185
				search.cleanup(tree);
				tree.setBoard(board, sign_to_move, false);
				search.setBoard(board, sign_to_move);
				const MCTSConfig &mcts_config = search.getConfig().mcts_config;
				std::unique_ptr<EdgeSelector> tmp = EdgeSelector::create(mcts_config.edge_selector_config);
				tree.setEdgeSelector(*tmp);
				tree.setEdgeGenerator(UnifiedGenerator(mcts_config.max_children, mcts_config.policy_expansion_threshold, mcts_config.policy_temperature, true));
			}
	};
}

extern "C"
{
	void* agref_sp_create(int rules, int rows, int cols, int draw_after, int max_batch_size, int max_simulations, const char *init_to,
			float exploration_constant, float information_leak_threshold, int use_solver, int solver_max_positions, agref_eval_fn eval_fn, void *ctx,
			int max_children, float policy_expansion_threshold, const char *final_policy, float final_exploration_constant,
			float policy_temperature)
	{
		GameConfig gc(static_cast<GameRules>(rules), rows, cols);
		if (draw_after > 0)
			gc.draw_after = draw_after;
		SelfplayConfig sc;
		sc.use_opening = false;
		sc.use_symmetries = false;
		sc.constraints.max_simulations = max_simulations;
		sc.final_selector.policy = (final_policy != nullptr and final_policy[0] != 0) ? final_policy : "max_visit";
		sc.final_selector.exploration_constant = final_exploration_constant;
		sc.search_config.mcts_config.policy_temperature = policy_temperature;
		sc.device_config = { DeviceConfig() };
		sc.device_config[0].batch_size = max_batch_size;
		sc.search_config.max_batch_size = max_batch_size;
		sc.search_config.tree_config.information_leak_threshold = information_leak_threshold;
		sc.search_config.mcts_config.edge_selector_config.policy = "puct";
		sc.search_config.mcts_config.edge_selector_config.init_to = init_to;
		sc.search_config.mcts_config.edge_selector_config.exploration_constant = exploration_constant;
		sc.search_config.tss_config.max_positions = solver_max_positions;
		if (max_children > 0)
		{ // MCTSConfig: prune_weak_moves keeps at most this many edges of an unproven position (EdgeGenerator.cpp:69-84)
			sc.search_config.mcts_config.max_children = max_children;
			sc.search_config.mcts_config.policy_expansion_threshold = policy_expansion_threshold;
		}
		agref::g_game_config = gc;
		agref::g_eval_fn = eval_fn;
		agref::g_eval_ctx = ctx;
		return new RefSelfplay(gc, sc, use_solver != 0);
	}
	// SelfplayConfig::use_symmetries for this instance's evaluator (GeneratorThread's constructor does the same, GeneratorManager.cpp:35)
	void agref_sp_use_symmetries(void *h, int on)
	{
		static_cast<RefSelfplay*>(h)->evaluator.useSymmetries(on != 0);
	}
	// the first n values of ag::randInt(r) on a fresh thread, i.e. from a newly constructed thread_local generator (src/utils/random.cpp:17-23):
	// what an NNEvaluator draws for its first n tasks when it is the only consumer on its thread (NNEvaluator.cpp:134-139)
	void agref_rand_ints(int r, int n, int32_t *out)
	{
		std::thread worker([=]()
		{
			for (int i = 0; i < n; i++)
				out[i] = ag::randInt(r);
		});
		worker.join();
	}
	void agref_sp_destroy(void *h)
	{
		delete static_cast<RefSelfplay*>(h);
	}
	void agref_sp_set_position(void *h, const int8_t *board, int sign_to_move)
	{ // GAME_NOT_STARTED -> GAMEPLAY (GameGenerator.cpp:48-61), starting from an arbitrary position
		RefSelfplay *sp = static_cast<RefSelfplay*>(h);
		for (int i = 0; i < sp->board.size(); i++)
			sp->board[i] = static_cast<Sign>(board[i]);
		sp->sign_to_move = static_cast<Sign>(sign_to_move);
		sp->moves.clear();
		{ // stones already on the board become the opening moves of the record: cross / circle in row-major order, interleaved
			const int cells = sp->board.size(), cols = sp->board.cols();
			int ci = 0, oi = 0;
			while (true)
			{
				while (ci < cells and sp->board[ci] != Sign::CROSS)
					ci++;
				while (oi < cells and sp->board[oi] != Sign::CIRCLE)
					oi++;
				if (ci >= cells and oi >= cells)
					break;
				if (ci < cells)
				{
					sp->moves.push_back(Move(ci / cols, ci % cols, Sign::CROSS));
					ci++;
				}
				if (oi < cells)
				{
					sp->moves.push_back(Move(oi / cols, oi % cols, Sign::CIRCLE));
					oi++;
				}
			}
		}
		sp->game_data_storage.clear();
		sp->outcome = GameOutcome::UNKNOWN;
		sp->tree.clear();
		sp->search.getSolver().clear();
		sp->prepare_search();
	}
	// one SELECT_SOLVE_EVALUATE + EXPAND_AND_BACKUP round; returns 0 = searching, 1 = a move was made, 2 = game over
	int agref_sp_step(void *h)
	{
		RefSelfplay *sp = static_cast<RefSelfplay*>(h);
		const int max_simulations = sp->selfplay_config.constraints.max_simulations;
		sp->search.select(sp->tree, max_simulations);
		if (sp->use_solver)
			sp->search.solve();
		sp->search.scheduleToNN(sp->evaluator);
		sp->evaluator.evaluateGraph();
		sp->search.generateEdges(sp->tree);
		sp->search.expand(sp->tree);
		sp->search.backup(sp->tree);

		const Value root_eval = sp->tree.getInfo( { }).getValue();
		const int simulations = get_simulations_for_move(root_eval.draw_rate, max_simulations, 50);
		if (sp->tree.getSimulationCount() > simulations or sp->tree.isRootProven())
		{ // make_move (GameGenerator.cpp:145-173)
			const Node root_node = sp->tree.getInfo( { });
			std::unique_ptr<EdgeSelector> selector = EdgeSelector::create(sp->selfplay_config.final_selector);
			const Move move = selector->select(&root_node)->getMove();
			SearchDataPack sample(root_node, sp->board);
			sp->game_data_storage.addSample(sample);
			Board::putMove(sp->board, move);
			sp->moves.push_back(move);
			sp->last_move = move;
			sp->sign_to_move = invertSign(move.sign);
			sp->outcome = getOutcome(sp->game_config.rules, sp->board, move, sp->game_config.draw_after);
			if (sp->outcome != GameOutcome::UNKNOWN)
			{ // GameGenerator.cpp:104-112
				sp->game_data_storage.setOutcome(sp->outcome);
				sp->game_data_storage.addMoves(sp->moves);
				return 2;
			}
			sp->prepare_search();
			return 1;
		}
		return 0;
	}
	void agref_sp_root(void *h, int32_t *visits, float *priors, float *q, float *value3, int32_t *root_visits, int32_t *n_edges)
	{
		RefSelfplay *sp = static_cast<RefSelfplay*>(h);
		const int cells = sp->board.size();
		std::memset(visits, 0, cells * sizeof(int32_t));
		std::memset(priors, 0, cells * sizeof(float));
		std::memset(q, 0, cells * sizeof(float));
		const Node root = sp->tree.getInfo( { });
		*root_visits = root.getVisits();
		*n_edges = root.numberOfEdges();
		value3[0] = root.getValue().win_rate;
		value3[1] = root.getValue().draw_rate;
		value3[2] = root.getValue().loss_rate();
		if (root.numberOfEdges() == 0)
			return;
		for (const Edge *edge = root.begin(); edge < root.end(); edge++)
		{
			const Move m = edge->getMove();
			const int idx = m.row * sp->board.cols() + m.col;
			visits[idx] = edge->getVisits();
			priors[idx] = edge->getPolicyPrior();
			q[idx] = edge->getExpectation();
		}
	}
	void agref_sp_board(void *h, int8_t *board, int32_t *sign_to_move, int32_t *outcome, int32_t *last_move)
	{
		RefSelfplay *sp = static_cast<RefSelfplay*>(h);
		for (int i = 0; i < sp->board.size(); i++)
			board[i] = static_cast<int8_t>(sp->board[i]);
		*sign_to_move = static_cast<int32_t>(sp->sign_to_move);
		*outcome = static_cast<int32_t>(sp->outcome);
		*last_move = sp->last_move.toShort();
	}
	void agref_sp_stats(void *h, uint64_t *out)
	{
		RefSelfplay *sp = static_cast<RefSelfplay*>(h);
		const SearchStats st = sp->search.getStats();
		out[0] = st.nb_network_evaluations;
		out[1] = st.nb_node_count;
		out[2] = st.nb_duplicate_nodes;
		out[3] = st.nb_information_leaks;
		out[4] = st.nb_proven_states;
		out[5] = st.nb_wasted_expansions;
	}
	// GameDataStorage::serialize of the game played so far (complete once agref_sp_step returned 2)
	size_t agref_sp_record(void *h, uint8_t *out, size_t capacity)
	{
		RefSelfplay *sp = static_cast<RefSelfplay*>(h);
		SerializedObject so;
		sp->game_data_storage.serialize(so);
		if (so.size() <= capacity)
			std::memcpy(out, so.data(), so.size());
		return so.size();
	}
	// SearchDataPack -> SearchDataStorage_v201::loadFrom -> serialize for one ply given as dense per-cell arrays
	size_t agref_serialize_sample_v201(int rows, int cols, const int8_t *board, const int32_t *visits, const float *prior, const float *win,
			const float *draw, const uint16_t *scores, uint16_t minimax_score, uint16_t flags, uint8_t *out, size_t capacity)
	{
		SearchDataPack pack(rows, cols);
		for (int i = 0; i < rows * cols; i++)
		{
			pack.board[i] = static_cast<Sign>(board[i]);
			pack.visit_count[i] = visits[i];
			pack.policy_prior[i] = prior[i];
			pack.action_values[i] = Value(win[i], draw[i]);
			pack.action_scores[i] = Score::from_short(scores[i]);
		}
		pack.minimax_score = Score::from_short(minimax_score);
		pack.flags = BitMask1D<uint16_t>(flags);
		SearchDataStorage_v201 storage;
		storage.loadFrom(pack);
		SerializedObject so;
		storage.serialize(so);
		if (so.size() <= capacity)
			std::memcpy(out, so.data(), so.size());
		return so.size();
	}
	// GameDataBuffer::load (GameDataBuffer.cpp:113-128) of a file written by our host writer; returns games, fills samples per game
	int agref_buffer_load(const char *path, int32_t *samples_per_game, int32_t *moves_per_game, int32_t *outcomes, int capacity, int32_t *rows,
			int32_t *cols, int32_t *rules)
	{
		GameDataBuffer buffer;
		try
		{
			buffer.load(path);
		}
		catch (const std::exception &ex)
		{
			std::fprintf(stderr, "agref_buffer_load: %s\n", ex.what());
			return -1;
		}
		*rows = buffer.getConfig().rows;
		*cols = buffer.getConfig().cols;
		*rules = static_cast<int32_t>(buffer.getConfig().rules);
		for (int i = 0; i < buffer.numberOfGames() and i < capacity; i++)
		{
			samples_per_game[i] = buffer.getGameData(i).numberOfSamples();
			moves_per_game[i] = buffer.getGameData(i).numberOfMoves();
			outcomes[i] = static_cast<int32_t>(buffer.getGameData(i).getOutcome());
		}
		return buffer.numberOfGames();
	}
	// ---- AlphaBetaSearch as a persistent object: transposition table and generation survive between calls -------------------------
	namespace
	{
		void export_keys(const AlphaBetaSearch &solver, int rows, int cols, uint64_t *keys)
		{ // [2 * cells][2]: low and high word of FastZobristHashing's key of (cell, CROSS) then (cell, CIRCLE)
			for (int r = 0; r < rows; r++)
				for (int c = 0; c < cols; c++)
					for (int s = 1; s <= 2; s++)
					{
						HashKey128 k;
						solver.shared_table.getHashFunction().updateHash(k, Move(r, c, static_cast<Sign>(s)));
						keys[2 * (2 * (r * cols + c) + s - 1) + 0] = k.getLow();
						keys[2 * (2 * (r * cols + c) + s - 1) + 1] = k.getHigh();
					}
		}
		int run_solver(AlphaBetaSearch &solver, const GameConfig &gc, const int8_t *board, int sign_to_move, int max_nodes, uint16_t *moves, uint16_t *scores,
				uint16_t *result_score, int32_t *flags)
		{
			matrix<Sign> b(gc.rows, gc.cols);
			for (int i = 0; i < gc.rows * gc.cols; i++)
				b[i] = static_cast<Sign>(board[i]);
			SearchTask task(gc);
			task.set(b, static_cast<Sign>(sign_to_move));
			solver.setDepthLimit(100); // Search::solve (Search.cpp:160-169)
			solver.setNodeLimit(max_nodes);
			solver.setTimeLimit(std::numeric_limits<double>::max());
			const int nodes = solver.solve(task);
			int n = 0;
			for (const Edge &e : task.getEdges())
			{
				const Move m = e.getMove();
				moves[n] = m.toShort();
				scores[n] = Score::to_short(task.getActionScores().at(m.row, m.col));
				n++;
			}
			*result_score = Score::to_short(task.getScore());
			*flags = static_cast<int>(task.mustDefend()) | (static_cast<int>(task.wasStaticallySolved()) << 1) | (static_cast<int>(task.wasRecursivelySolved()) << 2)
					| (nodes << 8);
			return n;
		}
		struct RefSolver
		{
				GameConfig gc;
				AlphaBetaSearch solver;
				RefSolver(const GameConfig &cfg) :
						gc(cfg),
						solver(cfg)
				{
				}
		};
	}
	void* agref_solver_create(int rules, int rows, int cols, int draw_after)
	{
		GameConfig gc(static_cast<GameRules>(rules), rows, cols);
		if (draw_after > 0)
			gc.draw_after = draw_after;
		return new RefSolver(gc);
	}
	void agref_solver_destroy(void *h)
	{
		delete static_cast<RefSolver*>(h);
	}
	void agref_solver_keys(void *h, uint64_t *keys)
	{
		RefSolver *s = static_cast<RefSolver*>(h);
		export_keys(s->solver, s->gc.rows, s->gc.cols, keys);
	}
	void agref_solver_next_generation(void *h)
	{
		static_cast<RefSolver*>(h)->solver.increaseGeneration();
	}
	void agref_solver_clear(void *h)
	{
		static_cast<RefSolver*>(h)->solver.clear();
	}
	// flags: bit0 must_defend, bit1 statically solved, bit2 recursively solved, bits 8.. node count
	int agref_solver_solve(void *h, const int8_t *board, int sign_to_move, int max_nodes, uint16_t *moves, uint16_t *scores, uint16_t *result_score,
			int32_t *flags)
	{
		RefSolver *s = static_cast<RefSolver*>(h);
		return run_solver(s->solver, s->gc, board, sign_to_move, max_nodes, moves, scores, result_score, flags);
	}
	void agref_sp_solver_keys(void *h, uint64_t *keys)
	{
		RefSelfplay *sp = static_cast<RefSelfplay*>(h);
		export_keys(sp->search.getSolver(), sp->game_config.rows, sp->game_config.cols, keys);
	}
	// AlphaBetaSearch::solve (AlphaBetaSearch.cpp:77-156) on one position with a node limit and a cleared table; outputs the task's edge
	// list (in order), their scores, the position score and flags (as above)
	int agref_solve(int rules, int rows, int cols, int draw_after, const int8_t *board, int sign_to_move, int max_nodes, uint16_t *moves,
			uint16_t *scores, uint16_t *result_score, int32_t *flags)
	{
		static std::unique_ptr<RefSolver> solver;
		GameConfig gc(static_cast<GameRules>(rules), rows, cols);
		if (draw_after > 0)
			gc.draw_after = draw_after;
		if (solver == nullptr or solver->gc.rules != gc.rules or solver->gc.rows != gc.rows or solver->gc.draw_after != gc.draw_after)
			solver = std::make_unique<RefSolver>(gc);
		solver->solver.clear();
		return run_solver(solver->solver, solver->gc, board, sign_to_move, max_nodes, moves, scores, result_score, flags);
	}
	// MoveGenerator::generate in a given MoveGeneratorMode on a fresh PatternCalculator, like MoveGenWrapper of the reference's
	// test/search/alpha_beta/test_move_generator.cpp:22-43. flags: bit0 must_defend, bit1 has_initiative; returns the list size
	int agref_generate(int rules, int rows, int cols, const int8_t *board, int sign_to_move, int mode, uint16_t *moves, uint16_t *scores, int32_t *flags)
	{
		GameConfig gc(static_cast<GameRules>(rules), rows, cols);
		matrix<Sign> b(rows, cols);
		for (int i = 0; i < rows * cols; i++)
			b[i] = static_cast<Sign>(board[i]);
		PatternCalculator calc(gc);
		MoveGenerator generator(gc, calc);
		ActionStack stack(1024 + rows * cols);
		calc.setBoard(b, static_cast<Sign>(sign_to_move));
		ActionList list(stack);
		generator.generate(list, static_cast<MoveGeneratorMode>(mode));
		for (int i = 0; i < list.size(); i++)
		{
			moves[i] = list[i].move.toShort();
			scores[i] = Score::to_short(list[i].score);
		}
		*flags = static_cast<int>(list.must_defend) | (static_cast<int>(list.has_initiative) << 1);
		return list.size();
	}
	// Integer-path rates of the pure reference code on one host core: setBoard + encode per second, addMove +
	// undoMove pairs per second, AlphaBetaSearch::solve(max_nodes) per second, each over the given boards for about `seconds` seconds.
	void agref_bench_integer_path(int rules, int rows, int cols, const int8_t *boards, int n_boards, double seconds, int max_nodes, double *rates)
	{
		GameConfig gc(static_cast<GameRules>(rules), rows, cols);
		const int cells = rows * cols;
		std::vector<matrix<Sign>> bs;
		std::vector<Sign> stm;
		std::vector<Move> moves;
		for (int i = 0; i < n_boards; i++)
		{
			matrix<Sign> b(rows, cols);
			int stones = 0, first_empty = -1;
			for (int j = 0; j < cells; j++)
			{
				b[j] = static_cast<Sign>(boards[static_cast<size_t>(i) * cells + j]);
				stones += (b[j] != Sign::NONE);
				if (b[j] == Sign::NONE and first_empty < 0 and j >= cells / 3)
					first_empty = j;
			}
			bs.push_back(b);
			stm.push_back((stones % 2 == 0) ? Sign::CROSS : Sign::CIRCLE);
			moves.push_back(Move(first_empty / cols, first_empty % cols, stm.back()));
		}
		PatternCalculator calc(gc);
		NNInputFeatures features(rows, cols);
		double t0 = getTime();
		long count = 0;
		while (getTime() - t0 < seconds)
			for (int i = 0; i < n_boards; i++, count++)
			{
				calc.setBoard(bs[i], stm[i]);
				features.encode(calc);
			}
		rates[0] = count / (getTime() - t0);
		t0 = getTime();
		count = 0;
		while (getTime() - t0 < seconds)
			for (int i = 0; i < n_boards; i++)
			{
				calc.setBoard(bs[i], stm[i]); // inside the timed region: one per 200 pairs
				for (int k = 0; k < 200; k++, count++)
				{
					calc.addMove(moves[i]);
					calc.undoMove(moves[i]);
				}
			}
		rates[1] = count / (getTime() - t0);
		AlphaBetaSearch solver(gc);
		solver.setDepthLimit(100);
		solver.setNodeLimit(max_nodes);
		solver.setTimeLimit(std::numeric_limits<double>::max());
		t0 = getTime();
		count = 0;
		while (getTime() - t0 < seconds)
			for (int i = 0; i < n_boards and getTime() - t0 < seconds; i++, count++)
			{
				SearchTask task(gc);
				task.set(bs[i], stm[i]);
				solver.solve(task);
			}
		rates[2] = count / (getTime() - t0);
	}
}
