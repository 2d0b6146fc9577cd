// Reference-side bindings for libagb200.so: what a maintainer of the reference adds to its tree to use the B200 engine
// (see INTEGRATION.md). Header-only, C++17, includes the reference's own headers; nothing here is part of the library.
// tests/test_oracle_cpu.py compiles this file against /root/reference/include to keep it in step with both sides.
#pragma once
#include <agb200.h>

#include <alphagomoku/dataset/GameDataBuffer.hpp>
#include <alphagomoku/dataset/GameDataStorage.hpp>
#include <alphagomoku/game/Move.hpp>
#include <alphagomoku/search/Score.hpp>
#include <alphagomoku/search/Value.hpp>
#include <alphagomoku/search/monte_carlo/SearchTask.hpp>
#include <alphagomoku/utils/configs.hpp>
#include <alphagomoku/utils/matrix.hpp>

#include <minml/utils/serialization.hpp>

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ag
{
	// Drop-in for NNEvaluator (include/alphagomoku/search/monte_carlo/NNEvaluator.hpp:42-83): same queue semantics, evaluation on the device
	class NNEvaluatorB200
	{
			AgbEngine *engine = nullptr;
			std::vector<std::pair<SearchTask*, int>> waiting_queue; // NNEvaluator::waiting_queue (task, symmetry or -1)
			std::vector<int8_t> boards, stm, symmetry;
			std::vector<float> policy, value, q;
			int batch_size = 0, cells = 0;
			bool q_head = false;
		public:
			NNEvaluatorB200(const GameConfig &game, const DeviceConfig &cfg, int blocks, int filters, bool qHead) :
					batch_size(cfg.batch_size),
					cells(game.rows * game.cols),
					q_head(qHead)
			{
				AgbConfig c { };
				c.rules = static_cast<int>(game.rules);
				c.rows = game.rows;
				c.cols = game.cols;
				c.draw_after = game.draw_after;
				c.device = 0;
				c.max_boards = cfg.batch_size;
				c.blocks = blocks;
				c.filters = filters;
				c.q_head = qHead;
				if (agb_create(&c, &engine) != AGB_OK)
					throw std::runtime_error(agb_last_error(nullptr)); // the reference's error convention (NNEvaluator.cpp:149)
			}
			NNEvaluatorB200(const NNEvaluatorB200&) = delete;
			NNEvaluatorB200& operator=(const NNEvaluatorB200&) = delete;
			~NNEvaluatorB200()
			{
				agb_destroy(engine);
			}
			void loadGraph(const void *blob, size_t bytes) // NNEvaluator::loadGraph (NNEvaluator.cpp:121-129)
			{
				if (agb_load_weights(engine, blob, bytes) != AGB_OK)
					throw std::logic_error(agb_last_error(engine));
			}
			void addToQueue(SearchTask &task) // NNEvaluator.cpp:134-139 (the caller picks the symmetry, or -1 for none)
			{
				waiting_queue.push_back( { &task, -1 });
			}
			void addToQueue(SearchTask &task, int sym) // NNEvaluator.cpp:140-146
			{
				waiting_queue.push_back( { &task, sym });
			}
			bool isQueueFull() const noexcept
			{
				return static_cast<int>(waiting_queue.size()) >= batch_size;
			}
			int getQueueSize() const noexcept
			{
				return static_cast<int>(waiting_queue.size());
			}
			void clearQueue() noexcept
			{
				waiting_queue.clear();
			}
			double evaluateGraph() // NNEvaluator.cpp:147-181: pack -> forward -> unpack
			{
				const int n = static_cast<int>(waiting_queue.size());
				if (n == 0)
					return 0.0;
				boards.resize(static_cast<size_t>(n) * cells);
				stm.resize(n);
				symmetry.resize(n);
				policy.resize(static_cast<size_t>(n) * cells);
				value.resize(static_cast<size_t>(n) * 3);
				q.resize(static_cast<size_t>(n) * cells * 3);
				bool any_symmetry = false;
				for (int i = 0; i < n; i++)
				{
					const matrix<Sign> &b = waiting_queue[i].first->getBoard();
					for (int j = 0; j < cells; j++)
						boards[static_cast<size_t>(i) * cells + j] = static_cast<int8_t>(b[j]);
					stm[i] = static_cast<int8_t>(waiting_queue[i].first->getSignToMove());
					symmetry[i] = static_cast<int8_t>(waiting_queue[i].second < 0 ? 0 : waiting_queue[i].second);
					any_symmetry = any_symmetry or waiting_queue[i].second > 0;
				}
				if (agb_evaluate(engine, boards.data(), stm.data(), any_symmetry ? symmetry.data() : nullptr, n, policy.data(), value.data(),
						q_head ? q.data() : nullptr) != AGB_OK)
					throw std::runtime_error(agb_last_error(engine));
				for (int i = 0; i < n; i++)
				{ // NNEvaluator::unpack_from_network (NNEvaluator.cpp:263-286); the engine has already undone the symmetry
					SearchTask &t = *waiting_queue[i].first;
					std::memcpy(t.getPolicy().data(), policy.data() + static_cast<size_t>(i) * cells, cells * sizeof(float));
					if (q_head)
						for (int j = 0; j < cells; j++)
						{
							const float *src = q.data() + (static_cast<size_t>(i) * cells + j) * 3;
							t.getActionValues()[j] = Value(src[0], src[1]);
						}
					t.setValue(Value(value[3 * i], value[3 * i + 1]));
					t.markAsProcessedByNetwork();
				}
				waiting_queue.clear();
				return 0.0;
			}
	};

	// AlphaBetaSearch::solve (src/search/alpha_beta/AlphaBetaSearch.cpp:77-156) for a batch of tasks, on the device
	inline void solve_tasks_b200(AgbEngine *engine, std::vector<SearchTask*> &tasks, int maxPositions)
	{
		const int n = static_cast<int>(tasks.size());
		if (n == 0)
			return;
		const int cells = tasks[0]->getBoard().size();
		std::vector<int8_t> boards(static_cast<size_t>(n) * cells), stm(n);
		std::vector<uint16_t> scores(n), moves(static_cast<size_t>(n) * cells), action_scores(static_cast<size_t>(n) * cells);
		std::vector<int32_t> n_actions(n), flags(n);
		for (int i = 0; i < n; i++)
		{
			for (int j = 0; j < cells; j++)
				boards[static_cast<size_t>(i) * cells + j] = static_cast<int8_t>(tasks[i]->getBoard()[j]);
			stm[i] = static_cast<int8_t>(tasks[i]->getSignToMove());
		}
		if (agb_solve(engine, boards.data(), stm.data(), n, maxPositions, scores.data(), n_actions.data(), moves.data(), action_scores.data(), flags.data())
				!= AGB_OK)
			throw std::runtime_error(agb_last_error(engine));
		for (int i = 0; i < n; i++)
		{ // what AlphaBetaSearch.cpp:114-135 writes into the task
			SearchTask &t = *tasks[i];
			for (int k = 0; k < n_actions[i]; k++)
			{
				const Move m(moves[static_cast<size_t>(i) * cells + k]);
				t.getActionScores().at(m.row, m.col) = Score::from_short(action_scores[static_cast<size_t>(i) * cells + k]);
				if (t.getActionScores().at(m.row, m.col).isProven())
					t.getActionValues().at(m.row, m.col) = t.getActionScores().at(m.row, m.col).convertToValue();
				t.addEdge(m);
			}
			t.setScore(Score::from_short(scores[i]));
			if (t.getScore().isProven())
			{
				t.setValue(t.getScore().convertToValue());
				t.setMovesLeft(t.getScore().getDistance());
			}
			if (flags[i] & 1)
				t.markAsDefensive();
			if (t.getScore().isProven())
				t.maskAsRecursivelySolved();
			if ((flags[i] >> 8) <= 1)
				t.markAsStaticallySolved();
			t.markAsProcessedBySolver();
		}
	}

	// GeneratorManager's surface (include/alphagomoku/selfplay/GeneratorManager.hpp:87-102) over the device engine: the whole self-play loop of
	// every GeneratorThread -- select, solve, evaluate, expand, backup, make move, record -- is agb_step; finished games arrive through
	// agb_pop_finished as GameDataStorage::serialize blobs (format 201) and are added to a real GameDataBuffer, so getGameBuffer() gives the
	// trainer exactly what the reference's manager gives it. saveState / loadState keep the reference's directory layout
	// (<working directory>/saved_state/buffer.bin written by GameDataBuffer::save; the games in flight in engine.bin, the engine's own blob).
	class GeneratorManagerB200
	{
			AgbEngine *engine = nullptr;
			GameDataBuffer game_buffer;
			std::string working_directory;
			std::vector<uint8_t> records;
			int steps_per_poll;
		public:
			// `config` carries GameConfig and the SelfplayConfig / SearchConfig fields the engine honours (agb_config_from_json fills it from the
			// reference's config.json); weights: fp32 blob (NetworkLoader's role, see NNEvaluator_b200.cpp)
			GeneratorManagerB200(const GameConfig &gameOptions, const AgbConfig &config, const void *weights, size_t weightBytes, int stepsPerPoll = 10) :
					game_buffer(gameOptions),
					records(64u << 20),
					steps_per_poll(stepsPerPoll)
			{
				if (agb_create(&config, &engine) != AGB_OK)
					throw std::runtime_error(std::string("GeneratorManagerB200 : ") + agb_last_error(nullptr));
				if (agb_load_weights(engine, weights, weightBytes) != AGB_OK)
					fail("agb_load_weights");
			}
			GeneratorManagerB200(const GeneratorManagerB200&) = delete;
			GeneratorManagerB200& operator=(const GeneratorManagerB200&) = delete;
			~GeneratorManagerB200()
			{
				agb_destroy(engine);
			}
			void setWorkingDirectory(const std::string &path)
			{
				working_directory = path;
			}
			const GameDataBuffer& getGameBuffer() const noexcept
			{
				return game_buffer;
			}
			GameDataBuffer& getGameBuffer() noexcept
			{
				return game_buffer;
			}
			bool hasEnoughGames(int numberOfGames) const noexcept
			{
				return game_buffer.numberOfGames() >= numberOfGames;
			}
			AgbEngine* getEngine() noexcept
			{
				return engine;
			}
			// start every game from a given position (or empty boards): the opening generator's output goes here (agb_generate_openings)
			void resetGames(const int8_t *boards = nullptr, const int8_t *signToMove = nullptr)
			{
				if (agb_selfplay_reset(engine, boards, signToMove) != AGB_OK)
					fail("agb_selfplay_reset");
			}
			// GeneratorManager::generate (GeneratorManager.cpp:177-218): play until the buffer holds numberOfGames games
			void generate(int numberOfGames)
			{
				while (not hasEnoughGames(numberOfGames))
				{
					if (agb_step(engine, steps_per_poll) != AGB_OK)
						fail("agb_step");
					collectFinishedGames();
				}
			}
			// GameGenerator.cpp:104-111 (manager.addToBuffer) for every game the device finished since the last call
			int collectFinishedGames()
			{
				size_t used = 0;
				int n_games = 0;
				if (agb_pop_finished(engine, records.data(), records.size(), &used, &n_games) != AGB_OK)
					fail("agb_pop_finished");
				SerializedObject so;
				so.save(records.data(), used);
				size_t offset = 0;
				for (int i = 0; i < n_games; i++)
					game_buffer.addGameData(GameDataStorage(so, offset, 201));
				if (offset != used)
					throw std::runtime_error("GeneratorManagerB200 : finished-game records do not parse");
				return n_games;
			}
			void saveState(bool saveBuffer)
			{ // GeneratorManager.cpp:240-263
				if (working_directory.empty())
					return;
				const std::string path = working_directory + "/saved_state/";
				if (saveBuffer)
					game_buffer.save(path + "buffer.bin");
				collectFinishedGames(); // nothing finished may stay on the device: the saved games are the ones in flight
				size_t used = 0;
				agb_save_games(engine, nullptr, 0, &used);
				std::vector<uint8_t> blob(used);
				if (agb_save_games(engine, blob.data(), blob.size(), &used) != AGB_OK)
					fail("agb_save_games");
				FILE *f = std::fopen((path + "engine.bin").c_str(), "wb");
				if (f == nullptr or std::fwrite(blob.data(), 1, used, f) != used)
					throw std::runtime_error("GeneratorManagerB200::saveState() : cannot write " + path + "engine.bin");
				std::fclose(f);
			}
			void loadState()
			{ // GeneratorManager.cpp:264-290
				if (working_directory.empty())
					return;
				const std::string path = working_directory + "/saved_state/";
				if (FILE *f = std::fopen((path + "buffer.bin").c_str(), "rb"))
				{
					std::fclose(f);
					game_buffer.load(path + "buffer.bin");
				}
				if (FILE *f = std::fopen((path + "engine.bin").c_str(), "rb"))
				{
					std::vector<uint8_t> blob;
					uint8_t chunk[65536];
					size_t n;
					while ((n = std::fread(chunk, 1, sizeof(chunk), f)) > 0)
						blob.insert(blob.end(), chunk, chunk + n);
					std::fclose(f);
					if (agb_load_games(engine, blob.data(), blob.size()) != AGB_OK)
						fail("agb_load_games");
				}
			}
		private:
			[[noreturn]] void fail(const char *what) const
			{
				throw std::runtime_error(std::string("GeneratorManagerB200 : ") + what + " : " + agb_last_error(engine));
			}
	};
} /* namespace ag */
