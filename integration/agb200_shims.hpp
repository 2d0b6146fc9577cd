// Reference-side bindings for libagb200.so: what a maintainer of the reference adds to its tree to use the B200 engine
// (see INTEGRATION.md). Header-only, C++17, includes the reference's own headers; nothing here is part of the library.
// tests/test_oracle_cpu.py compiles this file against /root/reference/include to keep it in step with both sides.
#pragma once
#include <agb200.h>

#include <alphagomoku/game/Move.hpp>
#include <alphagomoku/search/Score.hpp>
#include <alphagomoku/search/Value.hpp>
#include <alphagomoku/search/monte_carlo/SearchTask.hpp>
#include <alphagomoku/utils/configs.hpp>
#include <alphagomoku/utils/matrix.hpp>

#include <cstring>
#include <stdexcept>
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
} /* namespace ag */
