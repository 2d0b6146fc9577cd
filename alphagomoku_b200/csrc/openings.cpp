// a22: opening generation for the lockstep engine. Host code (compiled by g++, not nvcc: the floating-point expressions below
// must resolve to the same libm overloads as in the reference) on top of the device solver (agb_solve) and network (agb_evaluate).
//
// Reference: OpeningGenerator::generate (src/selfplay/OpeningGenerator.cpp:21-78), prepareOpening / generateOpeningMap /
// randomizeMove (src/utils/misc.cpp:84-170), randInt / randFloat (src/utils/random.cpp:27-56: std::mt19937 with
// uniform_int_distribution<int32_t> / uniform_real_distribution<float>).
//
// prepare_opening() draws from the generator exactly like the reference's prepareOpening(): seeded alike, both produce the same
// moves (tests/test_openings_cpu.py). generate() keeps the reference's acceptance rule (solver at 1000 positions must leave the
// position unproven, then |expectation - 0.5| < 0.1 + 0.01 * trials with the shared trials counter, in workspace order) but
// evaluates a whole workspace of candidates per device call instead of one task at a time.
#include <cuda_runtime.h>

#include "engine.hpp"
#include "openings_logic.hpp"

extern "C"
{
	int agb_solve(AgbEngine *e, const int8_t *boards_host, const int8_t *sign_to_move_host, int n, int max_positions, uint16_t *scores_host,
			int32_t *n_actions_host, uint16_t *moves_host, uint16_t *action_scores_host, int32_t *flags_host);
	int agb_evaluate(AgbEngine *e, const int8_t *boards_host, const int8_t *sign_to_move_host, const int8_t *symmetry_host, int n, float *policy_host,
			float *value_host, float *q_host);
}

namespace agb
{
	namespace
	{
		struct OpeningState
		{
				openings::Random random;
				int trials = 0; // OpeningGenerator::trials
				std::vector<uint8_t> pattern_table, threat_table; // host copies for the outcome check
				explicit OpeningState(uint32_t seed) :
						random(seed)
				{
				}
		};
		OpeningState* state_of(AgbEngine *e)
		{
			if (e->opening_rng == nullptr)
			{
				OpeningState *st = new OpeningState(static_cast<uint32_t>(e->cfg.seed));
				st->pattern_table.resize(1u << 20);
				st->threat_table.resize(4096);
				cudaMemcpy(st->pattern_table.data(), e->d_pattern, st->pattern_table.size(), cudaMemcpyDeviceToHost);
				cudaMemcpy(st->threat_table.data(), e->d_threat, st->threat_table.size(), cudaMemcpyDeviceToHost);
				e->opening_rng = st;
			}
			return static_cast<OpeningState*>(e->opening_rng);
		}
		std::vector<uint16_t> prepare_opening(AgbEngine *e, OpeningState &st, int min_moves, std::vector<int8_t> &board)
		{
			const Tables tables { st.pattern_table.data(), st.threat_table.data() };
			return openings::prepare_opening(e->cfg.rules, e->cfg.rows, e->cfg.cols, tables, st.random, min_moves, board);
		}
	}
	void openings_destroy(AgbEngine *e)
	{
		delete static_cast<OpeningState*>(e->opening_rng);
		e->opening_rng = nullptr;
	}
}

extern "C"
{
	using namespace agb;

	int agb_seed_openings(AgbEngine *e, uint32_t seed)
	{
		OpeningState *st = state_of(e);
		st->random.generator.seed(seed);
		st->trials = 0;
		return AGB_OK;
	}
	int agb_prepare_opening(AgbEngine *e, int min_moves, uint16_t *moves_host, int32_t *n_moves)
	{
		if (moves_host == nullptr or n_moves == nullptr)
			return e->fail(AGB_EINVAL, "null pointer");
		OpeningState *st = state_of(e);
		std::vector<int8_t> board;
		const std::vector<uint16_t> moves = prepare_opening(e, *st, min_moves, board);
		*n_moves = static_cast<int32_t>(moves.size());
		std::copy(moves.begin(), moves.end(), moves_host);
		return AGB_OK;
	}
	int agb_generate_openings(AgbEngine *e, int count, int8_t *boards_host, int8_t *sign_to_move_host)
	{
		if (boards_host == nullptr or sign_to_move_host == nullptr)
			return e->fail(AGB_EINVAL, "null pointer");
		if (e->net == nullptr)
			return e->fail(AGB_ESTATE, "no weights loaded");
		OpeningState *st = state_of(e);
		const int cells = e->cells;
		const int workspace = std::min(std::max(count, 8), std::min(e->cfg.max_boards, 256)); // openings under construction per device call
		std::vector<int8_t> boards(static_cast<size_t>(workspace) * cells), stm(workspace), board;
		std::vector<uint16_t> scores(workspace);
		std::vector<float> policy(static_cast<size_t>(workspace) * cells), value(static_cast<size_t>(workspace) * 3);
		int completed = 0;
		while (completed < count)
		{
			// one unproven candidate per workspace entry: up to 100 attempts each (OpeningGenerator.cpp:55-71), the solver calls batched
			std::vector<int> pending(workspace);
			std::iota(pending.begin(), pending.end(), 0);
			std::vector<uint8_t> ready(workspace, 0);
			for (int attempt = 0; attempt < 100 and not pending.empty(); attempt++)
			{
				std::vector<int8_t> cand_boards(pending.size() * cells), cand_stm(pending.size());
				for (size_t k = 0; k < pending.size(); k++)
				{
					const std::vector<uint16_t> moves = prepare_opening(e, *st, 1, board);
					std::copy(board.begin(), board.end(), cand_boards.begin() + k * cells);
					cand_stm[k] = static_cast<int8_t>(moves.empty() ? CROSS : 3 - (moves.back() & 3));
				}
				std::vector<uint16_t> cand_scores(pending.size());
				const int rc = agb_solve(e, cand_boards.data(), cand_stm.data(), static_cast<int>(pending.size()), 1000, cand_scores.data(), nullptr, nullptr, nullptr,
						nullptr);
				if (rc != AGB_OK)
					return rc;
				std::vector<int> still;
				for (size_t k = 0; k < pending.size(); k++)
				{
					const uint16_t s = cand_scores[k];
					const bool proven = ((s >> 13) & 3) != 2 and s != 0x0000 and s != 0xFFFF; // Score::isProven
					if (proven)
						still.push_back(pending[k]);
					else
					{
						const int w = pending[k];
						std::copy(cand_boards.begin() + k * cells, cand_boards.begin() + (k + 1) * cells, boards.begin() + static_cast<size_t>(w) * cells);
						stm[w] = cand_stm[k];
						ready[w] = 1;
					}
				}
				pending.swap(still);
			}
			// network evaluation of the scheduled candidates, then the balance test in workspace order (OpeningGenerator.cpp:43-53)
			const int rc = agb_evaluate(e, boards.data(), stm.data(), nullptr, workspace, policy.data(), value.data(), nullptr);
			if (rc != AGB_OK)
				return rc;
			for (int w = 0; w < workspace and completed < count; w++)
			{
				if (not ready[w])
					continue;
				const float expectation = value[3 * w + 0] + 0.5f * value[3 * w + 1]; // Value::getExpectation
				const float balance = std::fabs(expectation - 0.5f);
				if (balance < (0.1f + 0.01f * st->trials))
				{
					std::copy(boards.begin() + static_cast<size_t>(w) * cells, boards.begin() + static_cast<size_t>(w + 1) * cells,
							boards_host + static_cast<size_t>(completed) * cells);
					sign_to_move_host[completed] = stm[w];
					completed++;
					st->trials = 0;
				}
				else
					st->trials++;
			}
		}
		return AGB_OK;
	}
}
