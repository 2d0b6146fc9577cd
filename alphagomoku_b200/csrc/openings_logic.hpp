// a22 (host logic): random openings exactly as the reference draws them. Included by openings.cpp (the library) and by
// tests/hostsim (CPU parity test against the reference's prepareOpening with the same seed).
//
// Reference: prepareOpening / generateOpeningMap / randomizeMove (src/utils/misc.cpp:84-170), randInt / randFloat
// (src/utils/random.cpp:27-56: std::mt19937 with uniform_int_distribution<int32_t> / uniform_real_distribution<float>).
// Compile with g++ (not nvcc) and without FMA contraction: the expressions below must resolve to the same libm overloads and
// round like the reference's.
#pragma once
#include "patterns_logic.cuh"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <random>
#include <vector>

namespace agb
{
	namespace openings
	{
		struct Random
		{
				std::mt19937 generator;
				explicit Random(uint32_t seed) :
						generator(seed)
				{
				}
				int32_t rand_int(int r)
				{
					std::uniform_int_distribution<int32_t> dist(0, r - 1);
					return dist(generator);
				}
				float rand_float()
				{
					std::uniform_real_distribution<float> dist(0.0f, 1.0f);
					return dist(generator);
				}
		};

		// generateOpeningMap (misc.cpp:108-141). `dist` is NOT cleared for an empty board: the reference accumulates onto what the
		// previous attempt of the same prepareOpening() call left there.
		inline void opening_map(const std::vector<int8_t> &board, std::vector<float> &dist, int rows, int cols, Random &st)
		{
			const bool empty = std::all_of(board.begin(), board.end(), [](int8_t v)
			{	return v == NONE;});
			if (empty)
			{
				for (int i = 0; i < rows; i++)
					for (int j = 0; j < cols; j++)
					{
						float d = std::hypot(0.5 + i - 0.5 * rows, 0.5 + j - 0.5 * cols) - 1;
						dist[i * cols + j] += pow(1.5f, -d);
					}
				return;
			}
			for (size_t i = 0; i < board.size(); i++)
				if (board[i] != NONE)
					dist[i] = 0.0f;
				else
					dist[i] = 1.0e-6f;
			float tmp = 2.0f + st.rand_float();
			for (int k = 0; k < rows; k++)
				for (int l = 0; l < cols; l++)
					if (board[k * cols + l] != NONE)
					{
						for (int i = 0; i < rows; i++)
							for (int j = 0; j < cols; j++)
								if (board[i * cols + j] == NONE)
								{
									float d = std::hypot(i - k, j - l) - 1;
									dist[i * cols + j] += pow(tmp, -d);
								}
					}
		}
		// randomizeMove (misc.cpp:84-102): index of the sampled cell (may be == size when rounding leaves r >= the total)
		inline int randomize_move(const std::vector<float> &policy, Random &st)
		{
			float r = std::accumulate(policy.begin(), policy.end(), 0.0f);
			if (r == 0.0f)
				return st.rand_int(static_cast<int>(policy.size()));
			r *= st.rand_float();
			float sum = 0.0f;
			size_t i = 0;
			for (; i < policy.size(); i++)
			{
				sum += policy[i];
				if (r < sum)
					break;
			}
			return static_cast<int>(i);
		}
		// prepareOpening (misc.cpp:142-170): moves as Move::toShort words; `board` receives the position
		inline std::vector<uint16_t> prepare_opening(int rules, int rows, int cols, const Tables &tables, Random &st, int min_moves, std::vector<int8_t> &board)
		{
			std::vector<float> map_dist(rows * cols, 0.0f);
			board.assign(rows * cols, NONE);
			while (true)
			{
				std::vector<uint16_t> result;
				std::fill(board.begin(), board.end(), static_cast<int8_t>(NONE));
				int sign_to_move = CROSS;
				int opening_moves = std::max(min_moves, st.rand_int(6) + st.rand_int(6) + st.rand_int(6));
				if (st.rand_int(1000) == 0)
					opening_moves = 0;
				int last_row = 0, last_col = 0;
				for (int i = 0; i < opening_moves; i++)
				{
					opening_map(board, map_dist, rows, cols, st);
					int cell = randomize_move(map_dist, st);
					if (cell >= rows * cols) // cannot be represented; the reference would assert here
						cell = rows * cols - 1;
					last_row = cell / cols;
					last_col = cell % cols;
					result.push_back(static_cast<uint16_t>(sign_to_move | (last_row << 2) | (last_col << 9)));
					board[cell] = static_cast<int8_t>(sign_to_move);
					sign_to_move = 3 - sign_to_move;
				}
				if (result.empty())
					return result;
				bool overflow = false;
				// getOutcome with its default draw rule (board full), as prepareOpening calls it (misc.cpp:167)
				if (plogic::outcome_of(board.data(), rows, rules, 0, last_row, last_col, 3 - sign_to_move, tables, overflow) == 0)
					return result;
			}
		}
	}
}
