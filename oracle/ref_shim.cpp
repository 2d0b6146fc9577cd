// TEST INFRASTRUCTURE -- C-ABI shim over the UNMODIFIED reference classes (compiled from /root/reference by
// oracle/Makefile into oracle/_ref/libagref.so). It is the checker for the restatement in oracle/ and for the
// CUDA path in tests/; nothing in the product links or loads it.
//
// Every entry point is a thin call into a reference class; no algorithm lives here.
#include <alphagomoku/game/Board.hpp>
#include <alphagomoku/game/rules.hpp>
#include <alphagomoku/utils/misc.hpp>
#include <alphagomoku/networks/NNInputFeatures.hpp>
#include <alphagomoku/patterns/DefensiveMoveTable.hpp>
#include <alphagomoku/patterns/PatternCalculator.hpp>
#include <alphagomoku/patterns/PatternTable.hpp>
#include <alphagomoku/patterns/ThreatTable.hpp>
#include <alphagomoku/utils/augmentations.hpp>
#include <alphagomoku/utils/configs.hpp>

#include <cstdint>
#include <cstring>
#include <memory>

using namespace ag;

namespace
{
	matrix<Sign> to_matrix(const int8_t *board, int rows, int cols)
	{
		matrix<Sign> result(rows, cols);
		for (int i = 0; i < rows * cols; i++)
			result[i] = static_cast<Sign>(board[i]);
		return result;
	}
	struct Calc
	{
			GameConfig cfg;
			PatternCalculator calc;
			Calc(GameRules rules, int rows, int cols) :
					cfg(rules, rows, cols),
					calc(cfg)
			{
			}
	};
}

extern "C"
{
	// ---- static tables (PatternTable.hpp:108-127, ThreatTable.hpp:79-91) ---------------------------------------
	// pattern_types[1<<20]: PatternEncoding byte per narrowed index; half_open_3[1<<20]: bit0 cross, bit1 circle;
	// update_mask[1<<20][2]: raw UpdateMask words (may be null); threats[4096][2]: cross, circle ThreatType
	void agref_dump_tables(int rules, uint8_t *pattern_types, uint8_t *half_open_3, uint32_t *update_mask, uint8_t *threats)
	{
		const PatternTable &pt = PatternTable::get(static_cast<GameRules>(rules));
		const ThreatTable &tt = ThreatTable::get(static_cast<GameRules>(rules));
		for (uint32_t i = 0; i < (1u << 20); i++)
		{
			const uint32_t expanded = (i & 1023u) | ((i & 1047552u) << 2u);
			const NormalPattern np(expanded);
			const PatternEncoding enc = pt.getPatternType(np);
			pattern_types[i] = static_cast<uint8_t>(enc.forCross()) | (static_cast<uint8_t>(enc.forCircle()) << 4);
			half_open_3[i] = static_cast<uint8_t>(pt.isHalfOpenThree(np, Sign::CROSS)) | (static_cast<uint8_t>(pt.isHalfOpenThree(np, Sign::CIRCLE)) << 1);
			if (update_mask != nullptr)
				for (int c = 0; c < 2; c++)
				{
					const UpdateMask um = pt.getUpdateMask(np, (c == 0) ? Sign::CROSS : Sign::CIRCLE);
					uint32_t raw = 0;
					for (int k = 0; k < 11; k++)
						raw |= static_cast<uint32_t>(um.get(k)) << (2 * k);
					update_mask[2 * i + c] = raw;
				}
		}
		for (int f0 = 0; f0 < 8; f0++)
			for (int f1 = 0; f1 < 8; f1++)
				for (int f2 = 0; f2 < 8; f2++)
					for (int f3 = 0; f3 < 8; f3++)
					{
						DirectionGroup<PatternType> g;
						g.horizontal = static_cast<PatternType>(f0);
						g.vertical = static_cast<PatternType>(f1);
						g.diagonal = static_cast<PatternType>(f2);
						g.antidiagonal = static_cast<PatternType>(f3);
						const int idx = f0 + (f1 << 3) + (f2 << 6) + (f3 << 9);
						threats[2 * idx + 0] = static_cast<uint8_t>(tt.getThreat<Sign::CROSS>(g));
						threats[2 * idx + 1] = static_cast<uint8_t>(tt.getThreat<Sign::CIRCLE>(g));
					}
	}
	// DefensiveMoveTable::getMoves (DefensiveMoveTable.cpp:393-477) for one 13-cell window
	uint16_t agref_defensive_moves(int rules, uint32_t extended_pattern, int defender_sign, int pattern_type)
	{
		const DefensiveMoveTable &dt = DefensiveMoveTable::get(static_cast<GameRules>(rules));
		return dt.getMoves(ExtendedPattern(extended_pattern), static_cast<Sign>(defender_sign), static_cast<PatternType>(pattern_type)).raw();
	}
	uint16_t agref_open3_promotion_moves(uint32_t normal_pattern)
	{
		return getOpenThreePromotionMoves(NormalPattern(normal_pattern)).raw();
	}

	// ---- PatternCalculator handle (PatternCalculator.hpp:79-189) --------------------------------------------------
	void* agref_calc_create(int rules, int rows, int cols)
	{
		return new Calc(static_cast<GameRules>(rules), rows, cols);
	}
	void agref_calc_destroy(void *h)
	{
		delete static_cast<Calc*>(h);
	}
	void agref_calc_set_board(void *h, const int8_t *board, int sign_to_move)
	{
		Calc *c = static_cast<Calc*>(h);
		c->calc.setBoard(to_matrix(board, c->cfg.rows, c->cfg.cols), static_cast<Sign>(sign_to_move));
	}
	void agref_calc_add_move(void *h, int row, int col, int sign)
	{
		static_cast<Calc*>(h)->calc.addMove(Move(row, col, static_cast<Sign>(sign)));
	}
	void agref_calc_undo_move(void *h, int row, int col, int sign)
	{
		static_cast<Calc*>(h)->calc.undoMove(Move(row, col, static_cast<Sign>(sign)));
	}
	// pattern_types[cells][4]: PatternEncoding-style byte per direction (low nibble cross, high nibble circle);
	// threats[cells][2]; legal[cells]; forbidden[cells] (isForbidden(CROSS,...)); raw[cells][4] normal patterns
	void agref_calc_dump(void *h, uint8_t *pattern_types, uint8_t *threats, uint8_t *legal, uint8_t *forbidden, uint32_t *raw)
	{
		Calc *c = static_cast<Calc*>(h);
		for (int row = 0; row < c->cfg.rows; row++)
			for (int col = 0; col < c->cfg.cols; col++)
			{
				const int idx = row * c->cfg.cols + col;
				const TwoPlayerGroup<DirectionGroup<PatternType>> tpg = c->calc.getPatternsAt(row, col);
				for (int dir = 0; dir < 4; dir++)
				{
					if (pattern_types != nullptr)
						pattern_types[4 * idx + dir] = static_cast<uint8_t>(tpg.for_cross[dir]) | (static_cast<uint8_t>(tpg.for_circle[dir]) << 4);
					if (raw != nullptr)
						raw[4 * idx + dir] = c->calc.getNormalPatternAt(row, col, dir);
				}
				if (threats != nullptr)
				{
					threats[2 * idx + 0] = static_cast<uint8_t>(c->calc.getThreatAt(Sign::CROSS, row, col));
					threats[2 * idx + 1] = static_cast<uint8_t>(c->calc.getThreatAt(Sign::CIRCLE, row, col));
				}
				if (legal != nullptr)
					legal[idx] = c->calc.getLegalMovesMask().at(row, col);
				if (forbidden != nullptr)
					forbidden[idx] = c->calc.isForbidden(Sign::CROSS, row, col);
			}
	}
	// threat histogram of one colour: counts[10], locations[10][cap] as (row | col << 8), in list order
	void agref_calc_histogram(void *h, int sign, int32_t *counts, uint16_t *locations, int cap)
	{
		Calc *c = static_cast<Calc*>(h);
		const ThreatHistogram &hist = c->calc.getThreatHistogram(static_cast<Sign>(sign));
		for (int t = 0; t < 10; t++)
		{
			const LocationList &list = hist.get(static_cast<ThreatType>(t));
			counts[t] = static_cast<int32_t>(list.size());
			for (size_t i = 0; i < list.size() and static_cast<int>(i) < cap; i++)
				locations[t * cap + i] = list[i].toShort();
		}
	}
	void agref_calc_encode(void *h, uint32_t *features)
	{
		Calc *c = static_cast<Calc*>(h);
		NNInputFeatures f(c->cfg.rows, c->cfg.cols);
		f.encode(c->calc);
		std::memcpy(features, f.data(), f.sizeInBytes());
	}
	int agref_calc_sign_to_move(void *h)
	{
		return static_cast<int>(static_cast<Calc*>(h)->calc.getSignToMove());
	}

	// ---- stateless helpers -------------------------------------------------------------------------------------
	// NNInputFeatures::augment (NNInputFeatures.cpp:114-154), in place
	void agref_augment(uint32_t *features, int rows, int cols, int mode)
	{
		NNInputFeatures f(rows, cols);
		std::memcpy(f.data(), features, f.sizeInBytes());
		f.augment(mode);
		std::memcpy(features, f.data(), f.sizeInBytes());
	}
	// getOutcome (rules.cpp:110-133); returns GameOutcome as int
	int agref_get_outcome(int rules, int rows, int cols, const int8_t *board, int row, int col, int sign, int draw_after)
	{
		return static_cast<int>(getOutcome(static_cast<GameRules>(rules), to_matrix(board, rows, cols), Move(row, col, static_cast<Sign>(sign)), draw_after));
	}
	// isForbidden(board, move) -- the raw-board path (rules.cpp:134-173)
	int agref_is_forbidden(int rows, int cols, const int8_t *board, int row, int col, int sign)
	{
		return static_cast<int>(isForbidden(to_matrix(board, rows, cols), Move(row, col, static_cast<Sign>(sign))));
	}
	// apply_symmetry on a float / int32 plane (augmentations.hpp), out of place
	void agref_apply_symmetry_f32(float *dst, const float *src, int rows, int cols, int mode)
	{
		matrix<float> s(rows, cols), d(rows, cols);
		std::memcpy(s.data(), src, s.sizeInBytes());
		apply_symmetry(d, s, int_to_symmetry(mode));
		std::memcpy(dst, d.data(), d.sizeInBytes());
	}
	int agref_inverse_symmetry(int mode)
	{
		return static_cast<int>(get_inverse_symmetry(int_to_symmetry(mode)));
	}
	// prepareOpening (src/utils/misc.cpp:142-170) with the library's own thread-local generator (seed 0 in this debug build, advanced
	// by every randInt / randFloat the calling thread has made so far)
	int agref_prepare_opening(int rules, int rows, int cols, int min_moves, uint16_t *moves)
	{
		const std::vector<Move> result = prepareOpening(GameConfig(static_cast<GameRules>(rules), rows, cols), min_moves);
		for (size_t i = 0; i < result.size(); i++)
			moves[i] = result[i].toShort();
		return static_cast<int>(result.size());
	}
}
