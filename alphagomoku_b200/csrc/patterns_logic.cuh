// Per-cell pattern / threat / feature logic as host+device inline functions. patterns.cu wraps these in
// warp-per-board kernels; tests/hostsim compiles the same header with g++ to check the logic without a GPU.
//
// Line store (our own layout, same information as RawPatternCalculator.hpp:22-92): for a board of size S there are
// 6S-2 lines -- S rows, S columns, 2S-1 diagonals (col-row = const), 2S-1 antidiagonals (row+col = const) -- each one
// uint64 with 2 bits per cell, six ILLEGAL (0b11) padding cells in front of the first and behind the last cell.
// The 11-cell window centred on the cell at position t of a line is (line >> (2t + 2)) & 0x3FFFFF.
#pragma once
#include "agb_common.cuh"

namespace agb
{
	namespace plogic
	{
		AGB_HD inline int imin(int a, int b) { return a < b ? a : b; }
		AGB_HD inline int iabs(int a) { return a < 0 ? -a : a; }
		AGB_HD inline int dir_row_step(int dir) { return dir == 0 ? 0 : 1; }
		AGB_HD inline int dir_col_step(int dir) { return dir == 1 ? 0 : (dir == 3 ? -1 : 1); }

		AGB_HD inline int line_count(int S) { return 6 * S - 2; }
		AGB_HD inline int line_index(int dir, int r, int c, int S)
		{ // rows, then columns, then diagonals (by c - r), then antidiagonals (by r + c); written without branches: the solver calls it per move
			const int diagonal = (dir == 2) ? (3 * S - 1 - r + c) : (4 * S - 1 + r + c);
			const int straight = (dir == 0) ? r : (S + c);
			return (dir < 2) ? straight : diagonal;
		}
		AGB_HD inline int pos_in_line(int dir, int r, int c, int S)
		{
			const int a = (dir == 0) ? c : r;
			const int b = (dir == 3) ? (S - 1 - c) : ((dir == 1) ? r : c);
			return imin(a, b);
		}
		// first cell, direction and length of line l
		AGB_HD inline void line_geometry(int l, int S, int &r0, int &c0, int &dr, int &dc, int &len)
		{
			if (l < S)
			{
				r0 = l; c0 = 0; dr = 0; dc = 1; len = S;
			}
			else if (l < 2 * S)
			{
				r0 = 0; c0 = l - S; dr = 1; dc = 0; len = S;
			}
			else if (l < 4 * S - 1)
			{
				const int d = l - 2 * S - (S - 1); // col - row
				r0 = d >= 0 ? 0 : -d; c0 = d >= 0 ? d : 0; dr = 1; dc = 1; len = S - iabs(d);
			}
			else
			{
				const int a = l - (4 * S - 1); // row + col
				r0 = a < S ? 0 : a - S + 1; c0 = a < S ? a : S - 1; dr = 1; dc = -1; len = a < S ? a + 1 : 2 * S - 1 - a;
			}
		}
		AGB_HD inline uint64_t empty_line(int len)
		{
			return 0xFFFull | (0xFFFull << (12 + 2 * len));
		}
		AGB_HD inline uint64_t build_line(const int8_t *board, int S, int l)
		{
			int r0, c0, dr, dc, len;
			line_geometry(l, S, r0, c0, dr, dc, len);
			uint64_t w = empty_line(len);
			for (int t = 0; t < len; t++)
				w |= static_cast<uint64_t>(board[(r0 + t * dr) * S + c0 + t * dc] & 3) << (12 + 2 * t);
			return w;
		}
		AGB_HD inline uint32_t normal_window(const uint64_t *lines, int dir, int r, int c, int S)
		{
			return static_cast<uint32_t>(lines[line_index(dir, r, c, S)] >> (2 * pos_in_line(dir, r, c, S) + 2)) & 0x3FFFFFu;
		}
		AGB_HD inline uint32_t extended_window(const uint64_t *lines, int dir, int r, int c, int S)
		{
			return static_cast<uint32_t>(lines[line_index(dir, r, c, S)] >> (2 * pos_in_line(dir, r, c, S))) & 0x3FFFFFFu;
		}
		// the four direction bytes of an EMPTY cell (table byte incl. the half-open-three side bits)
		AGB_HD inline uint32_t classify_cell(const uint64_t *lines, const uint8_t *pattern_table, int r, int c, int S)
		{
			uint32_t p = 0;
#pragma unroll
			for (int dir = 0; dir < 4; dir++)
				p |= static_cast<uint32_t>(pattern_table[narrow_window(normal_window(lines, dir, r, c, S))]) << (8 * dir);
			return p;
		}
		AGB_HD inline uint32_t threat_index(uint32_t ptypes, int shift)
		{ // shift 0: cross nibbles, 4: circle nibbles
			const uint32_t p = ptypes >> shift;
			return (p & 7u) | (((p >> 8) & 7u) << 3) | (((p >> 16) & 7u) << 6) | (((p >> 24) & 7u) << 9);
		}
		AGB_HD inline uint8_t threat_of_cell(uint32_t ptypes, const uint8_t *threat_table)
		{ // low nibble: for cross, high nibble: for circle (ThreatTable.hpp:79-91)
			return (threat_table[threat_index(ptypes, 0)] & 0x0F) | (threat_table[threat_index(ptypes, 4)] & 0xF0);
		}

		// NNInputFeatures.cpp:15-32 + :65-103 for one cell (without the forbidden bit)
		AGB_HD inline uint32_t pattern_feature_bits(uint32_t nibbles)
		{ // nibbles: 4 bytes whose low 3 bits are the pattern types of one colour
			uint32_t r = 0;
#pragma unroll
			for (int dir = 0; dir < 4; dir++)
			{
				const uint32_t t = (nibbles >> (8 * dir)) & 7u;
				r |= (t == PT_OPEN_3) ? (1u << dir) : 0u;
				r |= (t == PT_HALF_OPEN_4) ? (1u << (4 + dir)) : 0u;
				r |= (t >= PT_OPEN_4) ? (1u << (8 + t - PT_OPEN_4)) : 0u;
			}
			return r;
		}
		AGB_HD inline uint32_t encode_cell(int cell_sign, uint32_t ptypes, int sign_to_move)
		{
			uint32_t f = (1u << 3) | ((sign_to_move == CROSS) ? (1u << 4) : (1u << 5));
			if (cell_sign == NONE)
				f |= 1u;
			else if (cell_sign == sign_to_move)
				f |= 2u;
			else if (cell_sign == CROSS or cell_sign == CIRCLE)
				f |= 4u;
			const uint32_t cross_bits = pattern_feature_bits(ptypes);
			const uint32_t circle_bits = pattern_feature_bits(ptypes >> 4);
			if (sign_to_move == CROSS)
				f |= (cross_bits << 8) | (circle_bits << 20);
			else
				f |= (cross_bits << 20) | (circle_bits << 8);
			return f;
		}

		// ---- renju: black forbidden moves on a raw board (rules.cpp:134-173, RawPatternCalculator.hpp:113-176) ------
		// The recursion of the reference places one extra black stone per level; those stones are kept in a small
		// overlay instead of copying the board.
		constexpr int kMaxOverlay = 12;
		struct Overlay
		{
				int8_t row[kMaxOverlay];
				int8_t col[kMaxOverlay];
				int count = 0;
				bool overflow = false;
		};
		AGB_HD inline int cell_at(const int8_t *board, int S, int r, int c, const Overlay &ov)
		{
			if (r < 0 or r >= S or c < 0 or c >= S)
				return ILLEGAL;
			for (int k = 0; k < ov.count; k++)
				if (ov.row[k] == r and ov.col[k] == c)
					return CROSS;
			return board[r * S + c];
		}
		AGB_HD inline uint32_t raw_window(const int8_t *board, int S, int r, int c, int dir, const Overlay &ov)
		{ // 11 cells centred on (r, c), centre forced empty
			uint32_t w = 0;
			const int dr = dir_row_step(dir), dc = dir_col_step(dir);
			for (int i = -5; i <= 5; i++)
				if (i != 0)
					w |= static_cast<uint32_t>(cell_at(board, S, r + i * dr, c + i * dc, ov)) << (2 * (i + 5));
			return w;
		}
		AGB_HD inline bool makes_straight_four(uint32_t window)
		{ // black stone on the centre, then any four consecutive black stones inside the 11 cells
			uint32_t w = window | (static_cast<uint32_t>(CROSS) << 10);
			for (int i = 0; i < 7; i++, w >>= 2)
				if ((w & 0xFFu) == 0x55u)
					return true;
			return false;
		}
		// cells that turn the open three through the (empty) centre into a four; bit k = window cell k
		// (getOpenThreePromotionMoves, DefensiveMoveTable.cpp:329-378)
		AGB_HD inline uint32_t open_three_promotions(uint32_t window)
		{
			// shapes "_XXX__", "_XX_X_", "_X_XX_", "__XXX_" as stone masks over 6 cells, and their promotion cells
			const uint32_t stones[4] = { 0b001110u, 0b010110u, 0b011010u, 0b011100u };
			const uint32_t promo[4] = { 0b110001u, 0b101001u, 0b100101u, 0b100011u };
			for (int s = 0; s < 4; s++)
				for (int j = 5; j >= 0; j--)
					if ((stones[s] >> j) & 1u)
					{ // shape cell j sits on the window centre
						const int start = 5 - j;
						bool ok = true;
						for (int m = 0; m < 6 and ok; m++)
						{
							const uint32_t want = (m != j and ((stones[s] >> m) & 1u)) ? CROSS : NONE;
							ok = ((window >> (2 * (start + m))) & 3u) == want;
						}
						if (ok)
							return promo[s] << start;
					}
			return 0;
		}
		AGB_HD inline bool is_forbidden_raw(const int8_t *board, int S, int r, int c, const Tables &tables, Overlay &ov)
		{
			uint32_t windows[4];
			uint32_t ptypes = 0;
			for (int dir = 0; dir < 4; dir++)
			{
				windows[dir] = raw_window(board, S, r, c, dir, ov);
				ptypes |= static_cast<uint32_t>(tables.pattern[narrow_window(windows[dir])] & 7u) << (8 * dir);
			}
			int threat = tables.threat[threat_index(ptypes, 0)] & 0x0F;
			if (threat == TT_FORK_3x3)
			{
				if (ov.count >= kMaxOverlay)
				{
					ov.overflow = true;
					return true;
				}
				ov.row[ov.count] = static_cast<int8_t>(r);
				ov.col[ov.count] = static_cast<int8_t>(c);
				ov.count++;
				for (int dir = 0; dir < 4; dir++)
					if (((ptypes >> (8 * dir)) & 7u) == PT_OPEN_3)
					{
						const uint32_t promotions = open_three_promotions(windows[dir]);
						const int dr = dir_row_step(dir), dc = dir_col_step(dir);
						bool real_three = false;
						for (int i = -5; i <= 5 and not real_three; i++)
							if (i != 0 and ((promotions >> (i + 5)) & 1u))
							{
								const int pr = r + i * dr, pc = c + i * dc;
								if (cell_at(board, S, pr, pc, ov) == NONE and makes_straight_four(raw_window(board, S, pr, pc, dir, ov))
										and not is_forbidden_raw(board, S, pr, pc, tables, ov))
									real_three = true;
							}
						if (not real_three)
							ptypes &= ~(7u << (8 * dir));
					}
				ov.count--;
				threat = tables.threat[threat_index(ptypes, 0)] & 0x0F;
			}
			return threat == TT_OVERLINE or threat == TT_FORK_4x4 or threat == TT_FORK_3x3;
		}

		// getOutcome (rules.cpp:110-133): the move may or may not already be on the board
		AGB_HD inline int outcome_of(const int8_t *board, int S, int rules, int draw_after, int r, int c, int sign, const Tables &tables,
				bool &overflow)
		{
			if (r < 0 or r >= S or c < 0 or c >= S)
				return 0;
			Overlay ov;
			bool win = false;
			for (int dir = 0; dir < 4; dir++)
			{
				const uint8_t e = tables.pattern[narrow_window(raw_window(board, S, r, c, dir, ov))];
				const int t = (sign == CROSS) ? (e & 7) : ((e >> 4) & 7);
				win = win or (t == PT_FIVE);
			}
			if (win)
				return sign == CROSS ? 2 : 3;
			if (rules == RULE_RENJU and sign == CROSS)
			{
				const bool f = is_forbidden_raw(board, S, r, c, tables, ov);
				overflow = overflow or ov.overflow;
				if (f)
					return 3;
			}
			int stones = 0;
			for (int i = 0; i < S * S; i++)
				stones += (board[i] != NONE);
			const bool is_draw = (draw_after > 0) ? (stones >= draw_after) : (stones == S * S);
			return is_draw ? 1 : 0;
		}

		// board symmetries: source cell that lands on (r, c) when Symmetry `mode` (0..7, augmentations.hpp:19-29) is applied
		// to a square board (apply_symmetry, augmentations.hpp:141-214)
		AGB_HD inline void symmetry_source(int mode, int S, int r, int c, int &sr, int &sc)
		{
			const int last = S - 1;
			switch (mode)
			{
				default:
				case 0: sr = r; sc = c; break;
				case 1: sr = last - r; sc = c; break; // flip vertically
				case 2: sr = r; sc = last - c; break; // flip horizontally
				case 3: sr = last - r; sc = last - c; break; // rotate 180
				case 4: sr = c; sc = r; break; // flip diagonally
				case 5: sr = last - c; sc = last - r; break; // flip antidiagonally
				case 6: sr = c; sc = last - r; break; // rotate 90
				case 7: sr = last - c; sc = r; break; // rotate 270
			}
		}
		AGB_HD inline int inverse_symmetry(int mode)
		{
			return mode == 6 ? 7 : (mode == 7 ? 6 : mode);
		}
		// permutation of the per-direction feature bits that goes with a board symmetry (NNInputFeatures.cpp:114-154):
		// reflections about an axis swap the two diagonals, reflections about a diagonal swap rows and columns,
		// quarter turns do both
		AGB_HD inline uint32_t permute_direction_bits(uint32_t f, int mode)
		{
			const bool swap_hv = (mode >= 4);
			const bool swap_da = (mode == 1 or mode == 2 or mode == 6 or mode == 7);
			const uint32_t groups = 0x0FF0FF00u; // bits 8-15 and 20-27: four nibbles of direction flags
			uint32_t out = 0;
			for (int g = 0; g < 4; g++)
			{
				const int base = (g < 2) ? (8 + 4 * g) : (20 + 4 * (g - 2));
				uint32_t nib = (f >> base) & 0xFu;
				if (swap_hv)
					nib = (nib & 0xCu) | ((nib & 1u) << 1) | ((nib >> 1) & 1u);
				if (swap_da)
					nib = (nib & 0x3u) | ((nib & 4u) << 1) | ((nib >> 1) & 4u);
				out |= nib << base;
			}
			return (f & ~groups) | out;
		}
	}
}
