// K5 (search layer): the reference's alpha-beta threat-space search on top of the static layer, as host+device code.
//
// Reference: AlphaBetaSearch::solve / recursive_solve / evaluate (src/search/alpha_beta/AlphaBetaSearch.cpp:77-156, 185-339,
// 345-365), ActionList / ActionStack (include/alphagomoku/search/alpha_beta/ActionList.hpp:262-420), SharedHashTable
// (include/alphagomoku/search/alpha_beta/SharedHashTable.hpp:27-220), FastZobristHashing (ZobristHashing.hpp:111-127),
// PatternCalculator::addMove / undoMove / update_around (src/patterns/PatternCalculator.cpp:68-106, 278-366) and
// PatternCalculator::isForbidden / is_3x3_forbidden (PatternCalculator.hpp:173-189, PatternCalculator.cpp:213-244).
//
// One warp owns one game: it solves that game's leaf positions one after another (they share the game's transposition
// table, so the order is part of the result). All 32 lanes execute the sequential logic in lockstep on identical data (a warp
// instruction costs the same with 1 or 32 active lanes) and split the data-parallel part, the 40-cell pattern update of every
// move made or taken back, between them. The recursion of the reference is unrolled into an explicit frame stack;
// the position is updated in place with the same incremental update as K2, which also keeps the ORDER of the threat lists
// identical to the reference's (swap-with-last removal, append on change, visiting order of update_around).
#pragma once
#include "solver_logic.cuh"

#ifndef AGB_SOLVER_NS
#define AGB_SOLVER_NS solver // solver_plain in the build without renju (solver_kernel.cuh)
#endif
namespace agb
{
	namespace AGB_SOLVER_NS
	{
		// ---- mutable position ----------------------------------------------------------------------------------------------------
		struct DynState
		{
				View v; // read side (same memory); stm / stones follow add and undo
				AGB_POSITION_ARRAY(int8_t, kBoard) board; // on the device: this warp's shared-memory copy (solver_logic.cuh)
				AGB_POSITION_ARRAY(uint64_t, kLines) lines;
				AGB_POSITION_ARRAY(uint32_t, kPtypes) ptypes;
				AGB_POSITION_ARRAY(uint8_t, kThreats) threats;
				AGB_POSITION_ARRAY(int32_t, kHistCount) hist_count;
				AGB_POSITION_ARRAY(uint16_t, kListIndex) list_index; // device only: see position_layout
				uint16_t *hist_cells;
				const uint8_t *threat_table;
				// MoveGenerator::forbidden_moves_cache: results of isForbidden within one generate() call
				uint16_t cache_loc[48];
				uint8_t cache_val[48];
				int cache_size = 0;
				uint32_t overflow = 0; // 1: forbidden-move recursion too deep, 2: forbidden cache full
				uint32_t n_adds = 0, n_quiet = 0, n_gen = 0, n_gen_actions = 0; // work counters (scheduling weights, solver.cu)
		};

		// The solver never walks the HALF_OPEN_3 lists and never asks for their length (MoveGenerator.cpp reads a cell's HALF_OPEN_3 threat
		// only through threat_at / the pattern table; AlphaBetaSearch::evaluate gives the type weight 0), and the slot's lists die with the
		// launch, so the search keeps only the lists of OPEN_3 and stronger in step: the longest list of a mid-game position costs nothing.
		AGB_HD inline bool list_is_kept(int type) { return type > TT_HALF_OPEN_3; }
#ifdef __CUDA_ARCH__
		// All lanes of the game's warp run these with identical arguments (lockstep); every lane does all of it, so each lane only ever depends on
		// its own earlier loads and stores. A cell is on at most one list per colour, and d.list_index remembers where: ThreatHistogram::remove's
		// linear search (the reference finds the FIRST match; a cell occurs once) becomes one load.
		__device__ __forceinline__ int list_index_of(uint16_t loc, int colour) { return 2 * ((loc & 255) * kMaxSize + (loc >> 8)) + colour; }
		__device__ __forceinline__ void dyn_hist_remove(DynState &d, int colour, int type, uint16_t loc)
		{ // ThreatHistogram::remove (ThreatHistogram.hpp:74-91)
			if (not list_is_kept(type))
				return;
			int32_t &count = d.hist_count[colour * kHistTypes + type];
			uint16_t *list = d.hist_cells + (colour * kHistTypes + type) * kCellPitch;
			const int len = count;
			const int found = d.list_index[list_index_of(loc, colour)];
			if (found < len)
			{
				const uint16_t last = list[len - 1];
				list[found] = last;
				d.list_index[list_index_of(last, colour)] = static_cast<uint16_t>(found);
				count = len - 1;
			}
		}
		__device__ __forceinline__ void dyn_hist_add(DynState &d, int colour, int type, uint16_t loc)
		{
			if (not list_is_kept(type))
				return;
			int32_t &count = d.hist_count[colour * kHistTypes + type];
			const int len = count;
			d.hist_cells[(colour * kHistTypes + type) * kCellPitch + len] = loc;
			d.list_index[list_index_of(loc, colour)] = static_cast<uint16_t>(len);
			count = len + 1;
		}
		// the index of the lists as K1 left them (the kept ones): lanes take interleaved entries
		__device__ __forceinline__ void dyn_build_list_index(DynState &d)
		{
#pragma unroll 1
			for (int k = 0; k < 2 * kHistTypes; k++)
				if (list_is_kept(k % kHistTypes))
				{
					const int len = d.hist_count[k];
					for (int i = threadIdx.x & 31; i < len; i += 32)
						d.list_index[list_index_of(d.hist_cells[k * kCellPitch + i], k / kHistTypes)] = static_cast<uint16_t>(i);
				}
			__syncwarp();
		}
#else
		AGB_HD_NOINLINE inline void dyn_hist_remove(DynState &d, int colour, int type, uint16_t loc)
		{ // ThreatHistogram::remove (ThreatHistogram.hpp:74-91)
			if (not list_is_kept(type))
				return;
			int32_t &count = d.hist_count[colour * kHistTypes + type];
			uint16_t *list = d.hist_cells + (colour * kHistTypes + type) * d.v.pitch;
			for (int i = 0; i < count; i++)
				if (list[i] == loc)
				{
					list[i] = list[count - 1];
					count--;
					return;
				}
		}
		AGB_HD inline void dyn_hist_add(DynState &d, int colour, int type, uint16_t loc)
		{
			if (not list_is_kept(type))
				return;
			int32_t &count = d.hist_count[colour * kHistTypes + type];
			d.hist_cells[(colour * kHistTypes + type) * d.v.pitch + count] = loc;
			count++;
		}
#endif
#ifdef __CUDA_ARCH__
		// Device form of addMove / undoMove. All 32 lanes of the game's warp run the solver in lockstep on identical data; here they split the
		// work: lane q (and q + 32) owns entry q = 4 * k + dir of the 40 neighbour cells (k-th offset of -5..-1, 1..5, like K2), so a lane only
		// ever needs the line word of ITS direction (dir = lane & 3). The list changes are then replayed by the whole warp in the visiting
		// order of update_around (PatternCalculator.cpp:320-330), which is the order of q.
		__device__ __noinline__ void dyn_update_neighbours_warp(DynState &d, int S, int r, int c, int dir, uint64_t my_line, int my_pos)
		{
			const int lane = threadIdx.x & 31;
			// the state's pointers in registers: the byte stores below could alias the fields of `d` and would force a reload after each
			const int8_t *const board = d.board;
			uint32_t *const ptypes = d.ptypes;
			uint8_t *const threats = d.threats;
			const uint8_t *const pattern_table = d.v.pattern_table, *const threat_table = d.threat_table;
			const int row_step = dir_row_step(dir), col_step = dir_col_step(dir);
			// Two passes of 32 and 8 entries, NOT unrolled, each: look the entry's cell up, store it, then replay the list changes of the pass in entry
			// order. A cell appears once among the 40 entries and the list changes depend only on its (old, new) threats, so replaying pass 0 before
			// pass 1 is looked up gives what "all lookups, then all replays" gave -- in half the code, which is what this kernel is short of.
#pragma unroll 1
			for (int half = 0; half < 2; half++)
			{
				uint32_t ot = 0, nt = 0, lc = 0;
				const int q = lane + 32 * half; // q & 3 == dir in both halves
				if (q < 40)
				{
					const int k = q >> 2;
					const int off = (k < 5) ? (k - 5) : (k - 4);
					const int nr = r + off * row_step, nc = c + off * col_step;
					const int ncell = nr * S + nc;
					if (nr >= 0 and nr < S and nc >= 0 and nc < S and board[ncell] == NONE)
					{
						const uint32_t window = static_cast<uint32_t>(my_line >> (2 * (my_pos + off) + 2)) & 0x3FFFFFu;
						const uint32_t byte = pattern_table[narrow_window(window)];
						const uint32_t p = (ptypes[ncell] & ~(0xFFu << (8 * dir))) | (byte << (8 * dir));
						ot = threats[ncell];
						nt = threat_of_cell(p, threat_table);
						lc = mk_loc(nr, nc);
						ptypes[ncell] = p;
						threats[ncell] = static_cast<uint8_t>(nt);
					}
				}
				// only changes that touch a kept list (OPEN_3 or stronger, either colour) are replayed
				const bool touches = ((ot & 15u) != (nt & 15u) and (list_is_kept(ot & 15u) or list_is_kept(nt & 15u)))
						or ((ot >> 4) != (nt >> 4) and (list_is_kept(ot >> 4) or list_is_kept(nt >> 4)));
				unsigned changed = __ballot_sync(0xFFFFFFFFu, touches);
				while (changed)
				{
					const int src = __ffs(changed) - 1;
					changed &= changed - 1;
					// one shuffle: old threats | new threats << 8 | location << 16
					const uint32_t packed = __shfl_sync(0xFFFFFFFFu, ot | (nt << 8) | (lc << 16), src);
					const uint16_t l = static_cast<uint16_t>(packed >> 16);
#pragma unroll 1
					for (int colour = 0; colour < 2; colour++)
					{
						const int o = (packed >> (4 * colour)) & 15, nw = (packed >> (8 + 4 * colour)) & 15;
						if (o != nw)
						{
							if (list_is_kept(o))
								dyn_hist_remove(d, colour, o, l);
							dyn_hist_add(d, colour, nw, l);
						}
					}
				}
			}
			__syncwarp();
		}
		__device__ __noinline__ void dyn_add_move(DynState &d, int r, int c, int sign)
		{ // PatternCalculator::addMove (PatternCalculator.cpp:68-86)
			d.n_adds++;
			const int S = d.v.S;
			const int dir = threadIdx.x & 3;
			const int li = line_index(dir, r, c, S), pos = pos_in_line(dir, r, c, S);
			uint64_t *const lines = d.lines;
			const uint64_t line = lines[li] | (static_cast<uint64_t>(sign) << (12 + 2 * pos));
			__syncwarp();
			if ((threadIdx.x & 31) < 4)
				lines[li] = line;
			const int centre = r * S + c;
			const uint8_t old_t = d.threats[centre];
			__syncwarp();
			d.board[centre] = static_cast<int8_t>(sign);
			d.ptypes[centre] = 0;
			d.threats[centre] = 0;
			if (list_is_kept(old_t & 15))
				dyn_hist_remove(d, 0, old_t & 15, mk_loc(r, c));
			if (list_is_kept(old_t >> 4))
				dyn_hist_remove(d, 1, old_t >> 4, mk_loc(r, c));
			dyn_update_neighbours_warp(d, S, r, c, dir, line, pos);
			d.v.stm = 3 - d.v.stm;
			d.v.stones++;
		}
		__device__ __noinline__ void dyn_undo_move(DynState &d, int r, int c, int sign)
		{ // PatternCalculator::undoMove (PatternCalculator.cpp:87-106)
			(void) sign;
			const int S = d.v.S;
			const int dir = threadIdx.x & 3;
			const int li = line_index(dir, r, c, S), pos = pos_in_line(dir, r, c, S);
			uint64_t *const lines = d.lines;
			const uint64_t line = lines[li] & ~(3ull << (12 + 2 * pos));
			__syncwarp();
			if ((threadIdx.x & 31) < 4)
				lines[li] = line;
			const int centre = r * S + c;
			// the centre's four direction bytes: lanes 0..3 look one up each
			const uint32_t mine = ((threadIdx.x & 31) < 4)
					? static_cast<uint32_t>(d.v.pattern_table[narrow_window(static_cast<uint32_t>(line >> (2 * pos + 2)) & 0x3FFFFFu)]) << (8 * dir) : 0u;
			const uint32_t p = __reduce_or_sync(0xFFFFFFFFu, mine);
			const uint8_t new_t = threat_of_cell(p, d.threat_table);
			d.board[centre] = NONE;
			d.ptypes[centre] = p;
			d.threats[centre] = new_t;
			dyn_hist_add(d, 0, new_t & 15, mk_loc(r, c));
			dyn_hist_add(d, 1, new_t >> 4, mk_loc(r, c));
			dyn_update_neighbours_warp(d, S, r, c, dir, line, pos);
			d.v.stm = 3 - d.v.stm;
			d.v.stones--;
		}
#else
		inline void dyn_update_neighbours(DynState &d, int r, int c, const uint64_t *line4, const int *pos4)
		{ // the 40 cells at distance 1..5 in the visiting order of update_around (PatternCalculator.cpp:320-330)
			const int S = d.v.S;
			for (int off = -5; off <= 5; off++)
				if (off != 0)
					for (int dir = 0; dir < 4; dir++)
					{
						const int nr = r + off * dir_row_step(dir), nc = c + off * dir_col_step(dir);
						if (nr < 0 or nr >= S or nc < 0 or nc >= S)
							continue;
						const int ncell = nr * S + nc;
						if (d.board[ncell] != NONE)
							continue;
						const uint32_t window = static_cast<uint32_t>(line4[dir] >> (2 * (pos4[dir] + off) + 2)) & 0x3FFFFFu;
						const uint32_t byte = d.v.pattern_table[narrow_window(window)];
						const uint32_t p = (d.ptypes[ncell] & ~(0xFFu << (8 * dir))) | (byte << (8 * dir));
						const uint8_t old_t = d.threats[ncell];
						const uint8_t new_t = threat_of_cell(p, d.threat_table);
						d.ptypes[ncell] = p;
						d.threats[ncell] = new_t;
						if (old_t != new_t)
						{
							const uint16_t loc = mk_loc(nr, nc);
							if ((old_t & 15) != (new_t & 15))
							{
								dyn_hist_remove(d, 0, old_t & 15, loc);
								dyn_hist_add(d, 0, new_t & 15, loc);
							}
							if ((old_t >> 4) != (new_t >> 4))
							{
								dyn_hist_remove(d, 1, old_t >> 4, loc);
								dyn_hist_add(d, 1, new_t >> 4, loc);
							}
						}
					}
		}
		inline void dyn_add_move(DynState &d, int r, int c, int sign)
		{ // PatternCalculator::addMove (PatternCalculator.cpp:68-86)
			d.n_adds++;
			const int S = d.v.S;
			uint64_t line4[4];
			int pos4[4];
			for (int dir = 0; dir < 4; dir++)
			{
				const int li = line_index(dir, r, c, S);
				pos4[dir] = pos_in_line(dir, r, c, S);
				d.lines[li] |= static_cast<uint64_t>(sign) << (12 + 2 * pos4[dir]);
				line4[dir] = d.lines[li];
			}
			const int centre = r * S + c;
			d.board[centre] = static_cast<int8_t>(sign);
			const uint8_t old_t = d.threats[centre];
			dyn_hist_remove(d, 0, old_t & 15, mk_loc(r, c));
			dyn_hist_remove(d, 1, old_t >> 4, mk_loc(r, c));
			d.ptypes[centre] = 0;
			d.threats[centre] = 0;
			dyn_update_neighbours(d, r, c, line4, pos4);
			d.v.stm = 3 - d.v.stm;
			d.v.stones++;
		}
		inline void dyn_undo_move(DynState &d, int r, int c, int sign)
		{ // PatternCalculator::undoMove (PatternCalculator.cpp:87-106)
			(void) sign;
			const int S = d.v.S;
			uint64_t line4[4];
			int pos4[4];
			for (int dir = 0; dir < 4; dir++)
			{
				const int li = line_index(dir, r, c, S);
				pos4[dir] = pos_in_line(dir, r, c, S);
				d.lines[li] &= ~(3ull << (12 + 2 * pos4[dir]));
				line4[dir] = d.lines[li];
			}
			const int centre = r * S + c;
			d.board[centre] = NONE;
			uint32_t p = 0;
			for (int dir = 0; dir < 4; dir++)
				p |= static_cast<uint32_t>(d.v.pattern_table[narrow_window(static_cast<uint32_t>(line4[dir] >> (2 * pos4[dir] + 2)) & 0x3FFFFFu)]) << (8 * dir);
			const uint8_t new_t = threat_of_cell(p, d.threat_table);
			d.ptypes[centre] = p;
			d.threats[centre] = new_t;
			dyn_hist_add(d, 0, new_t & 15, mk_loc(r, c));
			dyn_hist_add(d, 1, new_t >> 4, mk_loc(r, c));
			dyn_update_neighbours(d, r, c, line4, pos4);
			d.v.stm = 3 - d.v.stm;
			d.v.stones--;
		}
#endif

		// ---- renju forbidden moves on the live state (with the reference's side effect on the list order) ---------------------------
		constexpr int kMaxForbiddenDepth = 8;
		AGB_HD inline bool dyn_calc_is_forbidden(DynState &d, int r, int c, int depth);
		AGB_HD inline bool dyn_raw_straight_four(const DynState &d, int r, int c, int dir)
		{ // RawPatternCalculator::isStraightFourAt on the raw board
			Overlay none;
			return makes_straight_four(raw_window(d.board, d.v.S, r, c, dir, none));
		}
		AGB_HD_NOINLINE inline bool dyn_is_3x3_forbidden(DynState &d, int r, int c, int depth)
		{ // PatternCalculator::is_3x3_forbidden (PatternCalculator.cpp:213-244)
			if (depth >= kMaxForbiddenDepth)
			{
				d.overflow |= 1u;
				return true;
			}
			const int S = d.v.S;
			int open3_count = 0;
			for (int dir = 0; dir < 4; dir++)
				if (d.v.ptype_at(CROSS, r, c, dir) == PT_OPEN_3)
				{
					const uint32_t promotions = open_three_promotions(normal_window(d.lines, dir, r, c, S));
					d.board[r * S + c] = CROSS;
					for (int i = -5; i <= 5; i++)
						if ((promotions >> (i + 5)) & 1u)
						{
							const int pr = r + i * dir_row_step(dir), pc = c + i * dir_col_step(dir);
							if (pr < 0 or pr >= S or pc < 0 or pc >= S)
								continue;
							if (d.board[pr * S + pc] == NONE and dyn_raw_straight_four(d, pr, pc, dir))
							{
								d.board[r * S + c] = NONE;
								const int stm = d.v.stm;
								dyn_add_move(d, r, c, CROSS);
								const bool forbidden = dyn_calc_is_forbidden(d, pr, pc, depth + 1);
								dyn_undo_move(d, r, c, CROSS);
								d.v.stm = stm;
								d.board[r * S + c] = CROSS;
								if (not forbidden)
								{
									open3_count++;
									break;
								}
							}
						}
					d.board[r * S + c] = NONE;
				}
			return open3_count >= 2;
		}
		AGB_HD inline bool dyn_calc_is_forbidden(DynState &d, int r, int c, int depth)
		{ // PatternCalculator::isForbidden(CROSS, r, c) (PatternCalculator.hpp:173-189)
#ifdef AGB_SOLVER_NO_RENJU
			return false;
#endif
			if (d.v.rules != RULE_RENJU)
				return false;
			if (d.board[r * d.v.S + c] != NONE)
				return false;
			const int t = d.threats[r * d.v.S + c] & 15;
			if (t == TT_OVERLINE or t == TT_FORK_4x4)
				return true;
			if (t == TT_FORK_3x3)
				return dyn_is_3x3_forbidden(d, r, c, depth);
			return false;
		}
		AGB_HD_NOINLINE inline bool dyn_is_forbidden(DynState *d, int sign, int r, int c)
		{ // MoveGenerator::is_forbidden (MoveGenerator.cpp:1167-1180): cached per generate() call
#ifdef AGB_SOLVER_NO_RENJU
			return false;
#endif
			if (not (d->v.rules == RULE_RENJU and sign == CROSS))
				return false;
			const uint16_t loc = mk_loc(r, c);
			for (int i = 0; i < d->cache_size; i++)
				if (d->cache_loc[i] == loc)
					return d->cache_val[i] != 0;
			const bool result = dyn_calc_is_forbidden(*d, r, c, 0);
			if (d->cache_size < 48)
			{
				d->cache_loc[d->cache_size] = loc;
				d->cache_val[d->cache_size] = result ? 1 : 0;
				d->cache_size++;
			}
			else
				d->overflow |= 2u;
			return result;
		}

		// NNInputFeatures::encode (NNInputFeatures.cpp:104-110) asks isForbidden for every cell before the search starts; only its side
		// effect on the list order matters here (the feature words themselves come from K3)
		AGB_HD_NOINLINE inline void encode_forbidden_pass(DynState &d)
		{
#ifdef AGB_SOLVER_NO_RENJU
			return;
#endif
			if (d.v.rules != RULE_RENJU or d.v.stm != CROSS)
				return;
			const int n = d.v.count(CROSS, TT_FORK_3x3);
			if (n == 0)
				return; // nothing would be added or undone
			for (int r = 0; r < d.v.S; r++)
				for (int c = 0; c < d.v.S; c++)
					if ((d.threats[r * d.v.S + c] & 15) == TT_FORK_3x3)
						dyn_calc_is_forbidden(d, r, c, 0);
		}

		// ---- transposition table (one per game) --------------------------------------------------------------------------------------
		enum : int { BOUND_NONE = 0, BOUND_LOWER = 1, BOUND_UPPER = 2, BOUND_EXACT = 3 };
		struct HashTable
		{
				uint64_t *entries; // [size][2]: upper key word, packed data (SharedTableData)
				uint64_t bucket_mask; // buckets of 4 entries
				int generation; // SharedHashTable::m_base_generation
				const uint64_t *keys; // [2 * cells][2]: low, high word per (cell, colour)
		};
		constexpr uint64_t kKeyMask = 0xFFFF000000000000ull;
		constexpr uint64_t kEmptyEntryData = static_cast<uint64_t>(kScoreDefault) << 16; // SharedTableData(): bound NONE, depth 0, Score(), Move()
		AGB_HD inline uint64_t tt_pack(int bound, int depth, uint16_t score, uint16_t move)
		{
			return static_cast<uint64_t>(bound) | (static_cast<uint64_t>(depth) << 8) | (static_cast<uint64_t>(score) << 16) | (static_cast<uint64_t>(move) << 32);
		}
		AGB_HD inline int tt_bound(uint64_t e) { return static_cast<int>(e & 3ull); }
		AGB_HD inline int tt_generation(uint64_t e) { return static_cast<int>((e >> 2) & 63ull); }
		AGB_HD inline int tt_depth(uint64_t e) { return static_cast<int>((e >> 8) & 255ull); }
		AGB_HD inline uint16_t tt_score(uint64_t e) { return static_cast<uint16_t>((e >> 16) & 65535ull); }
		AGB_HD inline uint16_t tt_move(uint64_t e) { return static_cast<uint16_t>((e >> 32) & 65535ull); }
		AGB_HD inline void tt_clear(uint64_t *entries, size_t n_entries)
		{
			for (size_t i = 0; i < n_entries; i++)
			{
				entries[2 * i] = 0;
				entries[2 * i + 1] = kEmptyEntryData;
			}
		}
		AGB_HD inline void tt_prefetch(const HashTable &t, uint64_t lo)
		{ // SharedHashTable::prefetch (SharedHashTable.hpp:183-186): the bucket is on its way while the move is being added
#ifdef __CUDA_ARCH__
			asm volatile("prefetch.global.L2 [%0];" :: "l"(t.entries + 8 * (lo & t.bucket_mask)));
#else
			(void) t;
			(void) lo;
#endif
		}
		AGB_HD_NOINLINE inline uint64_t tt_seek(const HashTable &t, uint64_t lo, uint64_t hi)
		{
			const uint64_t *bucket = t.entries + 8 * (lo & t.bucket_mask);
			for (int i = 0; i < 4; i++)
				if (bucket[2 * i] == hi and (bucket[2 * i + 1] & kKeyMask) == (lo & kKeyMask))
					return bucket[2 * i + 1];
			return kEmptyEntryData;
		}
		AGB_HD_NOINLINE inline void tt_insert(HashTable &t, uint64_t lo, uint64_t hi, uint64_t value)
		{
			value &= ~(kKeyMask | 0xFCull);
			value |= static_cast<uint64_t>(t.generation) << 2;
			value |= lo & kKeyMask;
			uint64_t *bucket = t.entries + 8 * (lo & t.bucket_mask);
			if (sc_is_proven(tt_score(value)) or tt_bound(value) == BOUND_EXACT)
				for (int i = 0; i < 4; i++)
					if (bucket[2 * i] == hi and (bucket[2 * i + 1] & kKeyMask) == (lo & kKeyMask))
					{
						bucket[2 * i] = hi;
						bucket[2 * i + 1] = value;
						return;
					}
			int idx = 0, best = 0;
			for (int i = 0; i < 4; i++)
			{
				const int val = tt_depth(bucket[2 * i + 1]) - (t.generation - tt_generation(bucket[2 * i + 1]));
				if (i == 0 or val < best)
				{
					best = val;
					idx = i;
				}
			}
			bucket[2 * idx] = hi;
			bucket[2 * idx + 1] = value;
		}

		// ---- Score arithmetic of the search (search/Score.hpp) -----------------------------------------------------------------------
		constexpr uint16_t kScoreMinusInf = 0x0000, kScorePlusInf = 0xFFFF;
		AGB_HD inline int sc_pv(uint16_t s) { return (s >> 13) & 3; }
		AGB_HD inline int sc_ev(uint16_t s) { return static_cast<int>(s & 8191) - 4000; }
		AGB_HD inline bool sc_finite(uint16_t s) { return s != kScoreMinusInf and s != kScorePlusInf; }
		AGB_HD inline bool sc_is_loss(uint16_t s) { return sc_pv(s) == PV_LOSS and sc_finite(s); }
		AGB_HD inline int sc_distance(uint16_t s)
		{
			switch (sc_pv(s))
			{
				case PV_LOSS:
				case PV_DRAW:
					return sc_ev(s);
				case PV_WIN:
					return -sc_ev(s);
				default:
					return 0;
			}
		}
		AGB_HD inline uint16_t sc_negate(uint16_t s)
		{
			switch (sc_pv(s))
			{
				case PV_LOSS:
					return sc_finite(s) ? mk_score(PV_WIN, -sc_ev(s)) : kScorePlusInf;
				case PV_DRAW:
					return mk_score(PV_DRAW, sc_ev(s));
				case PV_WIN:
					return sc_finite(s) ? mk_score(PV_LOSS, -sc_ev(s)) : kScoreMinusInf;
				default:
					return mk_score(PV_UNKNOWN, -sc_ev(s));
			}
		}
		AGB_HD inline uint16_t sc_invert(uint16_t s, int delta)
		{ // invert_up (delta +1) / invert_down (delta -1) (Score.hpp): a loss in n becomes a win in n + delta and the other way round, a draw in n a
		  // draw in n + delta, an evaluation is negated, the infinities swap. Branch-free: the function sits in the hottest loop of a kernel that is
		  // bound by instruction supply, and its switch-of-switches form (equal on all 65 536 scores) was inlined five times.
			const int pv = (s >> 13) & 3, raw = s & 8191; // raw = 4000 + eval
			const int outer = ((pv ^ (pv >> 1)) & 1) ^ 1; // LOSS or WIN: the proven value flips
			const int mirrored = 8000 - raw + ((pv == PV_LOSS) ? -delta : ((pv == PV_WIN) ? delta : 0));
			const int value = (pv == PV_DRAW) ? raw + delta : mirrored;
			const uint16_t r = static_cast<uint16_t>(((pv ^ (outer ? 3 : 0)) << 13) | value);
			return (s == kScoreMinusInf) ? kScorePlusInf : ((s == kScorePlusInf) ? kScoreMinusInf : r);
		}

		// ---- the search --------------------------------------------------------------------------------------------------------------
		constexpr int kMaxFrames = 104; // depth limit 100 (Search::solve sets it) + root
		struct Frame
		{
				int32_t list_begin; // offset of this node's actions on the action stack
				int16_t list_size;
				int16_t index; // action being searched
				int16_t depth_remaining;
				uint16_t alpha, beta, original_alpha, best_score, best_move;
				uint16_t move; // move made to descend from this node
				uint8_t is_fully_expanded;
				uint8_t pad;
		};
		// What playing one root action would do to the position, worked out without touching the state (see preview_children)
		constexpr int kMaxChildOps = 14;
		struct ChildInfo
		{
				uint16_t eval; // AlphaBetaSearch::evaluate() of the position after the move, from its side to move
				uint8_t quiet; // 1: afterwards neither side has a 3x3 fork or anything stronger (MoveGenerator THREATS stages generate nothing)
				uint8_t n_ops; // 255: more open-three list changes than fit below, the move takes the regular path
				uint32_t ops[kMaxChildOps]; // cells entering or leaving an OPEN_3 list, in update order: loc | old threats << 16 | new threats << 24
		};
		struct SearchMemory
		{ // scratch of one game (global memory on the device)
				uint16_t *stack_moves, *stack_scores; // ActionStack
				int stack_capacity;
				Frame *frames; // [kMaxFrames]
				ChildInfo *children; // [cells], indexed by the cell of the root action
		};
		struct SearchOutput
		{
				uint16_t score = kScoreDefault;
				int n_actions = 0; // root actions are left at the bottom of the action stack, in their final order
				bool must_defend = false;
				int node_counter = 0;
				uint32_t overflow = 0; // 4: action stack full, 8: frame stack full (plus DynState's bits)
		};

		AGB_HD inline void hash_toggle(const HashTable &t, uint64_t &lo, uint64_t &hi, int S, uint16_t move)
		{ // FastZobristHashing::updateHash
			const int row = (move >> 2) & 127, col = (move >> 9) & 127, sign = move & 3;
			const uint64_t *k = t.keys + 2 * (2 * (row * S + col) + sign - 1);
			lo ^= k[0];
			hi ^= k[1];
		}
		AGB_HD_NOINLINE inline void hash_of_board(const HashTable &t, const DynState &d, uint64_t &lo, uint64_t &hi)
		{ // FastZobristHashing::getHash
			lo = 0;
			hi = 0;
#ifdef __CUDA_ARCH__
#pragma unroll 1
			for (int i = threadIdx.x & 31; i < d.v.cells; i += 32) // the lanes of the lockstep warp take interleaved cells
#else
			for (int i = 0; i < d.v.cells; i++)
#endif
				if (d.board[i] == CROSS or d.board[i] == CIRCLE)
				{
					const uint64_t *k = t.keys + 2 * (2 * i + d.board[i] - 1);
					lo ^= k[0];
					hi ^= k[1];
				}
#ifdef __CUDA_ARCH__
			for (int o = 16; o > 0; o >>= 1)
			{
				lo ^= __shfl_xor_sync(0xFFFFFFFFu, lo, o);
				hi ^= __shfl_xor_sync(0xFFFFFFFFu, hi, o);
			}
#endif
		}

		// ---- quiet root children without add/undo -------------------------------------------------------------------------------------
		// In a calm position (nobody has a 3x3 fork or better) nearly every root action leads to a position where the THREATS-mode
		// generator finds nothing, so the reference's 100-node budget is spent on "add the stone, see that nothing is there, evaluate,
		// take it back" (AlphaBetaSearch.cpp:262-339 with an empty child list). What such a visit leaves behind is fully determined by
		// (a) the threat histogram of the child position -> quiet or not, evaluate(); (b) the transposition table traffic; (c) the new
		// ORDER of the two OPEN_3 lists after addMove + undoMove (the only lists a later generate() reads that a quiet move can touch;
		// HALF_OPEN_3 lists are never read). (a) and the list edits of (c) are computed for all root actions at once, one action per
		// lane, straight from the line words; the visit itself then needs no pattern update at all.
		AGB_HD inline bool is_calm(const View &v)
		{
			for (int colour = 0; colour < 2; colour++)
				for (int t = TT_FORK_3x3; t < kHistTypes; t++)
					if (v.hist_count[colour * kHistTypes + t] != 0)
						return false;
			return true;
		}
		namespace eval_tables
		{ // AlphaBetaSearch::evaluate's weights per ThreatType (AlphaBetaSearch.cpp:345-365), for the side to move and for the other side
			AGB_MASK_TABLE int16_t own[kHistTypes] = { 0, 0, 19, 49, 76, 170, 33, 159, 252, 0 };
			AGB_MASK_TABLE int16_t other[kHistTypes] = { 0, 0, -1, -50, -45, -135, -14, -154, -496, 0 };
		}
		AGB_HD inline int eval_weight(int type, bool own)
		{
			return own ? eval_tables::own[type] : eval_tables::other[type];
		}
		AGB_HD inline void preview_child(const DynState &d, uint16_t move, ChildInfo &out)
		{ // valid for a calm parent only (the strong lists are empty, so "quiet" is "no strong threat appears")
			const int S = d.v.S;
			const int r = (move >> 2) & 127, c = (move >> 9) & 127, sign = move & 3;
			const int child_own = 3 - sign; // side to move after the move
			// evaluate() of the parent's histogram seen from the child's side to move, then corrected cell by cell
			int sum = 12;
#pragma unroll 1
			for (int t = TT_OPEN_3; t <= TT_FIVE; t++)
				sum += eval_weight(t, true) * d.v.count(child_own, t) + eval_weight(t, false) * d.v.count(3 - child_own, t);
			int strong = 0, n_ops = 0;
			const uint8_t centre_t = d.threats[r * S + c];
			for (int colour = 0; colour < 2; colour++)
				sum -= eval_weight((centre_t >> (4 * colour)) & 15, colour + 1 == child_own);
			if ((centre_t & 15) == TT_OPEN_3 or (centre_t >> 4) == TT_OPEN_3)
				out.ops[n_ops++] = mk_loc(r, c) | (static_cast<uint32_t>(centre_t) << 16);
#pragma unroll 1
			for (int off = -5; off <= 5; off++)
				if (off != 0)
#pragma unroll 1
					for (int dir = 0; dir < 4; dir++)
					{
						const int nr = r + off * dir_row_step(dir), nc = c + off * dir_col_step(dir);
						if (nr < 0 or nr >= S or nc < 0 or nc >= S)
							continue;
						const int ncell = nr * S + nc;
						if (d.board[ncell] != NONE)
							continue;
						const int pos = pos_in_line(dir, r, c, S);
						const uint64_t line = d.lines[line_index(dir, r, c, S)] | (static_cast<uint64_t>(sign) << (12 + 2 * pos));
						const uint32_t window = static_cast<uint32_t>(line >> (2 * (pos + off) + 2)) & 0x3FFFFFu;
						const uint32_t byte = d.v.pattern_table[narrow_window(window)];
						const uint32_t p = (d.ptypes[ncell] & ~(0xFFu << (8 * dir))) | (byte << (8 * dir));
						const uint8_t old_t = d.threats[ncell];
						const uint8_t new_t = threat_of_cell(p, d.threat_table);
						if (old_t == new_t)
							continue;
						bool touches_open3 = false;
#pragma unroll 1
						for (int colour = 0; colour < 2; colour++)
						{
							const int o = (old_t >> (4 * colour)) & 15, n = (new_t >> (4 * colour)) & 15;
							if (o == n)
								continue;
							const bool own = (colour + 1 == child_own);
							sum += eval_weight(n, own) - eval_weight(o, own);
							strong += (n >= TT_FORK_3x3) ? 1 : 0; // o < FORK_3x3 in a calm parent
							touches_open3 = touches_open3 or o == TT_OPEN_3 or n == TT_OPEN_3;
						}
						if (touches_open3)
						{
							if (n_ops < kMaxChildOps)
								out.ops[n_ops] = mk_loc(nr, nc) | (static_cast<uint32_t>(old_t) << 16) | (static_cast<uint32_t>(new_t) << 24);
							n_ops++;
						}
					}
			sum = sum < -1000 ? -1000 : (sum > 1000 ? 1000 : sum);
			out.eval = sc_eval(sum);
			out.quiet = (strong == 0) ? 1 : 0;
			out.n_ops = (n_ops <= kMaxChildOps) ? static_cast<uint8_t>(n_ops) : 255;
		}
		AGB_HD_NOINLINE inline void preview_children(const DynState &d, const uint16_t *moves, int n, ChildInfo *children)
		{
#ifdef __CUDA_ARCH__
			for (int j = threadIdx.x & 31; j < n; j += 32) // lockstep warp: one root action per lane
#else
			for (int j = 0; j < n; j++)
#endif
			{
				const uint16_t mv = moves[j];
				preview_child(d, mv, children[((mv >> 2) & 127) * d.v.S + ((mv >> 9) & 127)]);
			}
#ifdef __CUDA_ARCH__
			__syncwarp();
#endif
		}
		// the edits addMove + undoMove of this action make to the OPEN_3 lists, in the reference's order (PatternCalculator.cpp:278-366)
		AGB_HD_NOINLINE inline void replay_open3_edits(DynState &d, uint16_t move, const ChildInfo &info)
		{
			const uint16_t centre = mk_loc((move >> 2) & 127, (move >> 9) & 127);
#pragma unroll 1
			for (int phase = 0; phase < 2; phase++) // 0: addMove, 1: undoMove
#pragma unroll 1
				for (int k = 0; k < info.n_ops; k++)
				{
					const uint32_t op = info.ops[k];
					const uint16_t loc = static_cast<uint16_t>(op & 0xFFFFu);
					const int from = (phase == 0) ? (op >> 16) & 255 : (op >> 24) & 255;
					const int to = (phase == 0) ? (op >> 24) & 255 : (op >> 16) & 255;
#pragma unroll 1
					for (int colour = 0; colour < 2; colour++)
					{
						const int o = (from >> (4 * colour)) & 15, n = (to >> (4 * colour)) & 15;
						if (o == n)
							continue;
						if (o == TT_OPEN_3 and not (phase == 1 and loc == centre)) // undoMove only appends the centre
							dyn_hist_remove(d, colour, TT_OPEN_3, loc);
						if (n == TT_OPEN_3 and not (phase == 0 and loc == centre)) // addMove only removes the centre
							dyn_hist_add(d, colour, TT_OPEN_3, loc);
					}
				}
		}

		// One visit of a quiet child of a calm root, step by step what recursive_solve does for it (AlphaBetaSearch.cpp:185-339 with an
		// empty action list): table probe, node count, evaluate(), table store; plus the list edits of the add / undo pair around it.
		// Returns false (nothing done) when the child is not quiet and must take the regular path.
		AGB_HD_NOINLINE inline bool quiet_child_visit(DynState &d, HashTable &tt, SearchMemory &mem, Frame &f, uint16_t *am, uint16_t *as, int i, uint64_t &key_lo,
				uint64_t &key_hi, bool &children_previewed, SearchOutput &out)
		{
			if (not children_previewed)
			{
				preview_children(d, am, f.list_size, mem.children);
				children_previewed = true;
			}
			const uint16_t mv = am[i];
			const ChildInfo &info = mem.children[((mv >> 2) & 127) * d.v.S + ((mv >> 9) & 127)];
			if (info.quiet == 0 or info.n_ops == 255)
				return false;
			d.n_quiet++;
			const int depth = f.depth_remaining - 1;
			const uint16_t alpha = sc_invert(f.beta, -1), beta = sc_invert(f.alpha, -1);
			hash_toggle(tt, key_lo, key_hi, d.v.S, mv);
			uint16_t best_move = 0, returned = kScoreDefault;
			bool done = false;
			const uint64_t entry = tt_seek(tt, key_lo, key_hi);
			if (tt_bound(entry) != BOUND_NONE)
			{
				best_move = tt_move(entry);
				const uint16_t s = tt_score(entry);
				const int b = tt_bound(entry);
				if (sc_is_proven(s) or (tt_depth(entry) >= depth and (b == BOUND_EXACT or (b == BOUND_LOWER and s >= beta) or (b == BOUND_UPPER and s <= alpha))))
				{
					returned = s;
					done = true;
				}
			}
			if (not done)
			{
				out.node_counter++;
				returned = info.eval;
				if (depth > 0)
				{
					const int bound = (returned <= alpha) ? BOUND_UPPER : ((returned >= beta) ? BOUND_LOWER : BOUND_EXACT);
					tt_insert(tt, key_lo, key_hi, tt_pack(bound, depth, returned, best_move));
				}
			}
			hash_toggle(tt, key_lo, key_hi, d.v.S, mv);
			replay_open3_edits(d, mv, info);
			as[i] = sc_invert(returned, +1);
			return true;
		}

		// AlphaBetaSearch::solve (AlphaBetaSearch.cpp:77-156) without the time limit; `max_depth` is Search::solve's 100
		AGB_HD inline SearchOutput solve_position(DynState &d, HashTable &tt, SearchMemory &mem, int max_nodes, int max_depth)
		{
			SearchOutput out;
			uint64_t key_lo, key_hi;
			hash_of_board(tt, d, key_lo, key_hi);
			int stack_offset = 0, stack_max_offset = 0; // ActionStack::m_offset / m_max_offset
			int root_size = 0;
			bool root_fully_expanded = false;
			uint16_t result = kScoreDefault;
			Frame *frames = mem.frames;
			bool root_calm = false, children_previewed = false;

			for (int depth = 0; depth <= max_depth; depth += 4)
			{
				const int max_offset_before = stack_max_offset;
				// ---- recursive_solve(depth, min_value, max_value, root actions), unrolled -------------------------------------------
				int top = 0;
				frames[0].list_begin = 0;
				frames[0].list_size = static_cast<int16_t>(root_size);
				frames[0].depth_remaining = static_cast<int16_t>(depth);
				frames[0].alpha = kScoreMinusInf;
				frames[0].beta = kScorePlusInf;
				frames[0].is_fully_expanded = root_fully_expanded ? 1 : 0;
				bool entering = true;
				uint16_t returned = kScoreDefault;
				while (top >= 0)
				{
					Frame &f = frames[top];
					if (entering)
					{ // AlphaBetaSearch.cpp:197-245
						entering = false;
						f.best_move = 0; // Move()
						bool done = false;
						const uint64_t entry = tt_seek(tt, key_lo, key_hi);
						if (tt_bound(entry) != BOUND_NONE)
						{
							f.best_move = tt_move(entry);
							if (top > 0)
							{
								const uint16_t s = tt_score(entry);
								const int b = tt_bound(entry);
								if (sc_is_proven(s))
								{
									returned = s;
									done = true;
								}
								else if (tt_depth(entry) >= f.depth_remaining
										and (b == BOUND_EXACT or (b == BOUND_LOWER and s >= f.beta) or (b == BOUND_UPPER and s <= f.alpha)))
								{
									returned = s;
									done = true;
								}
							}
						}
						if (not done)
						{
							out.node_counter++;
							if (f.list_size == 0)
							{
								if (stack_offset + d.v.cells > mem.stack_capacity)
								{
									out.overflow |= 4u;
									returned = kScoreDefault;
									done = true;
								}
								else
								{
									d.cache_size = 0;
									MoveGenerator gen(d.v, mem.stack_moves + f.list_begin, mem.stack_scores + f.list_begin);
									gen.generate(top == 0 ? GEN_OPTIMAL : GEN_THREATS);
									d.n_gen++;
									d.n_gen_actions += gen.out.n_actions;
									f.list_size = static_cast<int16_t>(gen.out.n_actions);
									f.is_fully_expanded = gen.out.is_fully_expanded ? 1 : 0;
									stack_offset += gen.out.n_actions;
									if (stack_offset > stack_max_offset)
										stack_max_offset = stack_offset;
									if (top == 0)
									{
										root_size = gen.out.n_actions;
										root_fully_expanded = gen.out.is_fully_expanded;
										out.must_defend = gen.out.must_defend;
										root_calm = is_calm(d.v);
									}
									if (sc_is_proven(gen.out.score))
									{
										returned = gen.out.score;
										done = true;
									}
								}
							}
						}
						if (not done and f.depth_remaining <= 0)
						{
							returned = static_evaluation(d.v);
							done = true;
						}
						if (done)
						{ // return from this call: the child's list dies (ActionList::~ActionList)
							if (top > 0)
								stack_offset -= f.list_size;
							top--;
							continue;
						}
						f.original_alpha = f.alpha;
						f.best_score = kScoreMinusInf;
						f.index = 0;
					}
					else
					{ // back from the child of action f.index (AlphaBetaSearch.cpp:293-310)
						const int i = f.index;
						mem.stack_scores[f.list_begin + i] = sc_invert(returned, +1);
						const uint16_t mv = f.move;
						dyn_undo_move(d, (mv >> 2) & 127, (mv >> 9) & 127, mv & 3);
						hash_toggle(tt, key_lo, key_hi, d.v.S, mv);
						// rest of the loop body for this action
						const uint16_t s = mem.stack_scores[f.list_begin + i];
						if (s > f.best_score)
							f.best_score = s;
						if (s > f.alpha)
						{
							f.alpha = s;
							f.best_move = mem.stack_moves[f.list_begin + i];
						}
						if (s >= f.beta or sc_is_win(s))
							f.index = f.list_size; // break
						else
							f.index++;
					}

					// the action loop (AlphaBetaSearch.cpp:262-316)
					uint16_t *am = mem.stack_moves + f.list_begin;
					uint16_t *as = mem.stack_scores + f.list_begin;
					bool descended = false;
					while (f.index < f.list_size)
					{
						const int i = f.index;
						bool ordered = false;
						if (i == 0)
						{ // is_move_legal(best_move) -> moveCloserToFront(best_move, 0)
							const uint16_t bm = f.best_move;
							const int br = (bm >> 2) & 127, bc = (bm >> 9) & 127;
							if ((bm & 3) == d.v.stm and br < d.v.S and bc < d.v.S and d.board[br * d.v.S + bc] == NONE)
							{
								ordered = true;
#pragma unroll 1
								for (int j = 0; j < f.list_size; j++)
									if (am[j] == bm)
									{
										const uint16_t tm = am[0], ts = as[0];
										am[0] = am[j];
										as[0] = as[j];
										am[j] = tm;
										as[j] = ts;
										break;
									}
							}
						}
						if (not ordered)
						{ // selection step: first maximum of the remaining actions
							int idx = i;
#ifdef __CUDA_ARCH__
							{ // the lanes of the warp scan interleaved slices; ties go to the lowest index, like the sequential scan
								const int lane = threadIdx.x & 31;
								uint32_t best = 0; // (score << 16) | (0xFFFF - index): larger score first, then smaller index
#pragma unroll 1
								for (int j = i + lane; j < f.list_size; j += 32)
								{
									const uint32_t key = (static_cast<uint32_t>(as[j]) << 16) | static_cast<uint32_t>(0xFFFF - j);
									best = key > best ? key : best;
								}
								for (int o = 16; o > 0; o >>= 1)
								{
									const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, best, o);
									best = other > best ? other : best;
								}
								idx = 0xFFFF - static_cast<int>(best & 0xFFFFu);
							}
#else
							for (int j = i + 1; j < f.list_size; j++)
								if (as[idx] < as[j])
									idx = j;
#endif
							const uint16_t tm = am[i], ts = as[i];
							am[i] = am[idx];
							as[i] = as[idx];
							am[idx] = tm;
							as[idx] = ts;
						}
						if (sc_pv(as[i]) == PV_UNKNOWN and out.node_counter < max_nodes)
						{ // descend
							if (top + 1 >= kMaxFrames)
							{
								out.overflow |= 8u;
							}
							else if (top == 0 and root_calm and mem.children != nullptr and d.v.draw_after - (d.v.stones + 1) >= 2
									and quiet_child_visit(d, tt, mem, f, am, as, i, key_lo, key_hi, children_previewed, out))
							{ // the visit was carried out without touching the position; as[i] holds the child's score
							}
							else
							{
								const uint16_t mv = am[i];
								hash_toggle(tt, key_lo, key_hi, d.v.S, mv);
								tt_prefetch(tt, key_lo);
								Frame &child = frames[top + 1];
								child.list_begin = stack_offset;
								child.list_size = 0;
								child.depth_remaining = static_cast<int16_t>(f.depth_remaining - 1);
								child.alpha = sc_invert(f.beta, -1);
								child.beta = sc_invert(f.alpha, -1);
								child.is_fully_expanded = 0;
								f.move = mv;
								dyn_add_move(d, (mv >> 2) & 127, (mv >> 9) & 127, mv & 3);
								top++;
								entering = true;
								descended = true;
								break;
							}
						}
						const uint16_t s = as[i];
						if (s > f.best_score)
							f.best_score = s;
						if (s > f.alpha)
						{
							f.alpha = s;
							f.best_move = am[i];
						}
						if (s >= f.beta or sc_is_win(s))
							break;
						f.index++;
					}
					if (descended)
						continue;

					// after the loop (AlphaBetaSearch.cpp:317-339)
					if (f.list_size == 0 or (sc_is_loss(f.best_score) and not f.is_fully_expanded))
						f.best_score = static_evaluation(d.v);
					int bound;
					if (f.best_score <= f.original_alpha)
						bound = BOUND_UPPER;
					else
						bound = (f.best_score >= f.beta) ? BOUND_LOWER : BOUND_EXACT;
					tt_insert(tt, key_lo, key_hi, tt_pack(bound, f.depth_remaining, f.best_score, f.best_move));
					returned = f.best_score;
					if (top > 0)
						stack_offset -= f.list_size;
					top--;
				}
				result = returned;
				if (root_size == 0 or sc_is_proven(result) or out.node_counter >= max_nodes or stack_max_offset == max_offset_before)
					break;
			}
			out.score = result;
			out.n_actions = root_size;
			out.overflow |= d.overflow;
			return out;
		}
	}
}
