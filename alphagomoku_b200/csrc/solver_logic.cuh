// K5 (static layer): the per-leaf move generator and static evaluation of the reference's solver, as host+device code.
// One thread analyses one position on top of the K1 state (pattern types, threats, ordered threat lists, line words).
//
// Reference: MoveGenerator::generate and its stages (src/search/alpha_beta/MoveGenerator.cpp:159-223, 356-1137),
// PatternCalculator::getDefensiveMoves (include/alphagomoku/patterns/PatternCalculator.hpp:162-172),
// DefensiveMoveTable::getMoves and the table construction (src/patterns/DefensiveMoveTable.cpp:58-216, 393-586),
// AlphaBetaSearch::evaluate (src/search/alpha_beta/AlphaBetaSearch.cpp:345-365).
// With TSSConfig::max_positions <= 1 the reference's solve() is exactly: generate(OPTIMAL) at the root, then either the
// proven static score or evaluate() (AlphaBetaSearch.cpp:77-156, 185-245); no transposition-table access happens.
//
// Order matters everywhere (the action list becomes the edge list of the search node), so list operations keep the
// reference's semantics: swap-with-last removal, first-match lookups, append-if-absent unions.
#pragma once
#include "patterns_logic.cuh"

#ifndef AGB_SOLVER_NS
#define AGB_SOLVER_NS solver // solver_plain in the build without renju (solver_kernel.cuh)
#endif
namespace agb
{
	namespace AGB_SOLVER_NS
	{
		using namespace plogic;

		// ---- Score helpers (search/Score.hpp) --------------------------------------------------------------------
		enum : int { PV_LOSS = 0, PV_DRAW = 1, PV_UNKNOWN = 2, PV_WIN = 3 };
		AGB_HD inline uint16_t mk_score(int pv, int eval) { return static_cast<uint16_t>((pv << 13) | (4000 + eval)); }
		AGB_HD inline uint16_t sc_eval(int e) { return mk_score(PV_UNKNOWN, e); }
		AGB_HD inline uint16_t win_in(int n) { return mk_score(PV_WIN, -n); }
		AGB_HD inline uint16_t loss_in(int n) { return mk_score(PV_LOSS, n); }
		AGB_HD inline uint16_t draw_in(int n) { return mk_score(PV_DRAW, n); }
		AGB_HD inline bool sc_is_proven(uint16_t s) { return ((s >> 13) & 3) != PV_UNKNOWN and s != 0x0000 and s != 0xFFFF; }
		AGB_HD inline bool sc_is_win(uint16_t s) { return ((s >> 13) & 3) == PV_WIN and s != 0xFFFF; }
		constexpr uint16_t kScoreDefault = (PV_UNKNOWN << 13) | 4000;
		constexpr uint16_t kScoreMin = 0x0000;

		// ---- defensive move tables (built on the host at engine creation, 15 x 256 x 2 masks per rule) ----------------------
		// layout: [group 0..14][neighbour index 0..255][defender colour 0 cross / 1 circle]; groups 0-4 five, 5-8 open four,
		// 9-14 double four
		constexpr int kDefGroups = 15;
		// stone patterns (as 2-bit-per-cell words) that identify where the threat sits in the 13-cell window, for cross as attacker. The tables
		// live in constant memory on the device (a function-local array would be rebuilt on the stack by every call)
#ifdef __CUDA_ARCH__
#define AGB_MASK_TABLE static __constant__ const
#else
#define AGB_MASK_TABLE static const
#endif
		namespace mask_tables
		{
			AGB_MASK_TABLE uint32_t five[5] = { 85u, 277u, 325u, 337u, 340u };
			AGB_MASK_TABLE uint32_t open_four[4] = { 84u, 276u, 324u, 336u };
			AGB_MASK_TABLE uint32_t double_four[6] = { 4177u, 4369u, 4417u, 20549u, 20741u, 86037u };
			AGB_MASK_TABLE uint8_t double_four_len[6] = { 7, 7, 7, 8, 8, 9 };
			AGB_MASK_TABLE uint8_t double_four_off[6] = { 2, 3, 4, 2, 3, 2 };
			AGB_MASK_TABLE uint32_t half_open_four[20] = { 21u, 69u, 81u, 84u, 21u, 261u, 273u, 276u, 69u, 261u, 321u, 324u, 81u, 273u, 321u, 336u, 84u, 276u, 324u, 336u };
			AGB_MASK_TABLE uint8_t half_open_four_off[20] = { 3, 4, 5, 6, 2, 4, 5, 6, 2, 3, 5, 6, 2, 3, 4, 6, 2, 3, 4, 5 };
			AGB_MASK_TABLE uint32_t open_three[12] = { 20u, 68u, 80u, 20u, 260u, 272u, 68u, 260u, 320u, 80u, 272u, 320u };
			AGB_MASK_TABLE uint8_t open_three_off[12] = { 3, 4, 5, 2, 4, 5, 2, 3, 5, 2, 3, 4 };
		}
		struct DefMasks
		{
				static AGB_HD uint32_t five(int i) { return mask_tables::five[i]; }
				static AGB_HD uint32_t open_four(int i) { return mask_tables::open_four[i]; }
				static AGB_HD uint32_t double_four(int i) { return mask_tables::double_four[i]; }
				static AGB_HD int double_four_len(int i) { return mask_tables::double_four_len[i]; }
				static AGB_HD int double_four_off(int i) { return mask_tables::double_four_off[i]; }
				static AGB_HD uint32_t half_open_four(int i) { return mask_tables::half_open_four[i]; }
				static AGB_HD int half_open_four_off(int i) { return mask_tables::half_open_four_off[i]; }
				static AGB_HD uint32_t open_three(int i) { return mask_tables::open_three[i]; }
				static AGB_HD int open_three_off(int i) { return mask_tables::open_three_off[i]; }
		};
		AGB_HD inline uint32_t sub_pattern(uint32_t line, int start, int length)
		{
			return (line >> (2u * start)) & ((1u << (2u * length)) - 1u);
		}
		AGB_HD inline uint32_t neighbour_index(uint32_t line, int left, int right)
		{
			return ((line >> (2 * (left - 2))) & 15u) | (((line >> (2 * right)) & 15u) << 4);
		}
		AGB_HD inline bool sides_allow_five(int rules, int attacker, int defender, int first, int last)
		{ // CheckSides / the side conditions of DefendFive::is_five (DefensiveMoveTable.cpp:58-66, 96-116)
			const bool allow_overline = (rules == RULE_FREESTYLE) or (rules == RULE_RENJU and attacker == CIRCLE) or (rules == RULE_CARO6);
			const bool allow_blocked = (rules != RULE_CARO5 and rules != RULE_CARO6);
			if (not allow_overline and (first == attacker or last == attacker))
				return false;
			if (not allow_blocked and (first == defender and last == defender))
				return false;
			return true;
		}
		// DefensiveMoveTable::getMoves: 13-cell window around an attacker threat -> bit i set: cell (i - 6) along the line defends
		AGB_HD_NOINLINE inline uint32_t defensive_mask(const uint16_t *table, int rules, uint32_t window13, int defender, int threat)
		{
			const int attacker = 3 - defender;
			const uint32_t colour_scale = (attacker == CROSS) ? 1u : 2u; // circle masks are the cross masks with every stone doubled
			const int d = defender - 1;
			switch (threat)
			{
				case PT_FIVE:
#pragma unroll 1
					for (int i = 0; i < 5; i++)
						if (sub_pattern(window13, 2 + i, 5) == DefMasks::five(i) * colour_scale)
							return table[((0 + i) * 256 + neighbour_index(window13, 2 + i, 2 + i + 5)) * 2 + d];
					return 0;
				case PT_OPEN_4:
#pragma unroll 1
					for (int i = 0; i < 4; i++)
						if (sub_pattern(window13, 2 + i, 6) == DefMasks::open_four(i) * colour_scale)
							return table[((5 + i) * 256 + neighbour_index(window13, 2 + i, 2 + i + 6)) * 2 + d];
					return 0;
				case PT_DOUBLE_4:
#pragma unroll 1
					for (int i = 0; i < 6; i++)
					{
						const int len = DefMasks::double_four_len(i), begin = DefMasks::double_four_off(i);
						if (sub_pattern(window13, begin, len) == DefMasks::double_four(i) * colour_scale)
							return table[((9 + i) * 256 + neighbour_index(window13, begin, begin + len)) * 2 + d];
					}
					return 0;
				case PT_HALF_OPEN_4:
				{
					uint32_t result = 1u << 6;
#pragma unroll 1
					for (int i = 0; i < 20; i++)
					{
						const int begin = DefMasks::half_open_four_off(i);
						if (sub_pattern(window13, begin, 5) == DefMasks::half_open_four(i) * colour_scale
								and sides_allow_five(rules, attacker, defender, (window13 >> (2 * (begin - 1))) & 3, (window13 >> (2 * (begin + 5))) & 3))
						{
							uint32_t tmp = table[((0 + i / 4) * 256 + neighbour_index(window13, begin, begin + 5)) * 2 + d];
							const int shift = begin - (2 + i / 4);
							tmp = (shift >= 0) ? (tmp << shift) : (tmp >> (-shift));
							result |= tmp & 0xFFFFu;
							if (rules != RULE_CARO5 and rules != RULE_CARO6)
								return result;
						}
					}
					return result;
				}
				case PT_OPEN_3:
#pragma unroll 1
					for (int i = 0; i < 12; i++)
					{
						const int begin = DefMasks::open_three_off(i);
						if (sub_pattern(window13, begin, 6) == DefMasks::open_three(i) * colour_scale)
						{
							uint32_t result = table[((5 + i / 3) * 256 + neighbour_index(window13, begin, begin + 6)) * 2 + d];
							const int shift = begin - (2 + i / 3);
							result = ((shift >= 0) ? (result << shift) : (result >> (-shift))) & 0xFFFFu;
							return result | (1u << 6);
						}
					}
					return 0;
				default:
					return 0;
			}
		}

		// ---- host: construction of the tables by exhaustive mini-search on short lines (DefendFive, DefensiveMoveTable.cpp:118-216) ---
		namespace build
		{
			inline int cell(uint32_t line, int i) { return (line >> (2 * i)) & 3; }
			inline uint32_t with(uint32_t line, int i, int v) { return (line & ~(3u << (2 * i))) | (static_cast<uint32_t>(v) << (2 * i)); }
			inline bool has_five(uint32_t line, int size, int rules, int attacker, int defender)
			{
#pragma unroll 1
				for (int i = 1; i < size - 5; i++)
				{
					bool all = true;
#pragma unroll 1
					for (int k = 0; k < 5; k++)
						all = all and cell(line, i + k) == attacker;
					if (all and sides_allow_five(rules, attacker, defender, cell(line, i - 1), cell(line, i + 5)))
						return true;
				}
				return false;
			}
			inline int search(uint32_t line, int size, int rules, int attacker, int defender, int sign, int depth)
			{ // +1: `sign` to move can force the attacker's five, 0: nothing decided within depth, -1: no empty cell
				int outcome = -1;
#pragma unroll 1
				for (int i = 0; i < size; i++)
					if (cell(line, i) == NONE)
					{
						const uint32_t next = with(line, i, sign);
						if (has_five(next, size, rules, attacker, defender))
							return 1;
						const int tmp = (depth > 1) ? -search(next, size, rules, attacker, defender, 3 - sign, depth - 1) : 0;
						outcome = outcome > tmp ? outcome : tmp;
					}
				return outcome;
			}
			inline uint16_t defend(uint32_t line, int size, int offset, int rules, int defender, int depth)
			{
				const int attacker = 3 - defender;
				if (has_five(line, size, rules, attacker, defender))
					return 0;
				if (search(line, size, rules, attacker, defender, attacker, depth) == 0)
					return 0;
				uint16_t result = 0;
#pragma unroll 1
				for (int i = 0; i < size; i++)
					if (cell(line, i) == NONE)
						if (search(with(line, i, defender), size, rules, attacker, defender, attacker, depth) != 1)
							result |= static_cast<uint16_t>(1u << (offset + i));
				return result;
			}
			// table[kDefGroups][256][2]
			inline void defensive_table(int rules, uint16_t *table)
			{
#pragma unroll 1
				for (int g = 0; g < kDefGroups; g++)
				{
					uint32_t base;
					int len, off, depth;
					if (g < 5) { base = DefMasks::five(g); len = 5; off = 2 + g; depth = 1; }
					else if (g < 9) { base = DefMasks::open_four(g - 5); len = 6; off = 2 + (g - 5); depth = 3; }
					else { base = DefMasks::double_four(g - 9); len = DefMasks::double_four_len(g - 9); off = DefMasks::double_four_off(g - 9); depth = 3; }
#pragma unroll 1
					for (int j = 0; j < 256; j++)
#pragma unroll 1
						for (int d = 0; d < 2; d++)
						{ // defender d+1 faces the pattern drawn with the attacker's stones
							const uint32_t stones = base * ((d == 0) ? 2u : 1u); // cross defends circle stones and vice versa
							const uint32_t line = (j & 15u) | (stones << 4) | (static_cast<uint32_t>(j & 0xF0) << (2 * len));
							table[(g * 256 + j) * 2 + d] = defend(line, len + 4, off - 2, rules, d + 1, depth);
						}
				}
			}
		}

		// ---- view of one analysed position ---------------------------------------------------------------------------------------
		// On the device the position under search -- line words, pattern types, threats, board, list lengths of ONE position -- sits in the shared
		// memory of the warp that searches it (solver_kernel.cuh stages it there). The members below that "point" to those arrays are then not
		// pointers but empty handles whose address is a compile-time offset from the warp's base: every access is a plain LDS / STS instead of a
		// generic load through a pointer fetched from a stack object (no descriptor set-up, no dependent pointer load). The host build (tests)
		// keeps ordinary pointers.
		namespace position_layout
		{
			// kListIndex: for every cell and colour, where the cell sits in the threat list it is on ([row * kMaxSize + col][colour], uint16): removing a
			// cell from a list is then "move the last entry into its place" without searching for it
			constexpr int kLines = 0, kPtypes = kLinePitch * 8, kThreats = kPtypes + kMaxCells * 4, kBoard = kThreats + kMaxCells, kHistCount = kBoard + kMaxCells,
					kListIndex = kHistCount + 96, kBytes = kListIndex + kMaxCells * 4; // 5120 B per warp
		}
#ifdef __CUDA_ARCH__
		__device__ __forceinline__ unsigned char* warp_position()
		{
			extern __shared__ __align__(16) unsigned char solver_smem[];
			return solver_smem + (threadIdx.x >> 5) * position_layout::kBytes;
		}
		template<typename T, int kOffset>
		struct PositionArray
		{
				PositionArray() = default;
				template<typename U>
				__device__ __forceinline__ PositionArray(const U&) {} // "assigned" from the staged copy's address: it IS that copy
				__device__ __forceinline__ T* get() const { return reinterpret_cast<T*>(warp_position() + kOffset); }
				__device__ __forceinline__ T& operator[](int i) const { return get()[i]; }
				__device__ __forceinline__ operator T*() const { return get(); }
				__device__ __forceinline__ T* operator+(int i) const { return get() + i; }
		};
#define AGB_POSITION_ARRAY(type, offset) PositionArray<type, position_layout::offset>
#else
#define AGB_POSITION_ARRAY(type, offset) type*
#endif
		struct DynState;
		AGB_HD_NOINLINE inline bool dyn_is_forbidden(DynState *d, int sign, int r, int c); // solver_search.cuh: live state, reference side effects
		enum : int { GEN_BASIC = 0, GEN_THREATS = 1, GEN_OPTIMAL = 2, GEN_REDUCED = 3, GEN_LEGAL = 4 }; // MoveGeneratorMode (MoveGenerator.hpp)
		struct View
		{
				int S, cells, rules, stm, stones, draw_after, pitch; // pitch: cells per list in hist_cells
				AGB_POSITION_ARRAY(const int8_t, kBoard) board;
				AGB_POSITION_ARRAY(const uint64_t, kLines) lines;
				AGB_POSITION_ARRAY(const uint32_t, kPtypes) ptypes;
				AGB_POSITION_ARRAY(const uint8_t, kThreats) threats;
				const uint8_t *forbidden; // renju: PatternCalculator::isForbidden(CROSS, .) per cell
				AGB_POSITION_ARRAY(const int32_t, kHistCount) hist_count; // [2][10]
				const uint16_t *hist_cells; // [2][10][pitch]
				const uint8_t *pattern_table;
				const uint16_t *def_table;
				DynState *dyn = nullptr; // set when the position is searched in place (forbidden moves are then evaluated on the live state)

				AGB_HD int own() const { return stm; }
				AGB_HD int opp() const { return 3 - stm; }
				AGB_HD int count(int sign, int tt) const { return hist_count[(sign - 1) * kHistTypes + tt]; }
				AGB_HD uint16_t item(int sign, int tt, int i) const { return hist_cells[((sign - 1) * kHistTypes + tt) * pitch + i]; }
				AGB_HD int threat_at(int sign, int r, int c) const { return (threats[r * S + c] >> (4 * (sign - 1))) & 15; }
				AGB_HD int ptype_at(int sign, int r, int c, int dir) const { return (ptypes[r * S + c] >> (8 * dir + 4 * (sign - 1))) & 7; }
				AGB_HD bool half_open_three_at(int sign, int r, int c, int dir) const
				{
					return (pattern_table[narrow_window(normal_window(lines, dir, r, c, S))] >> (4 * (sign - 1) + 3)) & 1;
				}
				AGB_HD bool group_contains(int sign, int r, int c, int pt) const
				{
#pragma unroll 1
					for (int d = 0; d < 4; d++)
						if (ptype_at(sign, r, c, d) == pt)
							return true;
					return false;
				}
				AGB_HD int group_count(int sign, int r, int c, int pt) const
				{
					int n = 0;
#pragma unroll 1
					for (int d = 0; d < 4; d++)
						n += (ptype_at(sign, r, c, d) == pt);
					return n;
				}
				AGB_HD int direction_of(int sign, int r, int c, int pt) const
				{
#pragma unroll 1
					for (int d = 0; d < 4; d++)
						if (ptype_at(sign, r, c, d) == pt)
							return d;
					return 0;
				}
#ifdef AGB_SOLVER_NO_RENJU
				AGB_HD bool anything_forbidden_for(int) const { return false; } // this build of the solver serves the rule sets without forbidden moves
#else
				AGB_HD bool anything_forbidden_for(int sign) const { return rules == RULE_RENJU and sign == CROSS; }
#endif
				AGB_HD bool is_forbidden(int sign, int r, int c) const
				{
					if (dyn != nullptr)
						return dyn_is_forbidden(dyn, sign, r, c);
					return anything_forbidden_for(sign) and forbidden[r * S + c] != 0;
				}
				AGB_HD bool has_any_four(int sign) const
				{
					return count(sign, TT_HALF_OPEN_4) > 0 or count(sign, TT_FORK_4x3) > 0 or count(sign, TT_FORK_4x4) > 0 or count(sign, TT_OPEN_4) > 0;
				}
				AGB_HD int available_fours(int sign) const
				{
					return count(sign, TT_OPEN_4) + (anything_forbidden_for(sign) ? 0 : count(sign, TT_FORK_4x4)) + count(sign, TT_FORK_4x3)
							+ count(sign, TT_HALF_OPEN_4);
				}
		};
		AGB_HD inline int loc_row(uint16_t l) { return l & 255; }
		AGB_HD inline int loc_col(uint16_t l) { return l >> 8; }
		AGB_HD inline uint16_t mk_loc(int r, int c) { return static_cast<uint16_t>((c << 8) | r); }

		template<int N>
		struct LocList
		{ // StackVector<Location, N> (patterns/common.hpp:149-244)
				uint16_t data[N];
				int size = 0;
				AGB_HD bool contains(uint16_t v) const
				{
#pragma unroll 1
					for (int i = 0; i < size; i++)
						if (data[i] == v)
							return true;
					return false;
				}
				AGB_HD void add(uint16_t v) { if (size < N) data[size++] = v; }
				AGB_HD void remove_at(int i) { data[i] = data[--size]; }
				AGB_HD void remove_value(uint16_t v)
				{
#pragma unroll 1
					for (int i = 0; i < size; i++)
						if (data[i] == v)
						{
							data[i] = data[--size];
							return;
						}
				}
		};
		struct DefensiveSet
		{ // DefensiveMoves (MoveGenerator.cpp:93-121): running intersection
				LocList<24> list;
				bool not_initialized = true;
				template<int N>
				AGB_HD_NOINLINE void intersect(const LocList<N> &other)
				{
					if (not_initialized)
					{
#pragma unroll 1
						for (int i = 0; i < other.size; i++)
							list.add(other.data[i]);
						not_initialized = false;
					}
					else
					{
						int i = 0;
						while (i < list.size)
						{
							if (other.contains(list.data[i]))
								i++;
							else
								list.remove_at(i);
						}
					}
				}
				AGB_HD bool empty() const { return list.size == 0; }
		};

		struct Result
		{
				uint16_t score = kScoreDefault; // static score returned by generate()
				int n_actions = 0;
				bool must_defend = false, has_initiative = false, is_fully_expanded = false;
				uint16_t baseline = kScoreDefault;
		};

		// The generator. `moves` / `scores` receive the ordered action list (capacity: cells).
		struct MoveGenerator
		{
				const View &v;
				uint16_t *moves;
				uint16_t *scores;
				Result out;
				uint32_t added[kMaxSize]; // bitmask of cells already in the list (MoveGenerator::moves)
				// MoveGenerator::temporary_list: a copy of a threat list taken before a loop that may call is_forbidden(), which re-plays
				// moves on the calculator and so may reorder the original list (MoveGenerator.cpp:1181-1185)
				static constexpr int kTempCapacity = 96;
				uint16_t temp[kTempCapacity];
				int temp_size = 0;
				bool temp_overflow = false;
				AGB_HD_NOINLINE void copy_list(int sign, int tt)
				{
					temp_size = v.count(sign, tt);
					if (temp_size > kTempCapacity)
					{
						temp_size = kTempCapacity;
						temp_overflow = true;
					}
#pragma unroll 1
					for (int i = 0; i < temp_size; i++)
						temp[i] = v.item(sign, tt, i);
				}

				AGB_HD MoveGenerator(const View &view, uint16_t *m, uint16_t *s) : v(view), moves(m), scores(s)
				{
#pragma unroll
					for (int i = 0; i < kMaxSize; i++) // 20 words, unrolled into a few wide stores
						added[i] = 0;
				}
				AGB_HD uint16_t wire(uint16_t loc) const { return static_cast<uint16_t>(v.own() | (loc_row(loc) << 2) | (loc_col(loc) << 9)); }
				AGB_HD_NOINLINE void add_move(uint16_t loc, uint16_t s, bool override_duplicate)
				{
					const int r = loc_row(loc), c = loc_col(loc);
					if ((added[r] >> c) & 1u)
					{
						if (override_duplicate)
						{
							const uint16_t w = wire(loc);
#pragma unroll 1
							for (int i = 0; i < out.n_actions; i++)
								if (moves[i] == w)
								{
									scores[i] = s;
									return;
								}
						}
					}
					else
					{
						moves[out.n_actions] = wire(loc);
						scores[out.n_actions] = s;
						out.n_actions++;
						added[r] |= 1u << c;
					}
				}
				AGB_HD_NOINLINE void add_list(int sign, int tt, uint16_t s, bool override_duplicate = false)
				{
					const int n = v.count(sign, tt);
#pragma unroll 1
					for (int i = 0; i < n; i++)
						add_move(v.item(sign, tt, i), s, override_duplicate);
				}
				template<int N>
				AGB_HD_NOINLINE void add_all(const LocList<N> &l, uint16_t s = kScoreDefault)
				{
#pragma unroll 1
					for (int i = 0; i < l.size; i++)
						add_move(l.data[i], s, false);
				}
				// PatternCalculator::getDefensiveMoves: all cells of the table mask, in increasing offset order
				AGB_HD_NOINLINE LocList<7> raw_defensive_moves(int defender, int r, int c, int dir) const
				{
					const uint32_t window = extended_window(v.lines, dir, r, c, v.S);
					const int threat = v.ptype_at(3 - defender, r, c, dir);
					const uint32_t mask = defensive_mask(v.def_table, v.rules, window, defender, threat);
					LocList<7> result;
#pragma unroll 1
					for (int i = -6; i <= 6; i++)
						if ((mask >> (6 + i)) & 1u)
							result.add(mk_loc(r + i * dir_row_step(dir), c + i * dir_col_step(dir)));
					return result;
				}
				// MoveGenerator::get_defensive_moves (MoveGenerator.cpp:259-307)
				AGB_HD_NOINLINE LocList<7> get_defensive_moves(uint16_t loc, int dir)
				{
					const int r = loc_row(loc), c = loc_col(loc);
					LocList<7> result = raw_defensive_moves(v.own(), r, c, dir);
					if (v.anything_forbidden_for(v.own()))
					{
						int i = 0;
						while (i < result.size)
						{
							if (v.is_forbidden(v.own(), loc_row(result.data[i]), loc_col(result.data[i])))
							{
								add_move(result.data[i], loss_in(1), true);
								result.remove_at(i);
							}
							else
								i++;
						}
					}
					else if (v.anything_forbidden_for(v.opp()))
					{
						if (v.ptype_at(v.opp(), r, c, dir) == PT_OPEN_4)
						{
							const uint32_t raw = extended_window(v.lines, dir, r, c, v.S);
							int type = 0;
							if ((raw & 65520u) == 1344u)
								type = -1;
							if ((raw & 4193280u) == 344064u)
								type = +1;
							if (type != 0)
							{
								const int fr = r + 4 * type * dir_row_step(dir), fc = c + 4 * type * dir_col_step(dir);
								if (fr >= 0 and fr < v.S and fc >= 0 and fc < v.S and v.is_forbidden(v.opp(), fr, fc))
									result.add(mk_loc(r - type * dir_row_step(dir), c - type * dir_col_step(dir)));
							}
						}
					}
					return result;
				}
				AGB_HD_NOINLINE void create_remaining(const uint32_t *mask, uint16_t s = kScoreDefault)
				{ // MoveGenerator::create_remaining_moves
					for (int r = 0; r < v.S; r++)
					{
						uint32_t tmp = mask[r] & ~added[r];
						for (int c = 0; c < v.S; c++, tmp >>= 1)
							if (tmp & 1u)
							{
								moves[out.n_actions] = static_cast<uint16_t>(v.own() | (r << 2) | (c << 9));
								scores[out.n_actions] = s;
								out.n_actions++;
							}
						added[r] |= mask[r];
					}
				}
				AGB_HD_NOINLINE void legal_mask(uint32_t *mask) const
				{ // cold: used one move before the draw, and by the host form of the stencil
					for (int r = 0; r < v.S; r++)
					{
						uint32_t m = 0;
						for (int c = 0; c < v.S; c++)
							m |= static_cast<uint32_t>(v.board[r * v.S + c] == NONE) << c;
						mask[r] = m;
					}
				}
				// 7x7 stencils around stones (MoveGenerator.cpp:1011-1125); `sign` 0: around every stone (box-and-star), else star around `sign`
				AGB_HD_NOINLINE void stencil_mask(uint32_t *mask, int sign) const
				{
#ifdef __CUDA_ARCH__
					// lockstep warp (solver_search.cuh): lane r packs row r into bit masks (stones of interest, empty cells), takes the stone masks of
					// rows r-3..r+3 from its neighbours, and ORs each of them in shifted by every offset its stencil row has; then every lane
					// collects all rows
					const unsigned long long box = 0x493E3E773E3E49ull, star = 0x492A1C771C2A49ull; // the 7 stencil rows, one byte each (bit k: offset k - 3)
					const unsigned long long stencil = (sign == 0) ? box : star;
					const int lane = threadIdx.x & 31;
					uint32_t stones = 0, legal = 0;
					if (lane < v.S)
#pragma unroll 1
						for (int c = 0; c < v.S; c++)
						{
							const int s = v.board[lane * v.S + c];
							stones |= static_cast<uint32_t>((sign == 0) ? (s != NONE) : (s == sign)) << c;
							legal |= static_cast<uint32_t>(s == NONE) << c;
						}
					uint32_t mine = 0;
#pragma unroll 1
					for (int i = 0; i < 7; i++)
					{ // a stone in row lane + 3 - i puts stencil row i on this row
						const int sr = lane + 3 - i;
						const uint32_t theirs = __shfl_sync(0xFFFFFFFFu, stones, sr & 31);
						if (sr < 0 or sr >= v.S)
							continue;
						const uint32_t pattern = static_cast<uint32_t>(stencil >> (8 * i)) & 0x7Fu;
						uint64_t spread = 0; // bit c + k of it: a stone at column c seen through pattern bit k, i.e. column c + k - 3
#pragma unroll
						for (int k = 0; k < 7; k++)
							spread |= ((pattern >> k) & 1u) ? (static_cast<uint64_t>(theirs) << k) : 0ull;
						mine |= static_cast<uint32_t>((spread << 3) >> 6); // columns below 0 fall off, those beyond the board are masked below
					}
					if (sign == 0 and v.stones == 0 and lane == v.S / 2)
						mine |= 1u << (v.S / 2);
					mine &= legal & ((1u << v.S) - 1u);
					if (lane >= v.S)
						mine = 0;
					for (int r = 0; r < v.S; r++)
						mask[r] = __shfl_sync(0xFFFFFFFFu, mine, r);
#else
					const uint32_t box[7] = { 73u, 62u, 62u, 119u, 62u, 62u, 73u };
					const uint32_t star[7] = { 73u, 42u, 28u, 119u, 28u, 42u, 73u };
					uint32_t rows[kMaxSize + 7];
					for (int i = 0; i < kMaxSize + 7; i++)
						rows[i] = 0;
					for (int r = 0; r < v.S; r++)
						for (int c = 0; c < v.S; c++)
						{
							const int s = v.board[r * v.S + c];
							if ((sign == 0) ? (s != NONE) : (s == sign))
								for (int i = 0; i < 7; i++)
								{
									const uint32_t m = ((sign == 0) ? box[i] : star[i]) << 25;
									rows[r + i] |= m >> (28 - c);
								}
						}
					if (sign == 0 and v.stones == 0)
						rows[3 + v.S / 2] |= 1u << (v.S / 2);
					uint32_t legal[kMaxSize];
					legal_mask(legal);
					for (int r = 0; r < v.S; r++)
						mask[r] = rows[3 + r] & legal[r];
#endif
				}

				// ---- stages -----------------------------------------------------------------------------------------------------
				AGB_HD_NOINLINE uint16_t try_solve_own_fork_4x3(uint16_t loc)
				{ // MoveGenerator.cpp:952-994
					const uint16_t prior = sc_eval(15);
					if (v.anything_forbidden_for(v.own()))
						return prior;
					const int r = loc_row(loc), c = loc_col(loc);
					const int dir = v.direction_of(v.own(), r, c, PT_HALF_OPEN_4);
					LocList<7> def = raw_defensive_moves(v.opp(), r, c, dir);
					def.remove_value(loc);
					int best = TT_NONE;
#pragma unroll 1
					for (int i = 0; i < def.size; i++)
					{
						const int tt = v.threat_at(v.opp(), loc_row(def.data[i]), loc_col(def.data[i]));
						if ((tt != TT_FORK_4x4 and tt != TT_OVERLINE) or not v.anything_forbidden_for(v.opp()))
							best = best > tt ? best : tt;
					}
					switch (best)
					{
						default:
						case TT_NONE:
						case TT_HALF_OPEN_3:
						case TT_OPEN_3:
						case TT_FORK_3x3:
							return win_in(5);
						case TT_HALF_OPEN_4:
						case TT_FORK_4x3:
							return prior;
						case TT_FORK_4x4:
						case TT_OPEN_4:
							return loss_in(4);
						case TT_FIVE:
						case TT_OVERLINE:
							return loss_in(2);
					}
				}
				AGB_HD_NOINLINE uint16_t add_own_4x3_forks()
				{
					uint16_t result = kScoreDefault;
					const int n = v.count(v.own(), TT_FORK_4x3);
#pragma unroll 1
					for (int i = 0; i < n; i++)
					{
						const uint16_t loc = v.item(v.own(), TT_FORK_4x3, i);
						const uint16_t solution = try_solve_own_fork_4x3(loc);
						add_move(loc, solution, true);
						if (sc_is_proven(solution))
							result = result > solution ? result : solution;
					}
					return result;
				}
				AGB_HD_NOINLINE void add_own_half_open_fours()
				{
					const uint16_t prior = sc_eval(14);
					int hidden = 0;
					if (v.anything_forbidden_for(v.own()))
					{
						copy_list(v.own(), TT_FORK_3x3);
#pragma unroll 1
						for (int i = 0; i < temp_size; i++)
						{
							const uint16_t loc = temp[i];
							if (v.group_contains(v.own(), loc_row(loc), loc_col(loc), PT_HALF_OPEN_4) and not v.is_forbidden(v.own(), loc_row(loc), loc_col(loc)))
							{
								add_move(loc, prior, false);
								hidden++;
							}
						}
					}
					add_list(v.own(), TT_HALF_OPEN_4, prior);
					if (hidden + v.count(v.own(), TT_HALF_OPEN_4) > 0)
						out.has_initiative = true;
				}
				// each stage returns true when generation can stop; `score` then holds the stage's score
				AGB_HD bool try_win_in_1(uint16_t &score)
				{
					if (v.count(v.own(), TT_FIVE) > 0)
					{
						out.has_initiative = true;
						add_list(v.own(), TT_FIVE, win_in(1));
						score = win_in(1);
						return true;
					}
					return false;
				}
				AGB_HD_NOINLINE bool try_draw_in_1(uint16_t &score)
				{ // MoveGenerator.cpp:308-355
					out.baseline = draw_in(1);
					if (v.anything_forbidden_for(v.own()))
					{
						bool found = false;
						for (int r = 0; r < v.S; r++)
							for (int c = 0; c < v.S; c++)
								if (v.board[r * v.S + c] == NONE)
								{
									const int threat = v.threat_at(v.own(), r, c);
									if (threat == TT_FORK_4x4 or threat == TT_OVERLINE or (threat == TT_FORK_3x3 and v.is_forbidden(v.own(), r, c)))
										add_move(mk_loc(r, c), loss_in(1), false);
									else
									{
										add_move(mk_loc(r, c), draw_in(1), false);
										found = true;
									}
								}
						score = found ? draw_in(1) : loss_in(1);
					}
					else
					{
						uint32_t legal[kMaxSize];
						legal_mask(legal);
						create_remaining(legal, draw_in(1));
						score = draw_in(1);
					}
					return true;
				}
				AGB_HD_NOINLINE bool defend_loss_in_2(uint16_t &score)
				{ // MoveGenerator.cpp:372-461
					const int n_fives = v.count(v.opp(), TT_FIVE);
					if (n_fives == 0)
						return false;
					out.must_defend = true;
					out.baseline = loss_in(2);
					DefensiveSet defensive;
#pragma unroll 1
					for (int i = 0; i < n_fives; i++)
					{
						const uint16_t loc = v.item(v.opp(), TT_FIVE, i);
						const int dir = v.direction_of(v.opp(), loc_row(loc), loc_col(loc), PT_FIVE);
						defensive.intersect(get_defensive_moves(loc, dir));
						if (defensive.empty())
						{
							add_list(v.opp(), TT_FIVE, loss_in(2));
							score = loss_in(2);
							return true;
						}
					}
					uint16_t best = kScoreMin;
#pragma unroll 1
					for (int i = 0; i < defensive.list.size; i++)
					{
						const uint16_t loc = defensive.list.data[i];
						const int r = loc_row(loc), c = loc_col(loc);
						uint16_t response = kScoreDefault;
						switch (v.threat_at(v.own(), r, c))
						{
							case TT_FORK_3x3:
								if (v.anything_forbidden_for(v.own()))
								{
									if (v.group_contains(v.own(), r, c, PT_OPEN_4))
										response = win_in(3);
								}
								else if (not v.has_any_four(v.opp()))
									response = win_in(5);
								break;
							case TT_FORK_4x3:
							{
								const uint16_t solution = try_solve_own_fork_4x3(loc);
								response = sc_is_proven(solution) ? solution : sc_eval(15);
								break;
							}
							case TT_FORK_4x4:
							case TT_OPEN_4:
								response = win_in(3);
								break;
							default:
								if (v.group_contains(v.own(), r, c, PT_HALF_OPEN_4))
								{
									out.has_initiative = true;
									response = sc_eval(14);
								}
								break;
						}
						if (sc_is_win(response))
							out.has_initiative = true;
						add_move(loc, response, false);
						best = best > response ? best : response;
					}
					score = best;
					return true;
				}
				AGB_HD_NOINLINE bool try_win_in_3(uint16_t &score)
				{ // MoveGenerator.cpp:462-553
					int threats = 0;
					if (v.anything_forbidden_for(v.own()))
					{
						copy_list(v.own(), TT_FORK_3x3);
#pragma unroll 1
						for (int i = 0; i < temp_size; i++)
						{
							const uint16_t loc = temp[i];
							if (v.group_contains(v.own(), loc_row(loc), loc_col(loc), PT_OPEN_4) and not v.is_forbidden(v.own(), loc_row(loc), loc_col(loc)))
							{
								threats++;
								add_move(loc, win_in(3), false);
							}
						}
					}
					add_list(v.own(), TT_OPEN_4, win_in(3));
					threats += v.count(v.own(), TT_OPEN_4);
					if (v.count(v.own(), TT_FORK_4x4) > 0 and not v.anything_forbidden_for(v.own()))
					{
						threats += v.count(v.own(), TT_FORK_4x4);
						add_list(v.own(), TT_FORK_4x4, win_in(3));
					}
					if (v.anything_forbidden_for(v.opp()))
					{ // renju, white to move: a four whose only answer is a forbidden point for black
						copy_list(v.own(), TT_HALF_OPEN_4);
#pragma unroll 1
						for (int i = 0; i < temp_size; i++)
						{
							const uint16_t loc = temp[i];
							const int r = loc_row(loc), c = loc_col(loc);
							const int dir = v.direction_of(v.own(), r, c, PT_HALF_OPEN_4);
							bool winning = false;
							switch (v.threat_at(v.opp(), r, c))
							{
								default:
									break;
								case TT_FORK_3x3:
									if (v.ptype_at(v.opp(), r, c, dir) != PT_OPEN_3 and v.is_forbidden(v.opp(), r, c))
										winning = true;
									break;
								case TT_FORK_4x4:
								case TT_OVERLINE:
									winning = true;
									break;
							}
							if (winning)
							{
								const LocList<7> tmp = raw_defensive_moves(v.opp(), r, c, dir);
								const uint16_t original = (tmp.data[0] == loc) ? tmp.data[1] : tmp.data[0];
								add_move(original, win_in(3), false);
								score = win_in(3);
								return true;
							}
						}
					}
					if (threats > 0)
					{
						out.has_initiative = true;
						score = win_in(3);
						return true;
					}
					return false;
				}
				AGB_HD_NOINLINE bool defend_loss_in_4(uint16_t &score)
				{ // MoveGenerator.cpp:554-685
					const bool has_any_four = v.has_any_four(v.own());
					out.baseline = loss_in(4);
					if (v.rules != RULE_RENJU)
					{
						DefensiveSet defensive;
						const int n_open4 = v.count(v.opp(), TT_OPEN_4);
#pragma unroll 1
						for (int i = 0; i < n_open4; i++)
						{
							out.must_defend = true;
							const uint16_t loc = v.item(v.opp(), TT_OPEN_4, i);
							const int dir = v.direction_of(v.opp(), loc_row(loc), loc_col(loc), PT_OPEN_4);
							defensive.intersect(get_defensive_moves(loc, dir));
							if (defensive.empty() and not has_any_four)
							{
								add_list(v.opp(), TT_OPEN_4, loss_in(4));
								score = loss_in(4);
								return true;
							}
						}
						const int n_44 = v.count(v.opp(), TT_FORK_4x4);
#pragma unroll 1
						for (int i = 0; i < n_44; i++)
						{
							out.must_defend = true;
							const uint16_t loc = v.item(v.opp(), TT_FORK_4x4, i);
							const int r = loc_row(loc), c = loc_col(loc);
#pragma unroll 1
							for (int dir = 0; dir < 4; dir++)
							{
								const int pt = v.ptype_at(v.opp(), r, c, dir);
								if (pt == PT_OPEN_4 or pt == PT_DOUBLE_4)
									defensive.intersect(get_defensive_moves(loc, dir));
							}
							if (v.group_count(v.opp(), r, c, PT_HALF_OPEN_4) > 0)
							{
								LocList<24> storage;
#pragma unroll 1
								for (int dir = 0; dir < 4; dir++)
									if (v.ptype_at(v.opp(), r, c, dir) == PT_HALF_OPEN_4)
									{
										const LocList<7> tmp = get_defensive_moves(loc, dir);
#pragma unroll 1
										for (int k = 0; k < tmp.size; k++)
											if (not storage.contains(tmp.data[k]))
												storage.add(tmp.data[k]);
									}
								defensive.intersect(storage);
							}
							if (defensive.empty() and not has_any_four)
							{
								add_list(v.opp(), TT_FORK_4x4, loss_in(4));
								score = loss_in(4);
								return true;
							}
						}
						add_all(defensive.list);
					}
					else
					{
						copy_list(v.opp(), TT_OPEN_4);
#pragma unroll 1
						for (int i = 0; i < temp_size; i++)
						{
							out.must_defend = true;
							const uint16_t loc = temp[i];
							const int dir = v.direction_of(v.opp(), loc_row(loc), loc_col(loc), PT_OPEN_4);
							add_all(get_defensive_moves(loc, dir));
						}
						if (v.anything_forbidden_for(v.opp()))
						{
							copy_list(v.opp(), TT_FORK_3x3);
#pragma unroll 1
							for (int i = 0; i < temp_size; i++)
							{
								const uint16_t loc = temp[i];
								const int r = loc_row(loc), c = loc_col(loc);
								if (v.group_contains(v.opp(), r, c, PT_OPEN_4) and not v.is_forbidden(v.opp(), r, c))
								{
									out.must_defend = true;
									add_all(get_defensive_moves(loc, v.direction_of(v.opp(), r, c, PT_OPEN_4)));
								}
							}
						}
						if (not v.anything_forbidden_for(v.opp()))
						{
							copy_list(v.opp(), TT_FORK_4x4);
#pragma unroll 1
							for (int i = 0; i < temp_size; i++)
							{
								out.must_defend = true;
								const uint16_t loc = temp[i];
#pragma unroll 1
								for (int dir = 0; dir < 4; dir++)
								{
									const int pt = v.ptype_at(v.opp(), loc_row(loc), loc_col(loc), dir);
									if (pt == PT_HALF_OPEN_4 or pt == PT_OPEN_4 or pt == PT_DOUBLE_4)
										add_all(get_defensive_moves(loc, dir));
								}
							}
						}
					}
					if (out.must_defend)
					{
						out.has_initiative = has_any_four;
						const uint16_t best = add_own_4x3_forks();
						add_own_half_open_fours();
						score = sc_is_win(best) ? best : kScoreDefault;
						return true;
					}
					out.baseline = kScoreDefault;
					return false;
				}
				AGB_HD_NOINLINE bool try_win_in_5(uint16_t &score)
				{ // MoveGenerator.cpp:686-715
					uint16_t best = add_own_4x3_forks();
					if (not v.anything_forbidden_for(v.own()) and v.available_fours(v.opp()) == 0 and v.count(v.own(), TT_FORK_3x3) > 0)
					{
						add_list(v.own(), TT_FORK_3x3, win_in(5));
						best = best > win_in(5) ? best : win_in(5);
					}
					if (sc_is_win(best))
					{
						out.has_initiative = true;
						score = best;
						return true;
					}
					return false;
				}
				AGB_HD_NOINLINE bool defend_loss_in_6(uint16_t &score)
				{ // MoveGenerator.cpp:716-815
					if (v.available_fours(v.own()) > 0)
						return false;
					const int n43 = v.count(v.opp(), TT_FORK_4x3), n33 = v.count(v.opp(), TT_FORK_3x3);
					if (n43 > 0 or n33 > 0)
					{
						out.must_defend = true;
						out.baseline = loss_in(6);
					}
					// the reference walks the LIVE lists here and dereferences its iterator again for every get_defensive_moves() call, while
					// the pattern group was read once at the top of the iteration; is_forbidden() inside may have reordered the list in between
#pragma unroll 1
					for (int i = 0; i < n43; i++)
					{
						const uint16_t loc = v.item(v.opp(), TT_FORK_4x3, i);
						const int r = loc_row(loc), c = loc_col(loc);
						int group[4];
#pragma unroll 1
						for (int dir = 0; dir < 4; dir++)
							group[dir] = v.ptype_at(v.opp(), r, c, dir);
#pragma unroll 1
						for (int dir = 0; dir < 4; dir++)
							if (group[dir] == PT_OPEN_3)
								add_all(get_defensive_moves(v.item(v.opp(), TT_FORK_4x3, i), dir), sc_eval(0));
						int four_dir = 0;
#pragma unroll 1
						for (int dir = 3; dir >= 0; dir--)
							if (group[dir] == PT_HALF_OPEN_4)
								four_dir = dir;
						const LocList<7> four_defence = get_defensive_moves(v.item(v.opp(), TT_FORK_4x3, i), four_dir);
						add_all(four_defence, sc_eval(0));
#pragma unroll 1
						for (int k = 0; k < four_defence.size; k++)
						{
							const int dr = loc_row(four_defence.data[k]), dc = loc_col(four_defence.data[k]);
#pragma unroll 1
							for (int dir = 0; dir < 4; dir++)
							{
								const uint32_t reduced = static_cast<uint32_t>(v.lines[line_index(dir, dr, dc, v.S)] >> (2 * pos_in_line(dir, dr, dc, v.S) + 4)) & 0x3FFFFu;
#pragma unroll 1
								for (int i2 = -4; i2 <= 4; i2++)
									if (((reduced >> (2 * (i2 + 4))) & 3u) == 0u)
									{
										const int tr = dr + i2 * dir_row_step(dir), tc = dc + i2 * dir_col_step(dir);
										if (v.ptype_at(v.own(), tr, tc, dir) > PT_NONE or v.half_open_three_at(v.own(), tr, tc, dir))
											add_move(mk_loc(tr, tc), kScoreDefault, false);
									}
							}
						}
					}
#pragma unroll 1
					for (int i = 0; i < n33; i++)
					{
						const uint16_t loc = v.item(v.opp(), TT_FORK_3x3, i);
						const int r = loc_row(loc), c = loc_col(loc);
						int group[4];
#pragma unroll 1
						for (int dir = 0; dir < 4; dir++)
							group[dir] = v.ptype_at(v.opp(), r, c, dir);
#pragma unroll 1
						for (int dir = 0; dir < 4; dir++)
							if (group[dir] == PT_OPEN_3)
								add_all(get_defensive_moves(v.item(v.opp(), TT_FORK_3x3, i), dir), sc_eval(0));
						add_list(v.own(), TT_FORK_3x3, sc_eval(13));
						add_list(v.own(), TT_OPEN_3, sc_eval(1));
						uint32_t mask[kMaxSize];
						stencil_mask(mask, v.own());
#pragma unroll 1
						for (int rr = 0; rr < v.S; rr++)
						{
							uint32_t tmp = mask[rr] & ~added[rr];
#pragma unroll 1
							for (int cc = 0; cc < v.S; cc++, tmp >>= 1)
								if (tmp & 1u)
#pragma unroll 1
									for (int dir = 0; dir < 4; dir++)
										if (v.half_open_three_at(v.own(), rr, cc, dir))
										{
											add_move(mk_loc(rr, cc), sc_eval(1), false);
											break;
										}
						}
					}
					if (out.must_defend)
					{
						add_own_half_open_fours();
						score = kScoreDefault;
						return true;
					}
					return false;
				}
				AGB_HD_NOINLINE void mark_forbidden_moves()
				{ // MoveGenerator.cpp:995-1010
					add_list(v.own(), TT_OVERLINE, loss_in(1), true);
					add_list(v.own(), TT_FORK_4x4, loss_in(1), true);
					copy_list(v.own(), TT_FORK_3x3);
#pragma unroll 1
					for (int i = 0; i < temp_size; i++)
					{
						const uint16_t loc = temp[i];
						if (v.is_forbidden(CROSS, loc_row(loc), loc_col(loc)))
							add_move(loc, loss_in(1), true);
					}
				}
				AGB_HD void generate_optimal() { generate(GEN_OPTIMAL); }
				// true when neither side has a 3x3 fork or anything stronger: the THREATS stages (MoveGenerator.cpp:176-196) and
				// mark_forbidden_moves then find every list they walk empty
				AGB_HD bool is_quiet() const
				{
#ifdef __CUDA_ARCH__
					const int lane = threadIdx.x & 31; // the warp runs in lockstep (solver_search.cuh): one list length per lane
					const bool busy = lane < 2 * kHistTypes and (lane % kHistTypes) >= TT_FORK_3x3 and v.hist_count[lane] != 0;
					return __ballot_sync(0xFFFFFFFFu, busy) == 0u;
#else
					for (int colour = 0; colour < 2; colour++)
						for (int t = TT_FORK_3x3; t < kHistTypes; t++)
							if (v.hist_count[colour * kHistTypes + t] != 0)
								return false;
					return true;
#endif
				}
				// MoveGenerator::generate in THREATS or OPTIMAL mode (MoveGenerator.cpp:159-223)
				AGB_HD_NOINLINE void generate(int mode)
				{
					const int distance_to_draw = v.draw_after - v.stones;
					if (distance_to_draw <= 0)
					{
						out.score = mk_score(PV_DRAW, 0);
						return;
					}
					if (mode == GEN_THREATS and distance_to_draw >= 2 and is_quiet())
					{ // no list any stage below looks at has an entry: they would all fall through without adding a move
						out.score = kScoreDefault;
						out.is_fully_expanded = false;
						return;
					}
					uint16_t score = kScoreDefault;
					bool stop = try_win_in_1(score);
					if (not stop and distance_to_draw == 1)
						stop = try_draw_in_1(score);
					if (not stop and distance_to_draw >= 2)
						stop = defend_loss_in_2(score);
					if (not stop and distance_to_draw >= 3)
						stop = try_win_in_3(score);
					if (not stop and distance_to_draw >= 4)
						stop = defend_loss_in_4(score);
					if (not stop and distance_to_draw >= 5)
						stop = try_win_in_5(score);
					if (not stop and distance_to_draw >= 6)
						stop = defend_loss_in_6(score);
					if (not stop and distance_to_draw >= 3)
						add_own_half_open_fours();
					if (not stop and mode >= GEN_OPTIMAL)
					{
						if (distance_to_draw >= 6)
						{
							add_list(v.opp(), TT_FORK_3x3, sc_eval(3));
							add_list(v.opp(), TT_OPEN_3, sc_eval(2));
						}
						if (distance_to_draw >= 5)
						{
							add_list(v.own(), TT_FORK_3x3, sc_eval(13));
							add_list(v.own(), TT_OPEN_3, sc_eval(1));
						}
						if (distance_to_draw >= 3)
							add_list(v.opp(), TT_HALF_OPEN_4, sc_eval(4));
						uint32_t mask[kMaxSize];
						stencil_mask(mask, 0);
						create_remaining(mask);
					}
					if (v.anything_forbidden_for(v.own()))
						mark_forbidden_moves();
					out.is_fully_expanded = out.must_defend or mode >= GEN_OPTIMAL;
					out.score = stop ? score : kScoreDefault;
				}
		};

		// AlphaBetaSearch::evaluate (AlphaBetaSearch.cpp:345-365)
		AGB_HD_NOINLINE inline uint16_t static_evaluation(const View &v)
		{
			int result = 12;
#ifdef __CUDA_ARCH__
			{ // one list length per lane, then a warp sum (the warp runs in lockstep, solver_search.cuh)
				const int lane = threadIdx.x & 31;
				int term = 0;
				if (lane < 2 * kHistTypes)
				{
					const int t = lane % kHistTypes;
					const bool own = (lane / kHistTypes) == v.own() - 1;
					int w = 0;
					switch (t)
					{
						case TT_OPEN_3: w = own ? 19 : -1; break;
						case TT_FORK_3x3: w = own ? 49 : -50; break;
						case TT_HALF_OPEN_4: w = own ? 76 : -45; break;
						case TT_FORK_4x3: w = own ? 170 : -135; break;
						case TT_FORK_4x4: w = own ? 33 : -14; break;
						case TT_OPEN_4: w = own ? 159 : -154; break;
						case TT_FIVE: w = own ? 252 : -496; break;
						default: break;
					}
					term = w * v.hist_count[lane];
				}
				for (int o = 16; o > 0; o >>= 1)
					term += __shfl_xor_sync(0xFFFFFFFFu, term, o);
				result += term;
			}
#else
			const int own_values[10] = { 0, 0, 19, 49, 76, 170, 33, 159, 252, 0 };
			const int opp_values[10] = { 0, 0, -1, -50, -45, -135, -14, -154, -496, 0 };
			for (int i = TT_OPEN_3; i <= TT_FIVE; i++)
				result += own_values[i] * v.count(v.own(), i) + opp_values[i] * v.count(v.opp(), i);
#endif
			result = result < -1000 ? -1000 : (result > 1000 ? 1000 : result);
			return sc_eval(result);
		}
		// AlphaBetaSearch::solve with a node limit of one: the root is generated and either proven statically or evaluated
		AGB_HD inline Result solve_static(const View &v, uint16_t *moves, uint16_t *scores)
		{
			MoveGenerator gen(v, moves, scores);
			gen.generate_optimal();
			if (not sc_is_proven(gen.out.score))
				gen.out.score = static_evaluation(v);
			return gen.out;
		}
	}
}
