// Classification logic of the static tables as host+device inline functions, so that tables.cu (device build) and
// the host-compiled logic tests in tests/hostsim share one definition.
#pragma once
#include "agb_common.cuh"

#include <initializer_list>
#include <vector>

namespace agb
{
	namespace tables_logic
	{
		struct ShapeRule
		{
				uint8_t length;
				uint8_t type; // PatternType it proves
				uint8_t allowed[11]; // per cell: bit s set <=> Sign s may stand there
		};
		constexpr int kMaxRules = 64;

		// ---- host: shape rules ------------------------------------------------------------------------------
		constexpr uint8_t EMPTY = 1u << NONE, X = 1u << CROSS, O = 1u << CIRCLE, WALL = 1u << ILLEGAL, ANY = 0xF;

		struct RuleBuilder
		{
				int rules;
				int colour; // CROSS or CIRCLE
				std::vector<ShapeRule> out;

				uint8_t own() const { return colour == CROSS ? X : O; }
				uint8_t opp() const { return colour == CROSS ? O : X; }
				bool exact_five() const
				{ // "five but not overline": both colours in STANDARD, black only in RENJU (PatternClassifier.cpp:191-209)
					return rules == RULE_STANDARD or (rules == RULE_RENJU and colour == CROSS);
				}
				std::vector<uint8_t> cells(const char *shape) const
				{ // shapes are written for cross: 'X' own stone, '_' empty
					std::vector<uint8_t> r;
					for (const char *p = shape; *p; p++)
						r.push_back(*p == 'X' ? own() : EMPTY);
					return r;
				}
				void emit(int type, uint8_t left, const std::vector<uint8_t> &core, uint8_t right, bool has_left, bool has_right)
				{
					ShapeRule s { };
					s.type = static_cast<uint8_t>(type);
					int n = 0;
					if (has_left)
						s.allowed[n++] = left;
					for (uint8_t c : core)
						s.allowed[n++] = c;
					if (has_right)
						s.allowed[n++] = right;
					s.length = static_cast<uint8_t>(n);
					out.push_back(s);
				}
				// classes whose shapes have a stone (or the to-be-five gap) at an end: FIVE, HALF_OPEN_4, HALF_OPEN_3
				void add_end_sensitive(int type, std::initializer_list<const char*> shapes)
				{
					const uint8_t not_own = ANY & ~own(), not_opp = ANY & ~opp(), open = EMPTY | WALL;
					for (const char *sh : shapes)
					{
						const std::vector<uint8_t> core = cells(sh);
						if (exact_five())
							emit(type, not_own, core, not_own, true, true);
						else if (rules == RULE_CARO5)
						{ // at least one end not blocked by the opponent and no overline on the other
							emit(type, open, core, not_own, true, true);
							emit(type, not_own, core, open, true, true);
						}
						else if (rules == RULE_CARO6)
						{
							emit(type, not_opp, core, ANY, true, true);
							emit(type, ANY, core, not_opp, true, true);
						}
						else
							emit(type, 0, core, 0, false, false);
					}
				}
				// classes whose shapes already end in empty cells: OPEN_4, DOUBLE_4, OPEN_3
				void add_open(int type, std::initializer_list<const char*> shapes)
				{
					const uint8_t not_own = ANY & ~own(), not_opp = ANY & ~opp(), open = EMPTY | WALL;
					for (const char *sh : shapes)
					{
						const std::vector<uint8_t> core = cells(sh);
						if (exact_five())
							emit(type, not_own, core, not_own, true, true);
						else if (rules == RULE_CARO6)
							emit(type, not_opp, core, not_opp, true, true);
						else if (rules == RULE_CARO5)
							emit(type, open, core, open, true, true);
						else
							emit(type, 0, core, 0, false, false);
					}
				}
				void build()
				{ // priority order of PatternTable.cpp:49-66
					add_end_sensitive(PT_FIVE, { "XXXXX" });
					emit(PT_OVERLINE, 0, cells("XXXXXX"), 0, false, false);
					add_open(PT_OPEN_4, { "_XXXX_" });
					add_open(PT_DOUBLE_4, { "X_XXX_X", "XX_XX_XX", "XXX_X_XXX" });
					add_end_sensitive(PT_HALF_OPEN_4, { "_XXXX", "X_XXX", "XX_XX", "XXX_X", "XXXX_" });
					add_open(PT_OPEN_3, { "_XXX__", "_XX_X_", "_X_XX_", "__XXX_" });
					add_end_sensitive(PT_HALF_OPEN_3, { "__XXX", "_X_XX", "_XX_X", "_XXX_", "X__XX", "X_X_X", "X_XX_", "XX__X", "XX_X_", "XXX__" });
				}
		};

		// ---- device: one window per thread ------------------------------------------------------------------
		AGB_HD inline uint32_t expand_index(uint32_t i)
		{ // insert an empty centre cell: 20-bit index -> 22-bit window (PatternTable.hpp:142-145)
			return (i & 1023u) | ((i & 0xFFC00u) << 2);
		}
		AGB_HD inline uint32_t reverse_window(uint32_t w)
		{ // mirror the 11 two-bit cells
			uint32_t r = 0;
#pragma unroll
			for (int k = 0; k < 11; k++)
				r |= ((w >> (2 * k)) & 3u) << (2 * (10 - k));
			return r;
		}
		AGB_HD inline bool window_is_possible(uint32_t w)
		{ // walls only as a contiguous run from either end (Pattern.hpp:50-61); the centre is empty by construction
#pragma unroll
			for (int k = 0; k < 5; k++)
				if (((w >> (2 * k)) & 3u) != ILLEGAL and ((w >> (2 * k + 2)) & 3u) == ILLEGAL)
					return false;
#pragma unroll
			for (int k = 6; k < 11; k++)
				if (((w >> (2 * k - 2)) & 3u) == ILLEGAL and ((w >> (2 * k)) & 3u) != ILLEGAL)
					return false;
			return true;
		}
		AGB_HD inline int classify(uint32_t w, int colour, const ShapeRule *rules, int n)
		{
			uint8_t bits[11]; // one-hot sign per cell, centre taken by `colour`
#pragma unroll
			for (int k = 0; k < 11; k++)
				bits[k] = 1u << ((k == 5) ? colour : ((w >> (2 * k)) & 3u));
			for (int r = 0; r < n; r++)
			{
				const ShapeRule &rule = rules[r];
				for (int start = 0; start + rule.length <= 11; start++)
				{
					bool ok = true;
					for (int j = 0; j < rule.length; j++)
						ok = ok and ((rule.allowed[j] & bits[start + j]) != 0);
					if (ok)
						return rule.type;
				}
			}
			return PT_NONE;
		}
		AGB_HD inline uint8_t pattern_table_entry(uint32_t i, const ShapeRule *cross_rules, int n_cross, const ShapeRule *circle_rules, int n_circle)
		{
			// a window and its mirror image share one entry, computed from the smaller index (PatternTable.cpp:155-189)
			const uint32_t mirrored = narrow_window(reverse_window(expand_index(i)));
			const uint32_t w = expand_index(i < mirrored ? i : mirrored);
			uint8_t entry = 0;
			if (window_is_possible(w))
			{
				const int cross = classify(w, CROSS, cross_rules, n_cross);
				const int circle = classify(w, CIRCLE, circle_rules, n_circle);
				// half-open threes live in a side bit, the main type is demoted to NONE (PatternTable.cpp:172-183)
				entry = (cross == PT_HALF_OPEN_3) ? 0x08 : cross;
				entry |= (circle == PT_HALF_OPEN_3) ? 0x80 : (circle << 4);
			}
			return entry;
		}
		// ---- host: 4 pattern types -> threat (ThreatTable.cpp:52-96) -----------------------------------------
		inline void threat_of(const int t[4], int rules, int &for_cross, int &for_circle)
		{
			int n[8] = { 0 };
			for (int d = 0; d < 4; d++)
				n[t[d]]++;
			const int fours = n[PT_OPEN_4] + n[PT_HALF_OPEN_4];
			const bool fork44 = n[PT_DOUBLE_4] > 0 or fours >= 2;
			const bool fork43 = n[PT_OPEN_3] >= 1 and fours >= 1;
			const bool fork33 = n[PT_OPEN_3] >= 2;
			const auto both = [&](int v) { for_cross = for_circle = v; };
			if (n[PT_FIVE])
				return both(TT_FIVE);
			if (rules == RULE_RENJU)
			{ // columns: (black, white)
				if (n[PT_OVERLINE])
				{
					for_cross = TT_OVERLINE;
					for_circle = TT_FIVE;
					return;
				}
				if (fork44)
					return both(TT_FORK_4x4);
				if (n[PT_OPEN_4])
				{
					for_cross = fork33 ? TT_FORK_3x3 : TT_OPEN_4;
					for_circle = TT_OPEN_4;
					return;
				}
				if (fork43)
				{
					for_cross = fork33 ? TT_FORK_3x3 : TT_FORK_4x3;
					for_circle = TT_FORK_4x3;
					return;
				}
			}
			else
			{
				if (fork44)
					return both(TT_FORK_4x4);
				if (n[PT_OPEN_4])
					return both(TT_OPEN_4);
				if (fork43)
					return both(TT_FORK_4x3);
			}
			if (fork33)
				return both(TT_FORK_3x3);
			if (n[PT_HALF_OPEN_4])
				return both(TT_HALF_OPEN_4);
			if (n[PT_OPEN_3])
				return both(TT_OPEN_3);
			if (n[PT_HALF_OPEN_3])
				return both(TT_HALF_OPEN_3);
			both(TT_NONE);
		}
	}
}
