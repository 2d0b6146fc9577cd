// TEST-ONLY: compiles the host+device logic headers of alphagomoku_b200/csrc with g++ so the kernel logic can be
// checked against the oracle on a box without a GPU (tests marked "not gpu"). The product never builds or loads this.
#include "../../alphagomoku_b200/csrc/tables_logic.cuh"

#include <cstdint>
#include <vector>

using namespace agb;
using namespace agb::tables_logic;

extern "C" void hostsim_pattern_table(int rules, uint8_t *table)
{
	RuleBuilder cross { rules, CROSS, { } }, circle { rules, CIRCLE, { } };
	cross.build();
	circle.build();
	for (uint32_t i = 0; i < (1u << 20); i++)
		table[i] = pattern_table_entry(i, cross.out.data(), (int) cross.out.size(), circle.out.data(), (int) circle.out.size());
}
extern "C" void hostsim_threat_table(int rules, uint8_t *table)
{
	for (int idx = 0; idx < 4096; idx++)
	{
		const int t[4] = { idx & 7, (idx >> 3) & 7, (idx >> 6) & 7, (idx >> 9) & 7 };
		int cross, circle;
		threat_of(t, rules, cross, circle);
		table[idx] = static_cast<uint8_t>(cross | (circle << 4));
	}
}

#include "../../alphagomoku_b200/csrc/patterns_logic.cuh"
using namespace agb::plogic;

// sequential model of the K1+K3 kernel (one board): same per-cell functions, plain loops instead of a warp
extern "C" void hostsim_set_board(int rules, int S, const int8_t *board, int stm, const uint8_t *pattern_table, const uint8_t *threat_table,
		uint8_t *ptypes_out, uint8_t *threats_out, uint32_t *features_out, uint8_t *forbidden_out, int *overflow)
{
	uint64_t lines[kMaxLines];
	for (int l = 0; l < line_count(S); l++)
		lines[l] = build_line(board, S, l);
	Tables tables { pattern_table, threat_table };
	*overflow = 0;
	for (int r = 0; r < S; r++)
		for (int c = 0; c < S; c++)
		{
			const int idx = r * S + c;
			const uint32_t p = (board[idx] == NONE) ? classify_cell(lines, pattern_table, r, c, S) : 0u;
			const uint8_t t = (board[idx] == NONE) ? threat_of_cell(p, threat_table) : 0;
			for (int d = 0; d < 4; d++)
				ptypes_out[4 * idx + d] = (p >> (8 * d)) & 0x77;
			threats_out[2 * idx + 0] = t & 15;
			threats_out[2 * idx + 1] = t >> 4;
			uint32_t f = encode_cell(board[idx], p, stm);
			bool forb = false;
			if (rules == RULE_RENJU and board[idx] == NONE)
			{
				const int tc = t & 15;
				if (tc == TT_OVERLINE or tc == TT_FORK_4x4)
					forb = true;
				else if (tc == TT_FORK_3x3)
				{
					Overlay ov;
					forb = is_forbidden_raw(board, S, r, c, tables, ov);
					*overflow |= ov.overflow;
				}
			}
			forbidden_out[idx] = forb;
			if (forb and stm == CROSS)
				f |= 1u << 6;
			features_out[idx] = f;
		}
}
extern "C" uint32_t hostsim_open3_promotions(uint32_t window)
{
	return open_three_promotions(window);
}
extern "C" int hostsim_outcome(int rules, int S, const int8_t *board, int r, int c, int sign, int draw_after, const uint8_t *pattern_table,
		const uint8_t *threat_table)
{
	Tables tables { pattern_table, threat_table };
	bool overflow = false;
	return outcome_of(board, S, rules, draw_after, r, c, sign, tables, overflow);
}
extern "C" void hostsim_augment(uint32_t *dst, const uint32_t *src, int S, int mode)
{
	for (int r = 0; r < S; r++)
		for (int c = 0; c < S; c++)
		{
			int sr, sc;
			symmetry_source(mode, S, r, c, sr, sc);
			dst[r * S + c] = permute_direction_bits(src[sr * S + sc], mode);
		}
}

#include "../../alphagomoku_b200/csrc/records.cuh"
extern "C" size_t hostsim_serialize_sample_v201(int cells, const int8_t *board, const int32_t *visits, const float *prior, const float *win,
		const float *draw, const uint16_t *scores, uint16_t minimax_score, uint16_t flags, uint8_t *out)
{
	return agb::records::serialize_sample_v201(out, cells, board, visits, prior, win, draw, scores, minimax_score, flags);
}

#include "../../alphagomoku_b200/csrc/solver_search.cuh"
// sequential model of "K1 state -> K5 static solve" for one position (the device runs the same functions, one thread per task)
extern "C" void hostsim_defensive_table(int rules, uint16_t *table)
{
	agb::solver::build::defensive_table(rules, table);
}
extern "C" uint32_t hostsim_defensive_mask(const uint16_t *table, int rules, uint32_t window13, int defender, int threat)
{
	return agb::solver::defensive_mask(table, rules, window13, defender, threat);
}
extern "C" int hostsim_solve_static(int rules, int S, int draw_after, const int8_t *board, int stm, const uint8_t *pattern_table, const uint8_t *threat_table,
		const uint16_t *def_table, uint16_t *moves, uint16_t *scores, uint16_t *result_score, int32_t *flags)
{
	using namespace agb::solver;
	const int cells = S * S;
	uint64_t lines[kMaxLines];
	for (int l = 0; l < line_count(S); l++)
		lines[l] = build_line(board, S, l);
	std::vector<uint32_t> ptypes(cells, 0);
	std::vector<uint8_t> threats(cells, 0), forbidden(cells, 0);
	std::vector<int32_t> hist_count(2 * kHistTypes, 0);
	std::vector<uint16_t> hist_cells(2 * kHistTypes * cells, 0);
	Tables tables { pattern_table, threat_table };
	int stones = 0;
	for (int r = 0; r < S; r++)
		for (int c = 0; c < S; c++)
		{
			const int idx = r * S + c;
			stones += (board[idx] != NONE);
			if (board[idx] != NONE)
				continue;
			ptypes[idx] = classify_cell(lines, pattern_table, r, c, S);
			threats[idx] = threat_of_cell(ptypes[idx], threat_table);
			for (int colour = 0; colour < 2; colour++)
			{
				const int t = (threats[idx] >> (4 * colour)) & 15;
				if (t != TT_NONE)
					hist_cells[(colour * kHistTypes + t) * cells + hist_count[colour * kHistTypes + t]++] = mk_loc(r, c);
			}
			if (rules == RULE_RENJU)
			{
				const int tc = threats[idx] & 15;
				if (tc == TT_OVERLINE or tc == TT_FORK_4x4)
					forbidden[idx] = 1;
				else if (tc == TT_FORK_3x3)
				{
					Overlay ov;
					forbidden[idx] = is_forbidden_raw(board, S, r, c, tables, ov);
				}
			}
		}
	View v { S, cells, rules, stm, stones, draw_after > 0 ? draw_after : cells, cells, board, lines, ptypes.data(), threats.data(), forbidden.data(),
			hist_count.data(), hist_cells.data(), pattern_table, def_table };
	const Result res = solve_static(v, moves, scores);
	*result_score = res.score;
	*flags = static_cast<int>(res.must_defend) | (static_cast<int>(res.has_initiative) << 3);
	return res.n_actions;
}

// sequential model of one game's solver: K1 state -> K5 search (solve_position) with a persistent transposition table
struct HostSolver
{
		int rules, S, draw_after;
		std::vector<uint8_t> pattern_table, threat_table;
		std::vector<uint16_t> def_table;
		std::vector<uint64_t> keys, table;
		int generation = 0;
		std::vector<uint16_t> stack_moves, stack_scores;
		std::vector<agb::solver::Frame> frames;
		std::vector<agb::solver::ChildInfo> children;
};
extern "C" void* hostsim_solver_create(int rules, int S, int draw_after, const uint8_t *pattern_table, const uint8_t *threat_table, const uint16_t *def_table,
		const uint64_t *keys, size_t table_entries)
{
	HostSolver *h = new HostSolver();
	h->rules = rules;
	h->S = S;
	h->draw_after = draw_after > 0 ? draw_after : S * S;
	h->pattern_table.assign(pattern_table, pattern_table + (1 << 20));
	h->threat_table.assign(threat_table, threat_table + 4096);
	h->def_table.assign(def_table, def_table + agb::solver::kDefGroups * 256 * 2);
	h->keys.assign(keys, keys + 4 * S * S);
	h->table.resize(2 * table_entries);
	agb::solver::tt_clear(h->table.data(), table_entries);
	h->stack_moves.resize(S * S + 8192);
	h->stack_scores.resize(S * S + 8192);
	h->frames.resize(agb::solver::kMaxFrames);
	h->children.resize(S * S);
	return h;
}
extern "C" void hostsim_solver_destroy(void *p)
{
	delete static_cast<HostSolver*>(p);
}
extern "C" void hostsim_solver_next_generation(void *p)
{
	HostSolver *h = static_cast<HostSolver*>(p);
	h->generation = (h->generation + 1) % 64;
}
extern "C" void hostsim_solver_clear(void *p)
{
	HostSolver *h = static_cast<HostSolver*>(p);
	agb::solver::tt_clear(h->table.data(), h->table.size() / 2);
}
extern "C" int hostsim_solver_solve(void *p, const int8_t *board_in, int stm, int max_nodes, uint16_t *moves, uint16_t *scores, uint16_t *result_score,
		int32_t *flags)
{
	using namespace agb::solver;
	HostSolver *h = static_cast<HostSolver*>(p);
	const int S = h->S, cells = S * S;
	std::vector<int8_t> board(board_in, board_in + cells);
	std::vector<uint64_t> lines(kMaxLines);
	for (int l = 0; l < line_count(S); l++)
		lines[l] = build_line(board.data(), S, l);
	std::vector<uint32_t> ptypes(cells, 0);
	std::vector<uint8_t> threats(cells, 0);
	std::vector<int32_t> hist_count(2 * kHistTypes, 0);
	std::vector<uint16_t> hist_cells(2 * kHistTypes * cells, 0);
	int stones = 0;
	for (int r = 0; r < S; r++)
		for (int c = 0; c < S; c++)
		{ // PatternCalculator::setBoard: classify, then fill the lists in row-major order
			const int idx = r * S + c;
			stones += (board[idx] != NONE);
			if (board[idx] != NONE)
				continue;
			ptypes[idx] = classify_cell(lines.data(), h->pattern_table.data(), r, c, S);
			threats[idx] = threat_of_cell(ptypes[idx], h->threat_table.data());
			for (int colour = 0; colour < 2; colour++)
			{
				const int t = (threats[idx] >> (4 * colour)) & 15;
				if (t != TT_NONE)
					hist_cells[(colour * kHistTypes + t) * cells + hist_count[colour * kHistTypes + t]++] = mk_loc(r, c);
			}
		}
	DynState d;
	d.v = View { S, cells, h->rules, stm, stones, h->draw_after, cells, board.data(), lines.data(), ptypes.data(), threats.data(), nullptr, hist_count.data(),
			hist_cells.data(), h->pattern_table.data(), h->def_table.data(), &d };
	d.board = board.data();
	d.lines = lines.data();
	d.ptypes = ptypes.data();
	d.threats = threats.data();
	d.hist_count = hist_count.data();
	d.hist_cells = hist_cells.data();
	d.threat_table = h->threat_table.data();
	if (h->rules == RULE_RENJU)
	{ // NNInputFeatures::encode runs first in AlphaBetaSearch::solve and asks isForbidden for every empty cell in row-major order
		encode_forbidden_pass(d);
	}
	HashTable tt { h->table.data(), h->table.size() / 8 - 1, h->generation, h->keys.data() };
	SearchMemory mem { h->stack_moves.data(), h->stack_scores.data(), static_cast<int>(h->stack_moves.size()), h->frames.data(), h->children.data() };
	const SearchOutput out = solve_position(d, tt, mem, max_nodes, 100);
	for (int i = 0; i < out.n_actions; i++)
	{
		moves[i] = mem.stack_moves[i];
		scores[i] = mem.stack_scores[i];
	}
	*result_score = out.score;
	*flags = static_cast<int>(out.must_defend) | (static_cast<int>(out.node_counter <= 1) << 1) | (static_cast<int>(sc_is_proven(out.score)) << 2)
			| (out.node_counter << 8) | (out.overflow ? (1 << 30) : 0);
	return out.n_actions;
}

// a22: prepareOpening with a std::mt19937 seeded like the reference's debug build (seed 0)
#include "../../alphagomoku_b200/csrc/openings_logic.hpp"
extern "C" void* hostsim_openings_create(uint32_t seed)
{
	return new agb::openings::Random(seed);
}
extern "C" void hostsim_openings_destroy(void *p)
{
	delete static_cast<agb::openings::Random*>(p);
}
extern "C" int hostsim_prepare_opening(void *p, int rules, int S, int min_moves, const uint8_t *pattern_table, const uint8_t *threat_table, uint16_t *moves)
{
	std::vector<int8_t> board;
	const agb::Tables tables { pattern_table, threat_table };
	const std::vector<uint16_t> result = agb::openings::prepare_opening(rules, S, S, tables, *static_cast<agb::openings::Random*>(p), min_moves, board);
	std::copy(result.begin(), result.end(), moves);
	return static_cast<int>(result.size());
}

// MoveGenerator::generate(THREATS or OPTIMAL) on a position set up like PatternCalculator::setBoard; flags: bit0 must_defend, bit1 has_initiative
extern "C" int hostsim_generate(int rules, int S, const int8_t *board_in, int stm, int mode, const uint8_t *pattern_table, const uint8_t *threat_table,
		const uint16_t *def_table, uint16_t *moves, uint16_t *scores, int32_t *flags)
{
	using namespace agb::solver;
	const int cells = S * S;
	std::vector<int8_t> board(board_in, board_in + cells);
	std::vector<uint64_t> lines(kMaxLines);
	for (int l = 0; l < line_count(S); l++)
		lines[l] = build_line(board.data(), S, l);
	std::vector<uint32_t> ptypes(cells, 0);
	std::vector<uint8_t> threats(cells, 0);
	std::vector<int32_t> hist_count(2 * kHistTypes, 0);
	std::vector<uint16_t> hist_cells(2 * kHistTypes * cells, 0);
	int stones = 0;
	for (int r = 0; r < S; r++)
		for (int c = 0; c < S; c++)
		{
			const int idx = r * S + c;
			stones += (board[idx] != NONE);
			if (board[idx] != NONE)
				continue;
			ptypes[idx] = classify_cell(lines.data(), pattern_table, r, c, S);
			threats[idx] = threat_of_cell(ptypes[idx], threat_table);
			for (int colour = 0; colour < 2; colour++)
			{
				const int t = (threats[idx] >> (4 * colour)) & 15;
				if (t != TT_NONE)
					hist_cells[(colour * kHistTypes + t) * cells + hist_count[colour * kHistTypes + t]++] = mk_loc(r, c);
			}
		}
	DynState d;
	d.v = View { S, cells, rules, stm, stones, cells, cells, board.data(), lines.data(), ptypes.data(), threats.data(), nullptr, hist_count.data(), hist_cells.data(),
			pattern_table, def_table, &d };
	d.board = board.data();
	d.lines = lines.data();
	d.ptypes = ptypes.data();
	d.threats = threats.data();
	d.hist_count = hist_count.data();
	d.hist_cells = hist_cells.data();
	d.threat_table = threat_table;
	MoveGenerator gen(d.v, moves, scores);
	gen.generate(mode);
	*flags = static_cast<int>(gen.out.must_defend) | (static_cast<int>(gen.out.has_initiative) << 1);
	return gen.out.n_actions;
}

// Score arithmetic of the search (search/Score.hpp) as the device uses it
extern "C" uint16_t hostsim_score_negate(uint16_t s)
{
	return agb::solver::sc_negate(s);
}
extern "C" uint16_t hostsim_score_invert(uint16_t s, int delta)
{
	return agb::solver::sc_invert(s, delta);
}
extern "C" int hostsim_score_is_proven(uint16_t s)
{
	return agb::solver::sc_is_proven(s) ? 1 : 0;
}
