// K5: the per-leaf solver of the lockstep engine. One warp owns one game and solves that game's leaf positions of
// the current batch one after another, in task order, on top of the K1 state of their slots: staged move generation, static
// evaluation and (max_positions > 1) the alpha-beta threat-space search with the game's transposition table
// (solver_logic.cuh, solver_search.cuh). Replaces Search::solve -> AlphaBetaSearch::solve
// (src/search/monte_carlo/Search.cpp:160-183, src/search/alpha_beta/AlphaBetaSearch.cpp:77-156) and, for the positions that
// stay unproven or are tree roots, Search::scheduleToNN (Search.cpp:184-198).
//
// The positions of one game must be solved in order because they share the table (what task i stores, task i+1 may read);
// games are independent, so the parallelism is one warp per game; its lanes run the sequential logic in lockstep and share the
// per-move pattern updates (solver_search.cuh).
#include "engine.hpp"
#include "solver_kernel.cuh" // the kernel and its launch shapes, in the build that knows renju's forbidden moves

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

namespace agb
{
	namespace
	{

		// Two builds of the same kernel: <1, 28> keeps 72 registers per thread (28 warps fit an SM); <2, 28> is capped at 32 registers and
		// holds 56, for launches with many more games than that (resident warps matter more than spills there: 16384 games x 2 leaves
		// take 26 ms where 4096 x 8 take 43).
		//
		// Scheduling: the kernel is bound by the SM's instruction supply, which saturates at 8-16 unrelated warps, and
		// a game's cost is heavy-tailed, so "every game resident from the start" ends in a long tail of a few expensive games per SM
		// crawling at an equal share. Instead fewer warps are resident and each pulls games from a queue ordered by the cost the game
		// had in the previous launch (longest first; a game's cost correlates 0.9 from one step to the next): expensive games start at once and get a larger share, cheap ones fill the end.
		__global__ void order_games_kernel(const uint32_t *__restrict__ game_work, int game_begin, int games, int32_t *__restrict__ order, int32_t *__restrict__ next)
		{ // rank by counting: games <= 16384, so n^2 comparisons are a few hundred microseconds at worst
			const int i = blockIdx.x * blockDim.x + threadIdx.x;
			if (i == 0)
				*next = 0;
			if (i >= games)
				return;
			const uint32_t mine = game_work[game_begin + i];
			int rank = 0;
			for (int j = 0; j < games; j++)
			{
				const uint32_t other = game_work[game_begin + j];
				rank += (other > mine or (other == mine and j < i)) ? 1 : 0;
			}
			order[game_begin + rank] = i;
		}
		__global__ void clear_tables_kernel(uint64_t *table, size_t n_entries)
		{
			for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_entries; i += static_cast<size_t>(gridDim.x) * blockDim.x)
			{
				table[2 * i] = 0;
				table[2 * i + 1] = solver::kEmptyEntryData;
			}
		}
		uint64_t splitmix64(uint64_t &x)
		{
			x += 0x9E3779B97F4A7C15ull;
			uint64_t z = x;
			z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
			z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
			return z ^ (z >> 31);
		}
	}

	int solver_create(AgbEngine *e)
	{
		std::vector<uint16_t> table(solver::kDefGroups * 256 * 2);
		solver::build::defensive_table(e->cfg.rules, table.data());
		AGB_CUDA_CHECK(e, cudaMalloc(&e->d_def_table, table.size() * sizeof(uint16_t)));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_def_table, table.data(), table.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, e->stream));
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		return AGB_OK;
	}
	// per-game search memory; called by selfplay_create when the solver is on
	int solver_state_create(AgbEngine *e, int games, int batch, SolverState *st)
	{
		static_assert(sizeof(solver::Frame) == SolverState::kFrameBytes, "frame size");
		const AgbConfig &c = e->cfg;
		size_t entries = c.solver_table_entries > 0 ? static_cast<size_t>(c.solver_table_entries) : 65536;
		if (entries < 4 or (entries & (entries - 1)) != 0)
			return e->fail(AGB_EINVAL, "solver_table_entries must be a power of two >= 4 (entries of the per-game transposition table)");
		st->games = games;
		st->batch = batch;
		st->table_entries = entries;
		st->stack_capacity = e->cells + 64 * 128; // the root list plus room for the deepest lines of a 100-node search
		AGB_CUDA_CHECK(e, cudaMalloc(&st->table, static_cast<size_t>(games) * entries * 16));
		AGB_CUDA_CHECK(e, cudaMalloc(&st->keys, static_cast<size_t>(e->cells) * 4 * sizeof(uint64_t)));
		st->keys_stride = 0;
		AGB_CUDA_CHECK(e, cudaMalloc(&st->generation, games * sizeof(int32_t)));
		AGB_CUDA_CHECK(e, cudaMalloc(&st->game_slots, static_cast<size_t>(games) * batch * sizeof(int32_t)));
		AGB_CUDA_CHECK(e, cudaMalloc(&st->game_slot_count, games * sizeof(int32_t)));
		AGB_CUDA_CHECK(e, cudaMalloc(&st->stack_moves, static_cast<size_t>(games) * st->stack_capacity * sizeof(uint16_t)));
		AGB_CUDA_CHECK(e, cudaMalloc(&st->stack_scores, static_cast<size_t>(games) * st->stack_capacity * sizeof(uint16_t)));
		AGB_CUDA_CHECK(e, cudaMalloc(&st->frames, static_cast<size_t>(games) * solver::kMaxFrames * sizeof(solver::Frame)));
		AGB_CUDA_CHECK(e, cudaMalloc(&st->game_cycles, static_cast<size_t>(games) * 2 * sizeof(unsigned long long)));
		AGB_CUDA_CHECK(e, cudaMemset(st->game_cycles, 0, static_cast<size_t>(games) * 2 * sizeof(unsigned long long)));
		AGB_CUDA_CHECK(e, cudaMalloc(&st->children, static_cast<size_t>(games) * e->cells * sizeof(solver::ChildInfo)));
		AGB_CUDA_CHECK(e, cudaMalloc(&st->game_work, games * sizeof(uint32_t)));
		AGB_CUDA_CHECK(e, cudaMemset(st->game_work, 0, games * sizeof(uint32_t)));
		AGB_CUDA_CHECK(e, cudaMalloc(&st->order, games * sizeof(int32_t)));
		AGB_CUDA_CHECK(e, cudaMalloc(&st->next, games * sizeof(int32_t)));
		AGB_CUDA_CHECK(e, cudaMemset(st->generation, 0, games * sizeof(int32_t)));
		AGB_CUDA_CHECK(e, cudaMemset(st->game_slot_count, 0, games * sizeof(int32_t)));
		st->def_table = e->d_def_table;
		// hash keys: any fixed random words do (they only decide the bucket mapping); agb_set_solver_keys replaces them
		std::vector<uint64_t> keys(static_cast<size_t>(e->cells) * 4);
		uint64_t x = c.seed ^ 0x5EEDFACE0C0FFEEull;
		for (uint64_t &k : keys)
			k = splitmix64(x);
		AGB_CUDA_CHECK(e, cudaMemcpy(st->keys, keys.data(), keys.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
		return solver_state_reset(e, st);
	}
	int solver_state_reset(AgbEngine *e, SolverState *st)
	{ // AlphaBetaSearch::clear + a fresh generation counter
		clear_tables_kernel<<<148 * 8, 256, 0, e->stream>>>(st->table, static_cast<size_t>(st->games) * st->table_entries);
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		AGB_CUDA_CHECK(e, cudaMemsetAsync(st->generation, 0, st->games * sizeof(int32_t), e->stream));
		return AGB_OK;
	}
	void solver_state_destroy(SolverState *st)
	{
		cudaFree(st->table);
		cudaFree(st->keys);
		cudaFree(st->generation);
		cudaFree(st->game_slots);
		cudaFree(st->game_slot_count);
		cudaFree(st->stack_moves);
		cudaFree(st->stack_scores);
		cudaFree(st->frames);
		cudaFree(st->children);
		cudaFree(st->game_cycles);
		cudaFree(st->game_work);
		cudaFree(st->order);
		cudaFree(st->next);
		*st = SolverState { };
	}
	int launch_solve_games(AgbEngine *e, const SolverState &st, int game_begin, int game_count, const SolverOutputs &out, const uint8_t *slot_is_root, int *nn_list,
			int *nn_count, cudaStream_t stream, int solver_sms, bool green)
	{
		order_games_kernel<<<(game_count + 127) / 128, 128, 0, stream>>>(st.game_work, game_begin, game_count, st.order, st.next + game_begin);
		e->launches++;
		const int rc = (e->cfg.rules == AGB_RENJU) ? launch_solve_kernels(e, st, game_begin, game_count, out, slot_is_root, nn_list, nn_count, stream, solver_sms, green)
				: launch_solve_kernels_plain(e, st, game_begin, game_count, out, slot_is_root, nn_list, nn_count, stream, solver_sms, green);
		if (rc != AGB_OK)
			return rc;
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		return AGB_OK;
	}
}

// ---- agb_solve: the solver on caller-supplied positions (AlphaBetaSearch::solve as a service) -----------------------------------------
namespace agb
{
	struct SolveScratch
	{
			SolverState state { };
			SolverOutputs out { };
			uint8_t *slot_is_root = nullptr;
			int32_t *nn_list = nullptr, *nn_count = nullptr;
			int capacity = 0;
	};
	namespace
	{
		__global__ void iota_kernel(int32_t *slots, int32_t *counts, int n)
		{
			const int i = blockIdx.x * blockDim.x + threadIdx.x;
			if (i < n)
			{
				slots[i] = i;
				counts[i] = 1;
			}
		}
		int solve_scratch_create(AgbEngine *e)
		{
			SolveScratch *sc = new SolveScratch();
			e->solve_scratch = sc;
			const size_t entries = e->cfg.solver_table_entries > 0 ? static_cast<size_t>(e->cfg.solver_table_entries) : 65536;
			const size_t by_memory = std::max<size_t>(1, (static_cast<size_t>(4) << 30) / (entries * 16)); // at most 4 GiB of tables
			sc->capacity = static_cast<int>(std::min<size_t>(std::min(e->cfg.max_boards, 2048), by_memory));
			const size_t n = sc->capacity, cells = e->cells;
			int rc = solver_state_create(e, sc->capacity, 1, &sc->state);
			if (rc != AGB_OK)
				return rc;
			sc->out.pitch = static_cast<int>(cells);
			AGB_CUDA_CHECK(e, cudaMalloc(&sc->out.moves, n * cells * sizeof(uint16_t)));
			AGB_CUDA_CHECK(e, cudaMalloc(&sc->out.scores, n * cells * sizeof(uint16_t)));
			AGB_CUDA_CHECK(e, cudaMalloc(&sc->out.n_actions, n * sizeof(int32_t)));
			AGB_CUDA_CHECK(e, cudaMalloc(&sc->out.score, n * sizeof(uint16_t)));
			AGB_CUDA_CHECK(e, cudaMalloc(&sc->out.must_defend, n));
			AGB_CUDA_CHECK(e, cudaMalloc(&sc->out.nodes, n * sizeof(int32_t)));
			AGB_CUDA_CHECK(e, cudaMalloc(&sc->slot_is_root, n));
			AGB_CUDA_CHECK(e, cudaMalloc(&sc->nn_list, n * sizeof(int32_t)));
			AGB_CUDA_CHECK(e, cudaMalloc(&sc->nn_count, sizeof(int32_t)));
			AGB_CUDA_CHECK(e, cudaMemsetAsync(sc->slot_is_root, 0, n, e->stream));
			iota_kernel<<<(sc->capacity + 255) / 256, 256, 0, e->stream>>>(sc->state.game_slots, sc->state.game_slot_count, sc->capacity);
			AGB_CUDA_CHECK(e, cudaGetLastError());
			if (not e->solver_keys_host.empty() and e->solver_keys_host.size() == static_cast<size_t>(e->cells) * 4)
				AGB_CUDA_CHECK(e, cudaMemcpyAsync(sc->state.keys, e->solver_keys_host.data(), e->solver_keys_host.size() * sizeof(uint64_t), cudaMemcpyHostToDevice,
						e->stream));
			AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
			return AGB_OK;
		}
	}
	void solve_scratch_destroy(AgbEngine *e)
	{
		SolveScratch *sc = e->solve_scratch;
		if (sc == nullptr)
			return;
		solver_state_destroy(&sc->state);
		void *ptrs[] = { sc->out.moves, sc->out.scores, sc->out.n_actions, sc->out.score, sc->out.must_defend, sc->out.nodes, sc->slot_is_root, sc->nn_list,
				sc->nn_count };
		for (void *ptr : ptrs)
			if (ptr)
				cudaFree(ptr);
		delete sc;
		e->solve_scratch = nullptr;
	}
}

extern "C" int agb_solve(AgbEngine *e, const int8_t *boards_host, const int8_t *sign_to_move_host, int n, int max_positions, uint16_t *scores_host,
		int32_t *n_actions_host, uint16_t *moves_host, uint16_t *action_scores_host, int32_t *flags_host)
{
	const agb::DeviceGuard on_device(e);
	using namespace agb;
	if (boards_host == nullptr or sign_to_move_host == nullptr or scores_host == nullptr)
		return e->fail(AGB_EINVAL, "null pointer");
	if (n < 0 or max_positions < 1)
		return e->fail(AGB_EINVAL, "n must be >= 0 and max_positions >= 1");
	if (e->solve_scratch == nullptr)
	{
		const int rc = solve_scratch_create(e);
		if (rc != AGB_OK)
			return rc;
	}
	SolveScratch *sc = e->solve_scratch;
	const size_t cells = e->cells;
	const int rc_valid = validate_boards(e, boards_host, sign_to_move_host, static_cast<size_t>(n));
	if (rc_valid != AGB_OK)
		return rc_valid;
	const int saved = e->cfg.solver_max_positions;
	for (int begin = 0; begin < n; begin += sc->capacity)
	{
		const int count = std::min(sc->capacity, n - begin);
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io8, boards_host + begin * cells, count * cells, cudaMemcpyHostToDevice, e->stream));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io8b, sign_to_move_host + begin, count, cudaMemcpyHostToDevice, e->stream));
		int rc = launch_set_boards(e, e->d_io8, e->d_io8b, count, e->d_features); // K1 (+K3)
		if (rc != AGB_OK)
			return rc;
		rc = solver_state_reset(e, &sc->state); // every position starts from a cleared table (AlphaBetaSearch::clear)
		if (rc != AGB_OK)
			return rc;
		AGB_CUDA_CHECK(e, cudaMemsetAsync(sc->nn_count, 0, sizeof(int32_t), e->stream));
		e->cfg.solver_max_positions = max_positions;
		rc = launch_solve_games(e, sc->state, 0, count, sc->out, sc->slot_is_root, sc->nn_list, sc->nn_count, e->stream);
		e->cfg.solver_max_positions = saved;
		if (rc != AGB_OK)
			return rc;
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(scores_host + begin, sc->out.score, count * sizeof(uint16_t), cudaMemcpyDeviceToHost, e->stream));
		if (n_actions_host)
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(n_actions_host + begin, sc->out.n_actions, count * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
		if (moves_host)
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(moves_host + begin * cells, sc->out.moves, count * cells * sizeof(uint16_t), cudaMemcpyDeviceToHost, e->stream));
		if (action_scores_host)
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(action_scores_host + begin * cells, sc->out.scores, count * cells * sizeof(uint16_t), cudaMemcpyDeviceToHost, e->stream));
		if (flags_host)
		{ // bit0 must_defend, bits 8.. positions visited
			std::vector<uint8_t> md(count);
			std::vector<int32_t> nodes(count);
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(md.data(), sc->out.must_defend, count, cudaMemcpyDeviceToHost, e->stream));
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(nodes.data(), sc->out.nodes, count * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
			AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
			for (int i = 0; i < count; i++)
				flags_host[begin + i] = static_cast<int32_t>(md[i]) | (nodes[i] << 8);
		}
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
	}
	uint32_t status = 0;
	const int rc_status = take_status(e, &status);
	if (rc_status != AGB_OK)
		return rc_status;
	if (status != 0)
		return e->fail(AGB_EOVERFLOW, "device-side overflow, flags=" + std::to_string(status));
	return AGB_OK;
}
