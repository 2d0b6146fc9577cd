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
#include "solver_search.cuh"

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
		constexpr int kSolverSmemPerWarp = kLinePitch * 8 + kMaxCells * 6 + 96; // line words, pattern types (4 B), threats, board, list lengths of one position: 3520 B
		constexpr int kGreenResidentWarps = 28; // per SM inside the solver's green context (shared by the launches of all pipeline groups)
		constexpr int kResidentWarps = 10; // per SM, 72-register build: measured best of 4..28 (throughput is flat above ~8, the tail shorter below 28)
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
		template<int kSolverWarpsPerBlock, int kMinBlocks>
		__global__ void __launch_bounds__(kSolverWarpsPerBlock * 32, kMinBlocks) solve_games_kernel(BoardStore store, Tables tables, SolverState st, int game_begin, int games, int S, int rules,
				int draw_after, int max_nodes, SolverOutputs out, const uint8_t *__restrict__ slot_is_root, int *__restrict__ nn_list, int *__restrict__ nn_count,
				uint32_t *__restrict__ status)
		{
		  for (;;)
		  {
			int ticket = 0;
			if ((threadIdx.x & 31) == 0)
				ticket = atomicAdd(st.next + game_begin, 1);
			ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
			if (ticket >= games)
				return;
			const int local = st.order[game_begin + ticket];
			const int g = game_begin + local;
			const bool leader = (threadIdx.x & 31) == 0; // all lanes run the solver in lockstep (solver_search.cuh); one of them publishes
			const int cells = S * S;
			const int n_slots = st.game_slot_count[g];
			const long long t_begin = clock64();
			unsigned long long nodes_total = 0, n_adds = 0, n_quiet = 0, n_gen = 0;
			solver::HashTable tt { st.table + static_cast<size_t>(g) * st.table_entries * 2, st.table_entries / 4 - 1, st.generation[g], st.keys + static_cast<size_t>(g) * st.keys_stride };
			solver::SearchMemory mem { st.stack_moves + static_cast<size_t>(g) * st.stack_capacity, st.stack_scores + static_cast<size_t>(g) * st.stack_capacity,
					st.stack_capacity, reinterpret_cast<solver::Frame*>(st.frames) + static_cast<size_t>(g) * solver::kMaxFrames,
					reinterpret_cast<solver::ChildInfo*>(st.children) + static_cast<size_t>(g) * cells };
			// The position the search plays on -- board, line words, pattern types, threats: 3.4 KB, touched by every move made and taken back --
			// lives in this warp's shared memory for the duration of a solve: 28 warps' worth of it does not fit the L1 next to their lists,
			// tables and stacks (ncu: 66 % L1 hits, 8.8 stalled warp-cycles per instruction on loads), shared memory always hits. The slot's
			// global copy is never written back: the search leaves the position as it found it and the slot dies with the launch.
			extern __shared__ __align__(16) uint8_t solver_smem[];
			uint8_t *const my_smem = solver_smem + (threadIdx.x >> 5) * kSolverSmemPerWarp;
			uint64_t *const s_lines = reinterpret_cast<uint64_t*>(my_smem);
			uint32_t *const s_ptypes = reinterpret_cast<uint32_t*>(my_smem + kLinePitch * 8);
			uint8_t *const s_threats = my_smem + kLinePitch * 8 + kMaxCells * 4;
			int8_t *const s_board = reinterpret_cast<int8_t*>(my_smem + kLinePitch * 8 + kMaxCells * 5);
			int32_t *const s_hist_count = reinterpret_cast<int32_t*>(my_smem + kLinePitch * 8 + kMaxCells * 6);
			for (int k = 0; k < n_slots; k++)
			{
				const int slot = st.game_slots[static_cast<size_t>(g) * st.batch + k];
				const size_t cbase = static_cast<size_t>(slot) * kCellPitch;
				int stones = 0;
				__syncwarp();
				for (int i = threadIdx.x & 31; i < cells; i += 32) // the lanes take interleaved cells
				{
					const int8_t b = store.board[cbase + i];
					stones += (b != NONE);
					s_board[i] = b;
					s_ptypes[i] = store.ptypes[cbase + i];
					s_threats[i] = store.threats[cbase + i];
				}
				for (int i = threadIdx.x & 31; i < plogic::line_count(S); i += 32)
					s_lines[i] = store.lines[static_cast<size_t>(slot) * kLinePitch + i];
				if ((threadIdx.x & 31) < 2 * kHistTypes)
					s_hist_count[threadIdx.x & 31] = store.hist_count[static_cast<size_t>(slot) * 2 * kHistTypes + (threadIdx.x & 31)];
				for (int o = 16; o > 0; o >>= 1)
					stones += __shfl_xor_sync(0xFFFFFFFFu, stones, o);
				__syncwarp();
				solver::DynState d;
				d.board = s_board;
				d.lines = s_lines;
				d.ptypes = s_ptypes;
				d.threats = s_threats;
				d.hist_count = s_hist_count;
				d.hist_cells = store.hist_cells + static_cast<size_t>(slot) * 2 * kHistTypes * kCellPitch;
				d.threat_table = tables.threat;
				d.v = solver::View { S, cells, rules, store.sign_to_move[slot], stones, draw_after, kCellPitch, d.board, d.lines, d.ptypes, d.threats,
						store.forbidden + cbase, d.hist_count, d.hist_cells, tables.pattern, st.def_table, &d };
				solver::encode_forbidden_pass(d);
				const solver::SearchOutput res = solver::solve_position(d, tt, mem, max_nodes, 100);
				uint16_t *om = out.moves + static_cast<size_t>(slot) * out.pitch;
				uint16_t *os = out.scores + static_cast<size_t>(slot) * out.pitch;
				for (int i = threadIdx.x & 31; i < res.n_actions; i += 32)
				{
					om[i] = mem.stack_moves[i];
					os[i] = mem.stack_scores[i];
				}
				__syncwarp();
				out.n_actions[slot] = res.n_actions;
				out.score[slot] = res.score;
				out.must_defend[slot] = res.must_defend ? 1 : 0;
				out.nodes[slot] = res.node_counter;
				nodes_total += res.node_counter;
				n_adds += d.n_adds;
				n_quiet += d.n_quiet;
				n_gen += d.n_gen_actions;
				if (leader and res.overflow)
					atomicOr(status, res.overflow << 8); // bits 8..11, see AgbStats::overflow_flags
				if (leader and (slot_is_root[slot] or not solver::sc_is_proven(res.score)))
				{ // Search::scheduleToNN: roots and unproven positions go to the network; the evaluator draws their symmetry now, in task order
					nn_list[atomicAdd(nn_count, 1)] = slot;
					if (st.sym.task_sym != nullptr)
						st.sym.task_sym[slot] = static_cast<int8_t>(st.sym.draw(g));
				}
				__syncwarp();
			}
			if (leader)
			{
				st.game_cycles[2 * g] = static_cast<unsigned long long>(clock64() - t_begin);
				st.game_cycles[2 * g + 1] = nodes_total;
				// cost estimate in kilo-clocks of an undisturbed warp: least-squares fit on one-game-per-SM runs (tools/solver_work_fit.py, R^2 0.97)
				st.game_work[g] = static_cast<uint32_t>(24 * n_adds + 6 * n_quiet + 4 * n_gen);
			}
		  }
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
		const int draw_after = e->cfg.draw_after > 0 ? e->cfg.draw_after : e->cells;
		int sms = 148;
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->cfg.device);
		static const bool force_dense = getenv("AGB_SOLVER_DENSE") != nullptr; // tests: exercise the low-register build with few games
		static const int resident_env = getenv("AGB_SOLVER_RESIDENT") != nullptr ? atoi(getenv("AGB_SOLVER_RESIDENT")) : 0; // warps per SM (tuning)
		order_games_kernel<<<(game_count + 127) / 128, 128, 0, stream>>>(st.game_work, game_begin, game_count, st.order, st.next + game_begin);
		e->launches++;
		if (solver_sms > 0 and green)
		{ // the stream's green context holds solver_sms SMs: one-warp blocks, as many as fit (28 per SM at 72 registers); the launches of the other
		  // pipeline groups share these SMs, so a launch's tail (a few long games) runs next to the next group's games
			// resident warps per SM are capped by asking for (unused) dynamic shared memory: a warp among k unrelated ones runs at about 1 / k of the
			// SM's instruction supply, which saturates at 8-16 warps, and a launch ends with its longest game -- fewer residents shorten that chain
			const int resident = std::max(5, std::min(28, resident_env > 0 ? resident_env : kGreenResidentWarps));
			const int smem = resident >= 28 ? kSolverSmemPerWarp : (227 * 1024 / resident) & ~1023;
			solve_games_kernel<1, 28> <<<std::min(game_count, resident * solver_sms), 32, smem, stream>>>(e->store, e->tables, st, game_begin, game_count, e->cfg.rows,
					e->cfg.rules, draw_after, e->cfg.solver_max_positions, out, slot_is_root, nn_list, nn_count, e->d_status);
		}
		else if (solver_sms > 0)
		{ // side by side with the network kernel (AgbConfig::solver_sms): blocks of 28 warps at 72 registers fill an SM's register file, so such a
		  // block and a K4 CTA never share an SM, and they are launched as clusters of two so that they take whole TPCs and K4's CTA pairs find
		  // whole TPCs among the rest. Nothing inside the kernel uses the cluster.
			cudaLaunchConfig_t cfg = { };
			cfg.gridDim = dim3(static_cast<unsigned>(std::max(2, solver_sms & ~1)));
			cfg.blockDim = dim3(28 * 32);
			cfg.dynamicSmemBytes = 28 * kSolverSmemPerWarp;
			cfg.stream = stream;
			static const cudaError_t attr_set = cudaFuncSetAttribute(solve_games_kernel<28, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 28 * kSolverSmemPerWarp);
			AGB_CUDA_CHECK(e, attr_set);
			cudaLaunchAttribute attr[1];
			attr[0].id = cudaLaunchAttributeClusterDimension;
			attr[0].val.clusterDim.x = 2;
			attr[0].val.clusterDim.y = 1;
			attr[0].val.clusterDim.z = 1;
			cfg.attrs = attr;
			cfg.numAttrs = 1;
			AGB_CUDA_CHECK(e, cudaLaunchKernelEx(&cfg, solve_games_kernel<28, 1>, e->store, e->tables, st, game_begin, game_count, static_cast<int>(e->cfg.rows),
					static_cast<int>(e->cfg.rules), draw_after, static_cast<int>(e->cfg.solver_max_positions), out, slot_is_root, nn_list, nn_count, e->d_status));
		}
		else if (game_count <= 56 * sms and not force_dense)
		{
			const int resident = std::min(28, resident_env > 0 ? resident_env : kResidentWarps);
			solve_games_kernel<1, 28> <<<std::min(game_count, resident * sms), 32, kSolverSmemPerWarp, stream>>>(e->store, e->tables, st, game_begin, game_count, e->cfg.rows,
					e->cfg.rules, draw_after, e->cfg.solver_max_positions, out, slot_is_root, nn_list, nn_count, e->d_status);
		}
		else
		{
			const int resident = std::min(56, resident_env > 0 ? resident_env : 56);
			solve_games_kernel<2, 28> <<<std::min((game_count + 1) / 2, resident / 2 * sms), 64, 2 * kSolverSmemPerWarp, stream>>>(e->store, e->tables, st, game_begin, game_count,
					e->cfg.rows, e->cfg.rules, draw_after, e->cfg.solver_max_positions, out, slot_is_root, nn_list, nn_count, e->d_status);
		}
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
