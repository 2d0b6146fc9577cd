// K5: the solver kernel and its launch shapes. Included by two translation units: solver.cu (as is: the build that evaluates renju's forbidden
// moves, launch_solve_kernels) and solver_plain.cu (with AGB_SOLVER_NO_RENJU: every forbidden-move branch, the replay of isForbidden's side
// effects and their code are compiled out, launch_solve_kernels_plain). The kernel is bound by instruction supply: the plain
// build is a fifth smaller (6.6 k SASS instructions at the end of round 2) and serves freestyle, standard and caro.
#pragma once
#include "engine.hpp"
#include "solver_search.cuh"

#include <algorithm>
#include <cstdlib>

#ifdef AGB_SOLVER_NO_RENJU
#define AGB_SOLVER_LAUNCH launch_solve_kernels_plain
#else
#define AGB_SOLVER_LAUNCH launch_solve_kernels
#endif

namespace agb
{
	namespace solver_kernel = AGB_SOLVER_NS;
	int launch_solve_kernels(AgbEngine *e, const SolverState &st, int game_begin, int game_count, const SolverOutputs &out, const uint8_t *slot_is_root, int *nn_list,
			int *nn_count, cudaStream_t stream, int solver_sms, bool green);
	int launch_solve_kernels_plain(AgbEngine *e, const SolverState &st, int game_begin, int game_count, const SolverOutputs &out, const uint8_t *slot_is_root,
			int *nn_list, int *nn_count, cudaStream_t stream, int solver_sms, bool green);
	namespace
	{
		constexpr int kSolverSmemPerWarp = solver_kernel::position_layout::kBytes; // line words, pattern types (4 B), threats, board, list lengths and list index of one position: 5120 B
		constexpr int kGreenResidentWarps = 28; // per SM inside the solver's green context (shared by the launches of all pipeline groups)
		constexpr int kResidentWarps = 10; // per SM, 72-register build: measured best of 4..28 (throughput is flat above ~8, the tail shorter below 28)
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
			solver_kernel::HashTable tt { st.table + static_cast<size_t>(g) * st.table_entries * 2, st.table_entries / 4 - 1, st.generation[g], st.keys + static_cast<size_t>(g) * st.keys_stride };
			solver_kernel::SearchMemory mem { st.stack_moves + static_cast<size_t>(g) * st.stack_capacity, st.stack_scores + static_cast<size_t>(g) * st.stack_capacity,
					st.stack_capacity, reinterpret_cast<solver_kernel::Frame*>(st.frames) + static_cast<size_t>(g) * solver_kernel::kMaxFrames,
					reinterpret_cast<solver_kernel::ChildInfo*>(st.children) + static_cast<size_t>(g) * cells };
			// The position the search plays on -- board, line words, pattern types, threats: 3.4 KB, touched by every move made and taken back --
			// lives in this warp's shared memory for the duration of a solve: 28 warps' worth of it does not fit the L1 next to their lists,
			// tables and stacks (ncu: 66 % L1 hits, 8.8 stalled warp-cycles per instruction on loads), shared memory always hits. The slot's
			// global copy is never written back: the search leaves the position as it found it and the slot dies with the launch.
			extern __shared__ __align__(16) uint8_t solver_smem[];
			namespace layout = solver_kernel::position_layout;
			uint8_t *const my_smem = solver_smem + (threadIdx.x >> 5) * kSolverSmemPerWarp;
			uint64_t *const s_lines = reinterpret_cast<uint64_t*>(my_smem + layout::kLines);
			uint32_t *const s_ptypes = reinterpret_cast<uint32_t*>(my_smem + layout::kPtypes);
			uint8_t *const s_threats = my_smem + layout::kThreats;
			int8_t *const s_board = reinterpret_cast<int8_t*>(my_smem + layout::kBoard);
			int32_t *const s_hist_count = reinterpret_cast<int32_t*>(my_smem + layout::kHistCount);
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
				solver_kernel::DynState d;
				d.board = s_board;
				d.lines = s_lines;
				d.ptypes = s_ptypes;
				d.threats = s_threats;
				d.hist_count = s_hist_count;
				d.hist_cells = store.hist_cells + static_cast<size_t>(slot) * 2 * kHistTypes * kCellPitch;
				d.threat_table = tables.threat;
				d.v = solver_kernel::View { S, cells, rules, store.sign_to_move[slot], stones, draw_after, kCellPitch, d.board, d.lines, d.ptypes, d.threats,
						store.forbidden + cbase, d.hist_count, d.hist_cells, tables.pattern, st.def_table, &d };
#ifdef __CUDA_ARCH__
				solver_kernel::dyn_build_list_index(d);
#endif
				solver_kernel::encode_forbidden_pass(d);
				const solver_kernel::SearchOutput res = solver_kernel::solve_position(d, tt, mem, max_nodes, 100);
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
				if (leader and (slot_is_root[slot] or not solver_kernel::sc_is_proven(res.score)))
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
	}
	int AGB_SOLVER_LAUNCH(AgbEngine *e, const SolverState &st, int game_begin, int game_count, const SolverOutputs &out, const uint8_t *slot_is_root, int *nn_list,
			int *nn_count, cudaStream_t stream, int solver_sms, bool green)
	{
		const int draw_after = e->cfg.draw_after > 0 ? e->cfg.draw_after : e->cells;
		int sms = 148;
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->cfg.device);
		static const bool force_dense = getenv("AGB_SOLVER_DENSE") != nullptr; // tests: exercise the low-register build with few games
		static const int resident_env = getenv("AGB_SOLVER_RESIDENT") != nullptr ? atoi(getenv("AGB_SOLVER_RESIDENT")) : 0; // warps per SM (tuning)
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
		return AGB_OK;
	}
}
