// K5, static layer: one thread per leaf position runs the reference's staged move generator + static evaluation on the K1
// state of its slot (solver_logic.cuh), i.e. AlphaBetaSearch::solve with TSSConfig::max_positions = 1
// (src/search/alpha_beta/AlphaBetaSearch.cpp:77-156). Positions that are not proven (or are tree roots) are appended to the
// network batch (Search::scheduleToNN, src/search/monte_carlo/Search.cpp:184-198).
#include "engine.hpp"
#include "solver_logic.cuh"

#include <vector>

namespace agb
{
	namespace
	{
		__global__ void __launch_bounds__(128) solve_static_kernel(BoardStore store, Tables tables, const uint16_t *__restrict__ def_table, const int *__restrict__ n_dev,
				int S, int rules, int draw_after, SolverOutputs out, const uint8_t *__restrict__ slot_is_root, int *__restrict__ nn_list, int *__restrict__ nn_count)
		{
			const int slot = blockIdx.x * blockDim.x + threadIdx.x;
			if (slot >= *n_dev)
				return;
			const int cells = S * S;
			const size_t cbase = static_cast<size_t>(slot) * kCellPitch;
			int stones = 0;
			for (int i = 0; i < cells; i++)
				stones += (store.board[cbase + i] != NONE);
			solver::View v { S, cells, rules, store.sign_to_move[slot], stones, draw_after, kCellPitch, store.board + cbase,
					store.lines + static_cast<size_t>(slot) * kLinePitch, store.ptypes + cbase, store.threats + cbase, store.forbidden + cbase,
					store.hist_count + static_cast<size_t>(slot) * 2 * kHistTypes, store.hist_cells + static_cast<size_t>(slot) * 2 * kHistTypes * kCellPitch,
					tables.pattern, def_table };
			const solver::Result res = solver::solve_static(v, out.moves + static_cast<size_t>(slot) * out.pitch, out.scores + static_cast<size_t>(slot) * out.pitch);
			out.n_actions[slot] = res.n_actions;
			out.score[slot] = res.score;
			out.must_defend[slot] = res.must_defend ? 1 : 0;
			if (slot_is_root[slot] or not solver::sc_is_proven(res.score))
				nn_list[atomicAdd(nn_count, 1)] = slot;
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
	int launch_solve_static(AgbEngine *e, const int *n_dev, int max_n, const SolverOutputs &out, const uint8_t *slot_is_root, int *nn_list, int *nn_count)
	{
		const int draw_after = e->cfg.draw_after > 0 ? e->cfg.draw_after : e->cells;
		solve_static_kernel<<<(max_n + 127) / 128, 128, 0, e->stream>>>(e->store, e->tables, e->d_def_table, n_dev, e->cfg.rows, e->cfg.rules, draw_after, out,
				slot_is_root, nn_list, nn_count);
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		return AGB_OK;
	}
}
