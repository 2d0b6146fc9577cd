// K1 (set board), K2 (add / undo move), K3 (encode NN features), plus augment and outcome kernels.
//
// One warp owns one board slot. Board cells, line words, pattern types and threats live in a structure-of-arrays
// store in HBM (BoardStore, agb_common.cuh); consecutive lanes touch consecutive cells of the same slot, so every
// global access of a warp is one contiguous run. The 1 MiB pattern table and the 4 KiB threat table are read through
// the read-only path and stay resident in L2/L1.
//
// Reference semantics: PatternCalculator::setBoard / addMove / undoMove (src/patterns/PatternCalculator.cpp:40-106,
// 245-366), NNInputFeatures::encode / augment (src/networks/NNInputFeatures.cpp:59-154), getOutcome (src/game/rules.cpp:110-133).
#include "engine.hpp"
#include "patterns_logic.cuh"

namespace agb
{
	using namespace plogic;
	namespace
	{
		constexpr int kWarpsPerBlock = 4;
		constexpr unsigned kFull = 0xFFFFFFFFu;

		struct WarpScratch
		{
				uint64_t lines[kLinePitch];
				int8_t board[kCellPitch];
				int hist_base[2][kHistTypes];
		};

		__device__ inline uint16_t location_code(int r, int c)
		{ // Location::toShort, Move.hpp:66-69
			return static_cast<uint16_t>((c << 8) | r);
		}

		// features + forbidden flag of one cell from its pattern types and threat (board in shared memory)
		__device__ inline uint32_t finish_cell(const int8_t *sboard, int S, int rules, int stm, int r, int c, uint32_t p, uint8_t t,
				const Tables &tables, uint8_t &forbidden, uint32_t *status)
		{
			const int idx = r * S + c;
			uint32_t f = encode_cell(sboard[idx], p, stm);
			forbidden = 0;
			if (rules == RULE_RENJU and sboard[idx] == NONE)
			{ // PatternCalculator::isForbidden, PatternCalculator.hpp:173-189
				const int tc = t & 15;
				if (tc == TT_OVERLINE or tc == TT_FORK_4x4)
					forbidden = 1;
				else if (tc == TT_FORK_3x3)
				{
					Overlay ov;
					forbidden = is_forbidden_raw(sboard, S, r, c, tables, ov);
					if (ov.overflow)
						atomicOr(status, 1u);
				}
				if (forbidden and stm == CROSS)
					f |= 1u << 6;
			}
			return f;
		}

		// ---- K1 + K3 ---------------------------------------------------------------------------------------------
		__global__ void __launch_bounds__(kWarpsPerBlock * 32) set_boards_kernel(BoardStore store, Tables tables, const int8_t *__restrict__ boards,
				const int8_t *__restrict__ sign_to_move, int n, int S, int rules, uint32_t *__restrict__ features, uint32_t *status, const int *__restrict__ n_dev)
		{
			__shared__ WarpScratch scratch[kWarpsPerBlock];
			if (n_dev != nullptr)
				n = *n_dev; // batch size produced on the device (lockstep engine)
			const int lane = threadIdx.x & 31;
			const int warp = threadIdx.x >> 5;
			WarpScratch &ws = scratch[warp];
			const int cells = S * S;
			const int nlines = line_count(S);

			for (int b = blockIdx.x * kWarpsPerBlock + warp; b < n; b += gridDim.x * kWarpsPerBlock)
			{
				const int stm = sign_to_move[b];
				const size_t cell_base = static_cast<size_t>(b) * kCellPitch;
				for (int i = lane; i < cells; i += 32)
				{
					int8_t v = boards[static_cast<size_t>(b) * cells + i];
					if (v < NONE or v > CIRCLE)
					{ // not a Sign a board can hold: reported (kStatusBadInput), treated as empty so that nothing is indexed out of range
						atomicOr(status, kStatusBadInput);
						v = NONE;
					}
					ws.board[i] = v;
					store.board[cell_base + i] = v;
				}
				if (lane == 0 and stm != CROSS and stm != CIRCLE)
					atomicOr(status, kStatusBadInput);
				if (lane < 2 * kHistTypes)
					ws.hist_base[lane / kHistTypes][lane % kHistTypes] = 0;
				if (lane == 0)
					store.sign_to_move[b] = static_cast<int8_t>(stm);
				__syncwarp();
				for (int l = lane; l < nlines; l += 32)
				{
					const uint64_t w = build_line(ws.board, S, l);
					ws.lines[l] = w;
					store.lines[static_cast<size_t>(b) * kLinePitch + l] = w;
				}
				__syncwarp();
				for (int i0 = 0; i0 < cells; i0 += 32)
				{
					const int i = i0 + lane;
					const bool active = i < cells;
					uint32_t p = 0;
					uint8_t t = 0;
					int r = 0, c = 0;
					if (active)
					{
						r = i / S;
						c = i - r * S;
						if (ws.board[i] == NONE)
						{
							p = classify_cell(ws.lines, tables.pattern, r, c, S);
							t = threat_of_cell(p, tables.threat);
						}
						uint8_t forbidden;
						const uint32_t f = finish_cell(ws.board, S, rules, stm, r, c, p, t, tables, forbidden, status);
						store.ptypes[cell_base + i] = p;
						store.threats[cell_base + i] = t;
						store.forbidden[cell_base + i] = forbidden;
						features[static_cast<size_t>(b) * cells + i] = f;
					}
					// threat lists in row-major insertion order (prepare_threat_lists, PatternCalculator.cpp:261-277)
#pragma unroll
					for (int colour = 0; colour < 2; colour++)
					{
						const int type = active ? ((t >> (4 * colour)) & 15) : 0;
						const unsigned same = __match_any_sync(kFull, type);
						if (type != TT_NONE)
						{
							const int pos = ws.hist_base[colour][type] + __popc(same & ((1u << lane) - 1u));
							store.hist_cells[((static_cast<size_t>(b) * 2 + colour) * kHistTypes + type) * kCellPitch + pos] = location_code(r, c);
						}
						__syncwarp();
						if (type != TT_NONE and (same & ((1u << lane) - 1u)) == 0)
							ws.hist_base[colour][type] += __popc(same);
						__syncwarp();
					}
				}
				if (lane < 2 * kHistTypes)
					store.hist_count[static_cast<size_t>(b) * 2 * kHistTypes + lane] = ws.hist_base[lane / kHistTypes][lane % kHistTypes];
				__syncwarp();
			}
		}

		// ---- K3 alone: encode from the persistent state ------------------------------------------------------------
		__global__ void __launch_bounds__(kWarpsPerBlock * 32) encode_kernel(BoardStore store, Tables tables, int n, int S, int rules,
				uint32_t *__restrict__ features, uint32_t *status)
		{
			__shared__ int8_t sboards[kWarpsPerBlock][kCellPitch];
			const int lane = threadIdx.x & 31;
			const int warp = threadIdx.x >> 5;
			int8_t *sboard = sboards[warp];
			const int cells = S * S;
			for (int b = blockIdx.x * kWarpsPerBlock + warp; b < n; b += gridDim.x * kWarpsPerBlock)
			{
				const size_t cell_base = static_cast<size_t>(b) * kCellPitch;
				const int stm = store.sign_to_move[b];
				for (int i = lane; i < cells; i += 32)
					sboard[i] = store.board[cell_base + i];
				__syncwarp();
				for (int i = lane; i < cells; i += 32)
				{
					const int r = i / S, c = i - r * S;
					uint8_t forbidden;
					features[static_cast<size_t>(b) * cells + i] = finish_cell(sboard, S, rules, stm, r, c, store.ptypes[cell_base + i],
							store.threats[cell_base + i], tables, forbidden, status);
					store.forbidden[cell_base + i] = forbidden;
				}
				__syncwarp();
			}
		}

		// ---- K2 ------------------------------------------------------------------------------------------------------
		// ThreatHistogram::remove (find, overwrite with the last element, shrink) and ::add (append), ThreatHistogram.hpp:39-113
		__device__ inline void hist_remove(BoardStore &store, int b, int colour, int type, uint16_t loc, int lane)
		{
			if (type == TT_NONE)
				return;
			int *count = store.hist_count + (static_cast<size_t>(b) * 2 + colour) * kHistTypes + type;
			uint16_t *list = store.hist_cells + ((static_cast<size_t>(b) * 2 + colour) * kHistTypes + type) * kCellPitch;
			const int len = *count;
			int found = -1;
			for (int base = 0; base < len and found < 0; base += 32)
			{
				const int i = base + lane;
				const unsigned hit = __ballot_sync(kFull, i < len and list[i] == loc);
				if (hit)
					found = base + __ffs(hit) - 1;
			}
			__syncwarp();
			if (found >= 0 and lane == 0)
			{
				list[found] = list[len - 1];
				*count = len - 1;
			}
			__syncwarp();
		}
		__device__ inline void hist_add(BoardStore &store, int b, int colour, int type, uint16_t loc, int lane)
		{
			if (type == TT_NONE)
				return;
			if (lane == 0)
			{
				int *count = store.hist_count + (static_cast<size_t>(b) * 2 + colour) * kHistTypes + type;
				uint16_t *list = store.hist_cells + ((static_cast<size_t>(b) * 2 + colour) * kHistTypes + type) * kCellPitch;
				list[*count] = loc;
				*count += 1;
			}
			__syncwarp();
		}

		__global__ void __launch_bounds__(kWarpsPerBlock * 32) add_undo_kernel(BoardStore store, Tables tables, const uint16_t *__restrict__ moves, int n,
				int S, int undo, uint32_t *status)
		{
			const int lane = threadIdx.x & 31;
			const int warp = threadIdx.x >> 5;
			for (int b = blockIdx.x * kWarpsPerBlock + warp; b < n; b += gridDim.x * kWarpsPerBlock)
			{
				const uint32_t mv = moves[b];
				const int sign = mv & 3;
				if (sign == NONE)
					continue;
				const int r = (mv >> 2) & 127, c = (mv >> 9) & 127;
				const size_t cell_base = static_cast<size_t>(b) * kCellPitch;
				const int centre = r * S + c;
				// the reference asserts on these (PatternCalculator.cpp:70, 89); here the slot is left untouched and the call fails
				if (sign == ILLEGAL or r >= S or c >= S or store.board[cell_base + (r < S and c < S ? centre : 0)] != (undo ? sign : NONE))
				{
					if (lane == 0)
						atomicOr(status, kStatusBadInput);
					continue;
				}

				// the four lines through the move: lanes 0..3 own one each (RawPatternCalculator::addMove / undoMove)
				uint64_t line = 0;
				int pos = 0;
				if (lane < 4)
				{
					const size_t li = static_cast<size_t>(b) * kLinePitch + line_index(lane, r, c, S);
					pos = pos_in_line(lane, r, c, S);
					line = store.lines[li];
					if (undo)
						line &= ~(3ull << (12 + 2 * pos));
					else
						line |= static_cast<uint64_t>(sign) << (12 + 2 * pos);
					store.lines[li] = line;
				}
				if (lane == 0)
				{
					store.board[cell_base + centre] = undo ? NONE : sign;
					store.sign_to_move[b] = (store.sign_to_move[b] == CROSS) ? CIRCLE : CROSS;
				}
				__syncwarp();

				// centre cell (update_around, PatternCalculator.cpp:285-318)
				const uint16_t centre_loc = location_code(r, c);
				if (not undo)
				{
					const uint8_t old_t = store.threats[cell_base + centre];
					hist_remove(store, b, 0, old_t & 15, centre_loc, lane);
					hist_remove(store, b, 1, old_t >> 4, centre_loc, lane);
					if (lane == 0)
					{
						store.ptypes[cell_base + centre] = 0;
						store.threats[cell_base + centre] = 0;
					}
				}
				else
				{
					uint32_t byte = 0;
					if (lane < 4)
						byte = static_cast<uint32_t>(tables.pattern[narrow_window(static_cast<uint32_t>(line >> (2 * pos + 2)) & 0x3FFFFFu)]) << (8 * lane);
					byte |= __shfl_xor_sync(kFull, byte, 1);
					byte |= __shfl_xor_sync(kFull, byte, 2);
					const uint32_t p = __shfl_sync(kFull, byte, 0);
					const uint8_t new_t = threat_of_cell(p, tables.threat);
					if (lane == 0)
					{
						store.ptypes[cell_base + centre] = p;
						store.threats[cell_base + centre] = new_t;
					}
					hist_add(store, b, 0, new_t & 15, centre_loc, lane);
					hist_add(store, b, 1, new_t >> 4, centre_loc, lane);
				}

				// the 40 neighbours at distance 1..5 in the four directions: entry q = 4 * k + dir, k-th offset of -5..-1,1..5
				// (the order update_around visits them, PatternCalculator.cpp:320-330). Only empty in-board cells can change.
				uint8_t old_t[2] = { 0, 0 }, new_t[2] = { 0, 0 };
				uint16_t loc[2] = { 0, 0 };
#pragma unroll
				for (int half = 0; half < 2; half++)
				{
					const int q = lane + 32 * half;
					const int dir = q & 3;
					const int k = q >> 2;
					const int off = (k < 5) ? (k - 5) : (k - 4);
					const uint64_t dline = __shfl_sync(kFull, line, dir);
					const int dpos = __shfl_sync(kFull, pos, dir);
					if (q < 40)
					{
						const int nr = r + off * dir_row_step(dir), nc = c + off * dir_col_step(dir);
						if (nr >= 0 and nr < S and nc >= 0 and nc < S and store.board[cell_base + nr * S + nc] == NONE)
						{
							const int ncell = nr * S + nc;
							const uint32_t window = static_cast<uint32_t>(dline >> (2 * (dpos + off) + 2)) & 0x3FFFFFu;
							const uint32_t byte = tables.pattern[narrow_window(window)];
							const uint32_t p = (store.ptypes[cell_base + ncell] & ~(0xFFu << (8 * dir))) | (byte << (8 * dir));
							old_t[half] = store.threats[cell_base + ncell];
							new_t[half] = threat_of_cell(p, tables.threat);
							loc[half] = location_code(nr, nc);
							store.ptypes[cell_base + ncell] = p;
							store.threats[cell_base + ncell] = new_t[half];
						}
					}
				}
				__syncwarp();
				// threat lists: replay the changes in visiting order (update_feature_types_and_threats, PatternCalculator.cpp:345-363)
				unsigned changed0 = __ballot_sync(kFull, old_t[0] != new_t[0]);
				unsigned changed1 = __ballot_sync(kFull, old_t[1] != new_t[1]) & 0xFFu;
				for (int half = 0; half < 2; half++)
				{
					unsigned changed = half ? changed1 : changed0;
					while (changed)
					{
						const int src = __ffs(changed) - 1;
						changed &= changed - 1;
						const int o = __shfl_sync(kFull, static_cast<int>(old_t[half]), src);
						const int nw = __shfl_sync(kFull, static_cast<int>(new_t[half]), src);
						const uint16_t l = static_cast<uint16_t>(__shfl_sync(kFull, static_cast<int>(loc[half]), src));
						if ((o & 15) != (nw & 15))
						{
							hist_remove(store, b, 0, o & 15, l, lane);
							hist_add(store, b, 0, nw & 15, l, lane);
						}
						if ((o >> 4) != (nw >> 4))
						{
							hist_remove(store, b, 1, o >> 4, l, lane);
							hist_add(store, b, 1, nw >> 4, l, lane);
						}
					}
				}
				__syncwarp();
			}
		}

		// ---- augment: board symmetry + direction-bit permutation, one thread per output cell -----------------------
		__global__ void augment_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, const int8_t *__restrict__ symmetry, int n, int S)
		{
			const int cells = S * S;
			const long long gid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
			if (gid >= static_cast<long long>(n) * cells)
				return;
			const int b = static_cast<int>(gid / cells);
			const int i = static_cast<int>(gid - static_cast<long long>(b) * cells);
			const int r = i / S, c = i - r * S;
			const int mode = symmetry[b];
			int sr, sc;
			symmetry_source(mode, S, r, c, sr, sc);
			dst[gid] = permute_direction_bits(src[static_cast<size_t>(b) * cells + sr * S + sc], mode);
		}

		// apply_symmetry on float planes with `channels` values per cell (policy: 1, action values: 3); inverse = undo `mode`
		__global__ void symmetry_f32_kernel(const float *__restrict__ src, float *__restrict__ dst, const int8_t *__restrict__ symmetry, int n, int S,
				int channels, int inverse)
		{
			const int cells = S * S;
			const long long gid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
			if (gid >= static_cast<long long>(n) * cells)
				return;
			const int b = static_cast<int>(gid / cells);
			const int i = static_cast<int>(gid - static_cast<long long>(b) * cells);
			const int r = i / S, c = i - r * S;
			int mode = symmetry[b];
			if (inverse)
				mode = inverse_symmetry(mode);
			int sr, sc;
			symmetry_source(mode, S, r, c, sr, sc);
			for (int k = 0; k < channels; k++)
				dst[gid * channels + k] = src[(static_cast<size_t>(b) * cells + sr * S + sc) * channels + k];
		}

		// ---- getOutcome, one warp per (board, last move) ----------------------------------------------------------------
		__global__ void __launch_bounds__(kWarpsPerBlock * 32) outcome_kernel(Tables tables, const int8_t *__restrict__ boards,
				const uint16_t *__restrict__ moves, int n, int S, int rules, int draw_after, int8_t *__restrict__ outcomes, uint32_t *status)
		{
			__shared__ int8_t sboards[kWarpsPerBlock][kCellPitch];
			const int lane = threadIdx.x & 31;
			const int warp = threadIdx.x >> 5;
			int8_t *sboard = sboards[warp];
			const int cells = S * S;
			for (int b = blockIdx.x * kWarpsPerBlock + warp; b < n; b += gridDim.x * kWarpsPerBlock)
			{
				for (int i = lane; i < cells; i += 32)
					sboard[i] = boards[static_cast<size_t>(b) * cells + i];
				__syncwarp();
				if (lane == 0)
				{
					const uint32_t mv = moves[b];
					bool overflow = false;
					outcomes[b] = static_cast<int8_t>(outcome_of(sboard, S, rules, draw_after, (mv >> 2) & 127, (mv >> 9) & 127, mv & 3, tables, overflow));
					if (overflow)
						atomicOr(status, 1u);
				}
				__syncwarp();
			}
		}

		int grid_for(int n)
		{ // enough blocks for every slot, capped at a few waves of the 148 SMs (the kernels stride over slots)
			const int blocks = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
			const int cap = 148 * 16;
			return blocks < cap ? (blocks > 0 ? blocks : 1) : cap;
		}
	}

	int launch_set_boards(AgbEngine *e, const int8_t *boards_dev, const int8_t *stm_dev, int n, uint32_t *features_dev)
	{
		set_boards_kernel<<<grid_for(n), kWarpsPerBlock * 32, 0, e->stream>>>(e->store, e->tables, boards_dev, stm_dev, n, e->cfg.rows, e->cfg.rules,
				features_dev, e->d_status, nullptr);
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		return AGB_OK;
	}
	BoardStore store_at(const BoardStore &s, int slot)
	{ // the store seen from `slot` on: slot i of the result is slot (slot + i) of `s`
		BoardStore r = s;
		const size_t b = static_cast<size_t>(slot);
		r.capacity = s.capacity - slot;
		r.board += b * kCellPitch;
		r.sign_to_move += b;
		r.lines += b * kLinePitch;
		r.ptypes += b * kCellPitch;
		r.threats += b * kCellPitch;
		r.forbidden += b * kCellPitch;
		r.hist_count += b * 2 * kHistTypes;
		r.hist_cells += b * 2 * kHistTypes * kCellPitch;
		return r;
	}
	int launch_set_boards_counted(AgbEngine *e, const int8_t *boards_dev, const int8_t *stm_dev, const int *n_dev, int max_n, uint32_t *features_dev,
			int slot_base, cudaStream_t stream)
	{ // slots [slot_base, slot_base + *n_dev) of the store, the board / sign / feature arrays being indexed by slot as well
		if (stream == nullptr)
			stream = e->stream;
		const size_t off = static_cast<size_t>(slot_base);
		set_boards_kernel<<<grid_for(max_n), kWarpsPerBlock * 32, 0, stream>>>(store_at(e->store, slot_base), e->tables, boards_dev + off * e->cells, stm_dev + off,
				max_n, e->cfg.rows, e->cfg.rules, features_dev + off * e->cells, e->d_status, n_dev);
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		return AGB_OK;
	}
	int launch_add_undo(AgbEngine *e, const uint16_t *moves_dev, int n, bool undo)
	{
		add_undo_kernel<<<grid_for(n), kWarpsPerBlock * 32, 0, e->stream>>>(e->store, e->tables, moves_dev, n, e->cfg.rows, undo ? 1 : 0, e->d_status);
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		return AGB_OK;
	}
	int launch_encode(AgbEngine *e, int n, uint32_t *features_dev)
	{
		encode_kernel<<<grid_for(n), kWarpsPerBlock * 32, 0, e->stream>>>(e->store, e->tables, n, e->cfg.rows, e->cfg.rules, features_dev, e->d_status);
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		return AGB_OK;
	}
	int launch_augment(AgbEngine *e, const uint32_t *src_dev, uint32_t *dst_dev, const int8_t *sym_dev, int n, cudaStream_t stream)
	{
		if (stream == nullptr)
			stream = e->stream;
		const long long total = static_cast<long long>(n) * e->cells;
		augment_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(src_dev, dst_dev, sym_dev, n, e->cfg.rows);
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		return AGB_OK;
	}
	int launch_symmetry_f32(AgbEngine *e, const float *src_dev, float *dst_dev, const int8_t *sym_dev, int n, int channels, bool inverse, cudaStream_t stream)
	{
		if (stream == nullptr)
			stream = e->stream;
		const long long total = static_cast<long long>(n) * e->cells;
		symmetry_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(src_dev, dst_dev, sym_dev, n, e->cfg.rows, channels, inverse ? 1 : 0);
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		return AGB_OK;
	}
	int launch_outcomes(AgbEngine *e, const int8_t *boards_dev, const uint16_t *moves_dev, int n, int8_t *out_dev)
	{
		const int draw_after = e->cfg.draw_after > 0 ? e->cfg.draw_after : e->cells;
		outcome_kernel<<<grid_for(n), kWarpsPerBlock * 32, 0, e->stream>>>(e->tables, boards_dev, moves_dev, n, e->cfg.rows, e->cfg.rules, draw_after, out_dev,
				e->d_status);
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		return AGB_OK;
	}
}
