// K4: the ResNet policy/value(/Q) forward as ONE persistent tcgen05 kernel + a small dense kernel for the value head.
//
// Replaces AGNetwork::asyncForwardLaunch -> ml::Graph::predict (src/networks/AGNetwork.cpp:61-68) for the graphs built by
// ResnetPV / ResnetPVQ (src/networks/networks.cpp:71-93, 143-168; blocks.cpp:32-127), BN folded (AGNetwork.cpp:136-149).
//
// Design (B200-first):
//  * one CTA owns one board for the WHOLE network: activations never leave shared memory between layers. Two bf16 board
//    images (x, h) of [C/8][rows][8] live in smem with a zero halo and two zero columns per row (row pitch S+2), so every
//    conv tap is the same image read at a shifted start address -- implicit GEMM without im2col and without border code.
//  * per layer: D[256 positions x F] (fp32, TMEM) = sum over taps of A_tap[256 x C] * W_tap[F x C]^T with tcgen05.mma
//    (M=128 x2 tiles, N=F, K=16), operands K-major / no swizzle (umma.cuh). One elected thread issues.
//  * weights stream from L2 through a ring of stages filled by cp.async.bulk (mbarrier complete_tx); one stage = the nine taps of one
//    16-channel slice of a layer's input, which is the unit the K loop consumes (it runs input-slice major).
//  * epilogue warps read TMEM (tcgen05.ld) 16 columns at a time, add bias, ReLU (+ residual, in place), write that slice of the next
//    layer's image and publish it on its own mbarrier: the next layer's MMAs on the slice start at once, into the other of two TMEM
//    accumulator sets, while the rest is still being drained. The policy 1x1 + softmax, the value 1x1 and the Q 1x1 + softmax are fused
//    into the epilogues of their convs.
//  * the CTA pairs draw their boards from a global ticket counter (a pair whose SMs free up late takes fewer).
//  * the kernel's code size is kept small on purpose (slice loops are loops, head code sits outside the trunk's loop): ten warps per SM
//    fetch it, and the solver kernel running beside it fetches its own code through the same L2.
//  * input: the 32-bit feature words are unpacked into the stem's bf16 image by the same warps (replaces ml::unpackInput).
#include "engine.hpp"
#include "umma.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace agb
{
	using namespace umma;

	enum : int { MODE_STEM = 0, MODE_CONV1 = 1, MODE_CONV2 = 2, MODE_POLICY = 3, MODE_QHEAD = 4 };
	constexpr int kMaxConvLayers = 48;
	constexpr int kMaxStages = 8;
	constexpr int kThreads = 320; // warp 0: weight producer, warp 1: MMA issuer, warps 2..9: epilogue
	constexpr int kEpilogueThreads = 256;
	constexpr int kTicketCounters = 32;
#ifndef AGB_NET_SPLIT_SLICES
#define AGB_NET_SPLIT_SLICES 4 // boards split over the CTA pair: 16-channel slices per hand-over (measured at 20x20: 1: 21.6, 2: 20.2, 4: see profiles, 8: 21.7 ms per 4096)
#endif

	struct ConvDesc
	{
			uint32_t w_offset; // bytes into the bf16 weight images
			uint32_t bias_offset; // floats into the bias array
			uint16_t cin_chunks; // C_in / 8
			uint8_t radius; // 1 (3x3) or 2 (5x5)
			uint8_t mode;
			uint8_t n_taps; // taps in the weight stream: k * k, rounded up to a multiple of nine with zero weights (one ring stage = nine taps)
			uint8_t real_taps; // k * k
			uint8_t last_trunk; // value 1x1 conv runs after this layer's epilogue
			uint8_t pad;
	};
	struct NetParams
	{
			ConvDesc layers[kMaxConvLayers];
			int n_layers;
			int S, P, F; // board size, row pitch (S + 2), filters
			int n_stages, stage_bytes; // weight ring: stage_bytes = kc_per_stage * (F / 2) * 16 (each CTA of a pair holds half of C_out)
			int kc_per_stage; // 8-channel weight slices ((F / 2) rows x 16 bytes) per stage, even
			int buf_bytes; // one trunk image: (F / 8) * img_rows * 16
			int img_rows, in_img_rows;
			const uint8_t *w_images;
			const float *bias;
			const float *policy_w1; // [F], then b1
			const float *value_w; // [4][F], then b[4]
			const float *q_w1; // [3][F], then b[3]
			const uint32_t *features; // [n][S*S]
			float *policy; // [n][S*S]
			float *value_hidden; // [n][S*S*4]
			float *q; // [n][S*S][3] or null
			int n_boards;
			const int *n_boards_dev; // if set, the batch size is read from device memory
			const int *gather; // if set, board i of this launch lives in slot gather[i] of the feature / output arrays
			int slot_base; // otherwise board i lives in slot slot_base + i
			int *ticket; // zeroed before the launch: the CTA pairs draw their boards from it (see draw_ticket in the kernel)
			long long *trace; // optional [n_layers][4] clock64 stamps of CTA 0's first board (AGB_NET_TRACE)
	};

	struct NetWeights
	{
			NetParams params { };
			uint8_t *d_w_images = nullptr;
			float *d_small = nullptr; // biases + head vectors
			float *d_wd1 = nullptr, *d_bd1 = nullptr, *d_wd2 = nullptr, *d_bd2 = nullptr;
			float *d_value_hidden = nullptr;
			float *d_policy = nullptr, *d_value = nullptr, *d_q = nullptr; // staging for host entry point
			// one board-ticket counter per stream K4 is launched on (launches on one stream are ordered, so its counter is free again when the next
			// launch's memset runs; launches on different streams may overlap and must not share one)
			int *d_tickets = nullptr; // [kTicketCounters]
			cudaStream_t ticket_streams[kTicketCounters] = { };
			int n_ticket_streams = 0;
			int dense_width = 0;
			size_t smem_bytes = 0;
			bool split = false; // one board per CTA pair (boards of more than 15 rows)
			bool loaded = false;
	};

	namespace
	{
		__device__ __forceinline__ void named_barrier_epilogue()
		{
			asm volatile("bar.sync 1, 256;" ::: "memory");
		}
		__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi)
		{ // cvt.rn.relu.bf16x2.f32: max(x, 0) folded into the rounding conversion
			uint32_t r;
			asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
			return r;
		}
		__device__ __forceinline__ float tanh_fast(float x)
		{ // tanh.approx.f32 (one MUFU instruction, relative error 2^-11): the action-value head's activation feeds a 3-way softmax stored as
		  // fp32 next to bf16 activations; tanhf's exact expansion is 30 instructions x 16 channels in the middle of the epilogue's slice loop
			float y;
			asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
			return y;
		}
		__device__ __forceinline__ void unpack_bf16(uint32_t w, float &lo, float &hi)
		{
			lo = __uint_as_float(w << 16);
			hi = __uint_as_float(w & 0xFFFF0000u);
		}

		// Straight-line MMA schedule of one 3x3, F -> F convolution with row pitch P (17: 15x15 board, 22: 20x20 board split over the CTA
		// pair) (the trunk and head convs: 40 of the 42 layers of a 20-block net). The issuing lane cannot hide latency, so everything that
		// can be a compile-time constant is one: tap offsets, K-slice offsets and the weight-ring geometry (one stage = the nine taps of one
		// 16-channel slice of the input). Per MMA pair the lane executes two 64-bit adds and the two tcgen05.mma instructions.
		// The K loop runs INPUT-CHANNEL-SLICE major: slice c of the input image is exactly what the previous layer's epilogue produces from
		// accumulator columns 16c..16c+15, so this layer's MMAs on slice c start as soon as that part of the epilogue is done
		// (chunk_ready[c]) while the rest of the epilogue is still draining the other accumulator set.
		template<int F, int P, int KS>
		__device__ __forceinline__ void issue_conv(uint32_t tmem_acc, uint64_t a_desc0, uint64_t b_desc0, uint32_t idesc, uint32_t a_kc_step,
				uint32_t stage_step, uint64_t *w_full, uint64_t *peer_full, uint64_t *w_empty, uint64_t *chunk_ready, uint32_t &ready_phase, int &stage,
				uint32_t &phase, int n_stages, int n_slices, long long *trace)
		{
			constexpr int kTaps = KS * KS, kStagesPerSlice = (kTaps + 8) / 9; // 3x3: one stage per 16-channel slice; the 5x5 stem: 25 taps + 2 of zero weights = 3
			// the slice loop is NOT unrolled (the taps inside are): a few loop instructions per 18 MMAs (1 152 tensor-pipe cycles) cost nothing,
			// 8 x 100 instructions of straight-line code did -- K4's instruction footprint is felt by its own epilogue warps and, through the L2, by K5
#pragma unroll 1
			for (int c = 0; c < n_slices; c++)
			{
#pragma unroll
				for (int s = 0; s < kStagesPerSlice; s++)
				{
					// the weights first: they arrived long ago, and at the start of a layer (MMA queue empty) every wait after the image slice is
					// ready would add to the bubble
					mbar_wait(&w_full[stage], phase); // our half of the weights
					mbar_wait_cluster(&peer_full[stage], phase); // the peer's half
					if (s == 0)
					{
						mbar_wait_cluster(&chunk_ready[c], (ready_phase >> c) & 1u); // input channels 16c..16c+15 written by both CTAs, their accumulator columns drained
						ready_phase ^= 1u << c;
					}
					tc_fence_after();
					if (s == 0 and c == 0 and trace != nullptr)
						*trace = clock64();
					const uint64_t b_stage = b_desc0 + static_cast<uint32_t>(stage) * stage_step;
					if (elect_one())
					{
#pragma unroll
						for (int tt = 0; tt < 9; tt++)
						{
							constexpr int kDummy = 0;
							const int tap = s * 9 + tt;
							const int row0 = (tap < kTaps) ? (tap / KS) * P + (tap % KS) : kDummy; // taps past the kernel carry zero weights: any finite rows do
							const uint64_t ad = a_desc0 + row0 + static_cast<uint32_t>(2 * c) * a_kc_step;
							const uint64_t bd = b_stage + tt * 2 * (F / 2);
							const bool acc = (tap != 0) or (c != 0);
							mma_pair_bf16(tmem_acc, ad, bd, idesc, acc);
							mma_pair_bf16(tmem_acc + F, ad + 128, bd, idesc, acc);
						}
						mma_pair_commit(&w_empty[stage], 3); // both CTAs may refill this stage once these MMAs have read it
					}
					__syncwarp();
					if (++stage == n_stages)
					{
						stage = 0;
						phase ^= 1;
					}
				}
			}
		}

		// SPLIT = false: one board per CTA (two per pair), boards up to 15x15. SPLIT = true: one board per PAIR, boards of 16..20 rows: rank 0
		// owns the upper rows, rank 1 the lower ones; after every layer each CTA writes its boundary row into the other's halo row through
		// distributed shared memory, so the tensor-core schedule is the same as for two independent boards.
		template<int F, bool SPLIT>
#ifdef AGB_NET_MAXNREG
		// experiment: a register cap leaves room for solver warps of another pipeline group next to the resident CTA
		__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(AGB_NET_MAXNREG) resnet_board_kernel(const __grid_constant__ NetParams prm)
#else
		__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) resnet_board_kernel(const __grid_constant__ NetParams prm)
#endif
		{
			extern __shared__ __align__(1024) uint8_t smem[];
			const int S = prm.S, P = prm.P;
			const int n_boards = prm.n_boards_dev ? *prm.n_boards_dev : prm.n_boards;
			const int NS = prm.n_stages;
			uint8_t *buf_x = smem;
			uint8_t *buf_h = smem + prm.buf_bytes;
			uint8_t *stages = smem + 2 * prm.buf_bytes;
			float *logits = reinterpret_cast<float*>(stages + NS * prm.stage_bytes); // [256]
			float *reduce = logits + 256; // [8]
			float *sbias2 = reduce + 8; // [2][128]
			float *svalue_w = sbias2 + 256; // [4][F] + [4]: the value head's 1x1 convolution, staged once per kernel
			float *xchg = svalue_w + 4 * 128 + 8; // [2] softmax (max, sum) of the peer's half board (SPLIT)
			uint64_t *bars = reinterpret_cast<uint64_t*>(xchg + 8);
			uint64_t *w_full = bars, *w_empty = bars + 8, *acc_full = bars + 16, *peer_full = bars + 18, *xbar = bars + 26, *chunk_ready = bars + 27;
			uint32_t *tmem_slot = reinterpret_cast<uint32_t*>(bars + 35);
			uint64_t *ticket_ready = bars + 36; // [4]
			int *tickets = reinterpret_cast<int*>(bars + 40); // [4]
			// CTA pair: rank 0 (leader) issues the MMAs for both boards; each CTA loads half of every weight tile
			const uint32_t rank = cluster_ctarank();
			// Boards are handed out dynamically, one unit (two boards; SPLIT: one) per draw from a global counter: a pair whose SMs become free late
			// (the solver's blocks of another pipeline group may sit on them when the launch starts) simply takes fewer units instead of holding the
			// whole launch up with a fixed share. The leader's producer lane draws, writes the ticket into a 4-slot ring in both CTAs and arrives
			// on the slot's barrier; every other role of the pair reads iteration i's ticket from slot i % 4. No role can be two iterations ahead
			// of another (the weight ring holds less than a layer, the MMA warp waits for the epilogue's image), so four slots never wrap.
			const int n_units = SPLIT ? n_boards : (n_boards + 1) / 2;
			const auto draw_ticket = [&](int it) -> int
			{ // leader CTA, warp 0, lane 0
				const int t = atomicAdd(prm.ticket, 1);
				const int slot = it & 3;
				tickets[slot] = t;
				st_peer_u32(map_to_rank(smem_u32(&tickets[slot]), 1), static_cast<uint32_t>(t));
				mbar_arrive(&ticket_ready[slot]);
				mbar_arrive_remote(&ticket_ready[slot], 1);
				return t;
			};
			const auto read_ticket = [&](int it) -> int
			{
				const int slot = it & 3;
				mbar_wait_cluster(&ticket_ready[slot], (it >> 2) & 1);
				return tickets[slot];
			};

			const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
			const uint32_t img_chunk_bytes = prm.img_rows * 16;
			const uint32_t in_chunk_bytes = prm.in_img_rows * 16;
			const uint32_t tmem_cols = 4 * F; // two accumulator sets (layer parity) of two M-tiles x F columns: 256 (F = 64) or all 512 columns

			if (threadIdx.x == 0)
			{
				for (int s = 0; s < NS; s++)
				{
					mbar_init(&w_full[s], 1);
					mbar_init(&w_empty[s], 1);
					mbar_init(&peer_full[s], 1);
				}
				for (int i = 0; i < 4; i++)
					mbar_init(&ticket_ready[i], 1);
				mbar_init(acc_full, 1);
				mbar_init(xbar, 1);
				for (int c = 0; c < 8; c++)
					mbar_init(&chunk_ready[c], 2 * (kEpilogueThreads / 32)); // one arrival per epilogue warp of both CTAs, on the leader's barrier
				fence_mbar_init();
			}
			if (warp == 1)
			{
				tmem_alloc_pair(tmem_slot, tmem_cols);
				tmem_relinquish_pair();
			}
			for (int i = threadIdx.x; i < 4 * F + 4; i += kThreads)
				svalue_w[i] = __ldg(prm.value_w + i);
			// x image: the halo and the two pad columns stay zero for the whole kernel (only valid cells are ever written)
			for (uint32_t i = threadIdx.x; i < static_cast<uint32_t>(prm.buf_bytes) / 16; i += kThreads)
				reinterpret_cast<uint4*>(buf_x)[i] = make_uint4(0, 0, 0, 0);
			tc_fence_before();
			__syncthreads();
			cluster_sync(); // barriers of both CTAs are initialised before any remote arrive / multicast commit
			tc_fence_after();
			const uint32_t tmem_base = *tmem_slot;

			if (warp == 0)
			{ // ===== weight producer: one lane streams every layer's taps through the stage ring =====
				if (lane == 0)
				{
					int stage = 0;
					uint32_t phase = 0;
					for (int it = 0; (rank == 0 ? draw_ticket(it) : read_ticket(it)) < n_units; it++)
						for (int l = 0; l < prm.n_layers; l++)
						{
							const ConvDesc &L = prm.layers[l];
							const int total_kc = L.n_taps * L.cin_chunks; // the layer's weights as a stream of 8-channel slices
							for (int kc0 = 0; kc0 < total_kc; kc0 += prm.kc_per_stage)
							{
								const uint32_t bytes = min(prm.kc_per_stage, total_kc - kc0) * (F / 2) * 16;
								mbar_wait_backoff(&w_empty[stage], phase ^ 1);
								mbar_arrive_expect_tx(&w_full[stage], bytes);
								bulk_g2s(stages + stage * prm.stage_bytes, prm.w_images + L.w_offset + (static_cast<size_t>(rank) * total_kc + kc0) * (F / 2) * 16, bytes,
										&w_full[stage]);
								if (++stage == NS)
								{
									stage = 0;
									phase ^= 1;
								}
							}
						}
				}
			}
			else if (warp == 1)
			{ // ===== MMA issuer (leader CTA) / weight-arrival relay (peer CTA) =====
				if (rank == 0)
				{ // the whole warp walks the (warp-uniform) schedule; one elected lane issues the MMAs and commits
					int stage = 0;
					uint32_t phase = 0, ready_phase = 0;
					const uint32_t idesc = idesc_bf16_f32(256, F); // M = 256: 128 positions of this CTA's board + 128 of the peer's
					for (int it = 0;; it++)
					{
						const int ticket = read_ticket(it);
						if (ticket >= n_units)
							break;
						const int b0 = SPLIT ? ticket : 2 * ticket;
						for (int l = 0; l < prm.n_layers; l++)
						{
							const ConvDesc &L = prm.layers[l];
							const int KW = 2 * L.radius + 1;
							const uint32_t a_lbo = (L.mode == MODE_STEM) ? in_chunk_bytes : img_chunk_bytes;
							const uint32_t a_img = smem_u32((L.mode == MODE_STEM or L.mode == MODE_CONV2) ? buf_h : buf_x);
							const int total_kc = L.n_taps * L.cin_chunks, cin_chunks = L.cin_chunks;
							// descriptors are advanced by adding to their 14-bit start-address field (16-byte units); no field can carry.
							// The same shared-memory offsets are valid in both CTAs of the pair.
							const uint64_t a_desc0 = smem_desc(a_img, a_lbo, 128), b_desc0 = smem_desc(smem_u32(stages), (F / 2) * 16, 128);
							const uint32_t a_kc_step = a_lbo >> 4, b_kc_step = F / 2, stage_step = prm.stage_bytes >> 4;
							const uint32_t tmem_acc = tmem_base + (l & 1) * 2 * F; // accumulator sets alternate: the previous layer's is still being drained
							long long *trace0 = (prm.trace and b0 == 0 and lane == 0) ? prm.trace + 6 * l : nullptr;
							int chunk = 0, tap = 0, ky = 0, kx = 0; // position in the layer's stream of weight slices: 16-channel slice of the input, then tap
							const bool fast = (S == (SPLIT ? 20 : 15) and prm.kc_per_stage == 18 and L.n_taps % 9 == 0);
							if (fast and L.radius == 1)
								issue_conv<F, SPLIT ? 22 : 17, 3>(tmem_acc, a_desc0, b_desc0, idesc, a_kc_step, stage_step, w_full, peer_full, w_empty, chunk_ready, ready_phase,
										stage, phase, NS, cin_chunks / 2, trace0);
							else if (fast and L.radius == 2)
								issue_conv<F, SPLIT ? 22 : 17, 5>(tmem_acc, a_desc0, b_desc0, idesc, a_kc_step, stage_step, w_full, peer_full, w_empty, chunk_ready, ready_phase,
										stage, phase, NS, cin_chunks / 2, trace0);
							else
							for (int kc0 = 0; kc0 < total_kc; kc0 += prm.kc_per_stage)
							{
								const int nkc = min(prm.kc_per_stage, total_kc - kc0);
								mbar_wait(&w_full[stage], phase); // our half of the weights
								mbar_wait_cluster(&peer_full[stage], phase); // the peer's half
								tc_fence_after();
								uint64_t bd = b_desc0 + stage * stage_step;
								for (int j = 0; j < nkc; j += 2)
								{ // one K=16 step: two 8-channel slices of one tap, for both M-tiles of both boards
									if (tap == 0)
									{ // a new 16-channel slice of the input image: wait until both CTAs have written it
										mbar_wait_cluster(&chunk_ready[chunk], (ready_phase >> chunk) & 1u);
										ready_phase ^= 1u << chunk;
										tc_fence_after();
										if (chunk == 0 and trace0 != nullptr)
											*trace0 = clock64();
									}
									const uint64_t ad = a_desc0 + (tap < L.real_taps ? ky * P + kx : 0) + 2 * chunk * a_kc_step; // taps past the kernel carry zero weights
									const bool acc = (kc0 + j) != 0;
									if (elect_one())
									{
										mma_pair_bf16(tmem_acc, ad, bd, idesc, acc);
										mma_pair_bf16(tmem_acc + F, ad + 128, bd, idesc, acc);
									}
									__syncwarp();
									bd += 2 * b_kc_step;
									if (++kx == KW)
									{
										kx = 0;
										ky++;
									}
									if (++tap == L.n_taps)
									{
										tap = 0;
										ky = 0;
										kx = 0; // the stream's taps past the kernel do not end on a row boundary
										chunk++;
									}
								}
								if (elect_one())
									mma_pair_commit(&w_empty[stage], 3); // both CTAs may refill this stage once these MMAs have read it
								__syncwarp();
								if (++stage == NS)
								{
									stage = 0;
									phase ^= 1;
								}
							}
							if (prm.trace and b0 == 0 and lane == 0)
							{
								prm.trace[6 * l + 1] = clock64();
							}
							if (elect_one())
								mma_pair_commit(acc_full, 3);
							__syncwarp();
						}
					}
				}
				else if (lane == 0)
				{ // peer CTA: tell the leader when our half of each weight stage has landed
					int stage = 0;
					uint32_t phase = 0;
					for (int it = 0; read_ticket(it) < n_units; it++)
						for (int l = 0; l < prm.n_layers; l++)
						{
							const int total_kc = prm.layers[l].n_taps * prm.layers[l].cin_chunks;
							for (int kc0 = 0; kc0 < total_kc; kc0 += prm.kc_per_stage)
							{
								mbar_wait_backoff(&w_full[stage], phase);
								mbar_arrive_remote(&peer_full[stage], 0);
								if (++stage == NS)
								{
									stage = 0;
									phase ^= 1;
								}
							}
						}
				}
			}
			else
			{ // ===== epilogue warps (also build the input image and run the fused heads) =====
				const int et = threadIdx.x - 64;
				const int quadrant = warp & 3; // TMEM lanes 32*quadrant .. +31 are the ones this warp may read
				const int tile = (warp - 2) >> 2; // which of the board's two 128-row M-tiles this warp drains (all F channels of its 32 rows)
				const int cells = S * S;
				uint32_t acc_phase = 0, xchg_phase = 0;
				// rows of the board this CTA computes: all of them, or (SPLIT) the upper / lower part
				const int rows0 = SPLIT ? (S + 1) / 2 : S;
				const int my_rows = SPLIT ? (rank == 0 ? rows0 : S - rows0) : S;
				const int row_begin = (SPLIT and rank == 1) ? rows0 : 0;
				const int halo_in = SPLIT ? 2 : 0; // rows of the neighbouring part that the 5x5 stem reads: taken from the features directly
				for (int it = 0;; it++)
				{
					const int ticket = read_ticket(it);
					if (ticket >= n_units)
						break;
					const int b0 = SPLIT ? ticket : 2 * ticket;
					const int bi = SPLIT ? b0 : b0 + rank; // this CTA's board in the launch
					const bool live = bi < n_boards;
					const int b = (live and prm.gather != nullptr) ? prm.gather[bi] : bi + prm.slot_base; // its slot in the feature / output arrays // an odd batch leaves the last peer without a board: it still runs every barrier
					// ---- prologue: feature words -> bf16 stem image (32 channels, halo 2) in the h buffer ----
					for (uint32_t i = et; i < 4 * in_chunk_bytes / 16; i += kEpilogueThreads)
						reinterpret_cast<uint4*>(buf_h)[i] = make_uint4(0, 0, 0, 0);
					named_barrier_epilogue();
					for (int lc = et; lc < (my_rows + 2 * halo_in) * S; lc += kEpilogueThreads)
					{
						const int y = lc / S - halo_in, x = lc - (y + halo_in) * S; // row relative to this CTA's first row
						const int gy = row_begin + y;
						if (gy < 0 or gy >= S)
							continue;
						const uint32_t f = live ? prm.features[static_cast<size_t>(b) * cells + gy * S + x] : 0u;
						const uint32_t idx = (y + 2) * P + x + 2;
#pragma unroll
						for (int k = 0; k < 4; k++)
						{
							const uint32_t bits = (f >> (8 * k)) & 0xFFu;
							uint4 v;
							v.x = ((bits & 1u) ? 0x3F80u : 0u) | ((bits & 2u) ? 0x3F800000u : 0u);
							v.y = ((bits & 4u) ? 0x3F80u : 0u) | ((bits & 8u) ? 0x3F800000u : 0u);
							v.z = ((bits & 16u) ? 0x3F80u : 0u) | ((bits & 32u) ? 0x3F800000u : 0u);
							v.w = ((bits & 64u) ? 0x3F80u : 0u) | ((bits & 128u) ? 0x3F800000u : 0u);
							*reinterpret_cast<uint4*>(buf_h + k * in_chunk_bytes + idx * 16) = v;
						}
					}
					fence_proxy_async();
					tc_fence_before();
					__syncwarp(); // the lane that arrives publishes the whole warp's writes
					if (lane == 0)
					{ // the stem reads 32 input channels = two 16-channel slices
						for (int c = 0; c < 2; c++)
						{
							if (rank == 0)
								mbar_arrive(&chunk_ready[c]);
							else
								mbar_arrive_remote(&chunk_ready[c], 0);
						}
					}
					// this warp's 32 accumulator rows = 32 positions of the padded board image
					const int p = tile * 128 + quadrant * 32 + lane;
					const int y = p / P, x = p - y * P;
					const bool valid = (y < my_rows) and (x < S);
					const uint32_t out_idx = p + P + 1;
					const bool send = SPLIT and valid and (rank == 0 ? (y == my_rows - 1) : (y == 0));
					const uint32_t peer_idx = (rank == 0) ? (x + 1) : ((rows0 + 1) * P + x + 1);

					for (int l = 0; l < prm.n_layers; l++)
					{
						const ConvDesc &L = prm.layers[l];
						// this layer's bias goes to shared memory while its MMAs are still running
						// (double buffered by layer parity: slower warps may still be reading the previous layer's copy)
						float *sbias = sbias2 + (l & 1) * 128;
						if (et < F)
							sbias[et] = __ldg(prm.bias + L.bias_offset + et);
						named_barrier_epilogue();
						mbar_wait(acc_full, acc_phase & 1); // try_wait suspends the warp in hardware: no polling load next to the MMA-issuing lane
						acc_phase++;
						tc_fence_after();
						if (prm.trace and bi == 0 and et == 0)
							prm.trace[6 * l + 2] = clock64();
						if (L.mode == MODE_STEM)
						{ // the stem image is dead now: give the h buffer its zero halo back
							for (uint32_t i = et; i < static_cast<uint32_t>(prm.buf_bytes) / 16; i += kEpilogueThreads)
								reinterpret_cast<uint4*>(buf_h)[i] = make_uint4(0, 0, 0, 0);
						}
						uint8_t *out_img = (L.mode == MODE_CONV1) ? buf_h : buf_x;
						float head[3] = { 0.f, 0.f, 0.f };
						// SPLIT: the row next to the other part also goes into the peer's halo row of the same image (distributed shared memory)
						const uint32_t peer_img = SPLIT ? map_to_rank(smem_u32(out_img), rank ^ 1u) : 0u;
						const bool hand_over = l + 1 < prm.n_layers;
						// The accumulator is drained 16 columns (= 16 output channels = one K slice of the next layer) at a time; after each slice the
						// warp publishes it (chunk_ready), so the next layer's MMAs on that slice run while the rest is still being drained. The TMEM load
						// of slice c+1 is in flight while slice c is processed.
						const uint32_t taddr = tmem_base + ((quadrant * 32u) << 16) + (l & 1) * 2 * F + tile * F;
						uint32_t v[2][16];
						tmem_ld16(taddr, v[0]);
						// SPLIT: publishing also covers the boundary row written into the peer's image, which needs the full proxy fence and a cluster-scope
						// release (1.5 k cycles against 0.1 k for the local ones): several slices per hand-over there, so that the epilogue stays shorter
						// than the next layer's MMAs
						constexpr int kSlicesPerHandOver = SPLIT ? AGB_NET_SPLIT_SLICES : 1;
						const auto publish = [&](int cb)
						{ // hand these slices of the image (and the drained accumulator columns) to the MMA warp
							if (hand_over and (cb + 1) % kSlicesPerHandOver == 0)
							{
								tc_fence_before();
								if constexpr (SPLIT)
									fence_proxy_async_all(); // covers the rows written into the peer's image
								else
									fence_proxy_async();
								__syncwarp(); // the lane that arrives publishes the whole warp's writes
								if (lane == 0)
								{
#pragma unroll
									for (int c = cb + 1 - kSlicesPerHandOver; c <= cb; c++)
									{
										if (rank == 0)
										{
											if constexpr (SPLIT)
												mbar_arrive_cluster(&chunk_ready[c]);
											else
												mbar_arrive(&chunk_ready[c]);
										}
										else
											mbar_arrive_remote(&chunk_ready[c], 0);
									}
								}
							}
						};
						// Two loops, each unrolled by two only: the 40 trunk layers run the first, whose body is what every epilogue warp fetches all the
						// time -- the kernel's code size is felt by K4 itself and, through the L2, by the solver beside it (fully unrolled with the head
						// code inside: 286 k evaluations/s in the step, like this: see profiles)
						if (L.mode != MODE_POLICY and L.mode != MODE_QHEAD)
						{
#pragma unroll 2
							for (int cb = 0; cb < F / 16; cb++)
							{
								const int c0 = cb * 16;
								uint4 res[2];
								if (L.mode == MODE_CONV2)
								{ // residual operand: x is updated in place
									res[0] = *reinterpret_cast<const uint4*>(buf_x + (c0 / 8) * img_chunk_bytes + out_idx * 16);
									res[1] = *reinterpret_cast<const uint4*>(buf_x + (c0 / 8 + 1) * img_chunk_bytes + out_idx * 16);
								}
								tmem_ld_wait();
								if (cb + 1 < F / 16)
									tmem_ld16(taddr + c0 + 16, v[(cb + 1) & 1]);
								float a[16];
#pragma unroll
								for (int j = 0; j < 16; j += 4)
								{
									const float4 bj = *reinterpret_cast<const float4*>(sbias + c0 + j);
									a[j] = __uint_as_float(v[cb & 1][j]) + bj.x;
									a[j + 1] = __uint_as_float(v[cb & 1][j + 1]) + bj.y;
									a[j + 2] = __uint_as_float(v[cb & 1][j + 2]) + bj.z;
									a[j + 3] = __uint_as_float(v[cb & 1][j + 3]) + bj.w;
								}
								if (L.mode == MODE_CONV2)
								{
#pragma unroll
									for (int h8 = 0; h8 < 2; h8++)
									{
										const uint4 r = res[h8];
										float lo, hi;
										unpack_bf16(r.x, lo, hi); a[8 * h8 + 0] += lo; a[8 * h8 + 1] += hi;
										unpack_bf16(r.y, lo, hi); a[8 * h8 + 2] += lo; a[8 * h8 + 3] += hi;
										unpack_bf16(r.z, lo, hi); a[8 * h8 + 4] += lo; a[8 * h8 + 5] += hi;
										unpack_bf16(r.w, lo, hi); a[8 * h8 + 6] += lo; a[8 * h8 + 7] += hi;
									}
								}
								if (valid)
								{
#pragma unroll
									for (int h8 = 0; h8 < 2; h8++)
									{
										uint4 o;
										o.x = pack_bf16_relu(a[8 * h8 + 0], a[8 * h8 + 1]);
										o.y = pack_bf16_relu(a[8 * h8 + 2], a[8 * h8 + 3]);
										o.z = pack_bf16_relu(a[8 * h8 + 4], a[8 * h8 + 5]);
										o.w = pack_bf16_relu(a[8 * h8 + 6], a[8 * h8 + 7]);
										*reinterpret_cast<uint4*>(out_img + (c0 / 8 + h8) * img_chunk_bytes + out_idx * 16) = o;
										if (send)
											st_peer_v4(peer_img + (c0 / 8 + h8) * img_chunk_bytes + peer_idx * 16, o);
									}
								}
								publish(cb);
							}
						}
						else
						{ // the two head convolutions: nothing is written to the image, the 1x1 convolutions behind them are accumulated per position
#pragma unroll 2
							for (int cb = 0; cb < F / 16; cb++)
							{
								const int c0 = cb * 16;
								tmem_ld_wait();
								if (cb + 1 < F / 16)
									tmem_ld16(taddr + c0 + 16, v[(cb + 1) & 1]);
								if (L.mode == MODE_QHEAD)
								{
#pragma unroll
									for (int j = 0; j < 16; j++)
									{
										const float t = tanh_fast(__uint_as_float(v[cb & 1][j]) + sbias[c0 + j]);
										head[0] += t * __ldg(prm.q_w1 + c0 + j);
										head[1] += t * __ldg(prm.q_w1 + F + c0 + j);
										head[2] += t * __ldg(prm.q_w1 + 2 * F + c0 + j);
									}
								}
								else
								{
#pragma unroll
									for (int j = 0; j < 16; j++)
										head[0] += fmaxf(__uint_as_float(v[cb & 1][j]) + sbias[c0 + j], 0.0f) * __ldg(prm.policy_w1 + c0 + j);
								}
								publish(cb);
							}
						}
						tc_fence_before();

						if (L.mode == MODE_POLICY)
						{ // 1x1 conv to one logit per cell, softmax over the board (createPolicyHead, blocks.cpp:99-107)
							logits[p] = valid ? (head[0] + __ldg(prm.policy_w1 + F)) : -INFINITY;
							named_barrier_epilogue();
							float m = logits[et];
							for (int o = 16; o > 0; o >>= 1)
								m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
							if (lane == 0)
								reduce[warp - 2] = m;
							named_barrier_epilogue();
							m = reduce[0];
							for (int w = 1; w < 8; w++)
								m = fmaxf(m, reduce[w]);
							const float e = expf(logits[et] - m);
							float s = e;
							for (int o = 16; o > 0; o >>= 1)
								s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
							named_barrier_epilogue();
							if (lane == 0)
								reduce[warp - 2] = s;
							named_barrier_epilogue();
							s = 0.0f;
							for (int w = 0; w < 8; w++)
								s += reduce[w];
							float out = e / s;
							if constexpr (SPLIT)
							{ // the two halves of the board merge their (max, sum) pairs: one exchange through the peer's shared memory
								if (et == 0)
								{
									const uint32_t peer_xchg = map_to_rank(smem_u32(xchg), rank ^ 1u);
									st_peer_f32(peer_xchg, m);
									st_peer_f32(peer_xchg + 4, s);
									mbar_arrive_remote(xbar, rank ^ 1u);
								}
								mbar_wait_cluster(xbar, xchg_phase & 1);
								xchg_phase++;
								const float m_peer = xchg[0], s_peer = xchg[1];
								const float m_all = fmaxf(m, m_peer);
								const float scale = expf(m - m_all);
								out = e * scale / (s * scale + s_peer * expf(m_peer - m_all));
							}
							{
								const int pe = et, ye = pe / P, xe = pe - ye * P;
								if (live and ye < my_rows and xe < S)
									prm.policy[static_cast<size_t>(b) * cells + (row_begin + ye) * S + xe] = out;
							}
							named_barrier_epilogue();
						}
						else if (L.mode == MODE_QHEAD)
						{ // 1x1 conv to 3 logits per cell, softmax over them (createActionValuesHead, blocks.cpp:119-127)
							if (live and valid and prm.q != nullptr)
							{
								float z[3];
								for (int k = 0; k < 3; k++)
									z[k] = head[k] + __ldg(prm.q_w1 + 3 * F + k);
								const float m = fmaxf(z[0], fmaxf(z[1], z[2]));
								const float e0 = expf(z[0] - m), e1 = expf(z[1] - m), e2 = expf(z[2] - m);
								const float inv = 1.0f / (e0 + e1 + e2);
								float *dst = prm.q + (static_cast<size_t>(b) * cells + (row_begin + y) * S + x) * 3;
								dst[0] = e0 * inv;
								dst[1] = e1 * inv;
								dst[2] = e2 * inv;
							}
						}
						else if (L.last_trunk)
						{ // value head 1x1 conv F -> 4 + ReLU on the final trunk output (createValueHead, blocks.cpp:108-111)
							named_barrier_epilogue();
							for (int lc = et; lc < my_rows * S; lc += kEpilogueThreads)
							{
								const int yv = lc / S, xv0 = lc - yv * S;
								const int cell = (row_begin + yv) * S + xv0;
								const uint32_t idx = (yv + 1) * P + xv0 + 1;
								float s4[4] = { 0.f, 0.f, 0.f, 0.f };
								for (int ch = 0; ch < F / 8; ch++)
								{
									const uint4 r = *reinterpret_cast<const uint4*>(buf_x + ch * img_chunk_bytes + idx * 16);
									float xv[8];
									unpack_bf16(r.x, xv[0], xv[1]);
									unpack_bf16(r.y, xv[2], xv[3]);
									unpack_bf16(r.z, xv[4], xv[5]);
									unpack_bf16(r.w, xv[6], xv[7]);
#pragma unroll
									for (int k = 0; k < 4; k++)
#pragma unroll
										for (int j = 0; j < 8; j++)
											s4[k] += xv[j] * svalue_w[k * F + ch * 8 + j];
								}
								float4 o;
								o.x = fmaxf(s4[0] + svalue_w[4 * F + 0], 0.f);
								o.y = fmaxf(s4[1] + svalue_w[4 * F + 1], 0.f);
								o.z = fmaxf(s4[2] + svalue_w[4 * F + 2], 0.f);
								o.w = fmaxf(s4[3] + svalue_w[4 * F + 3], 0.f);
								if (live)
									*reinterpret_cast<float4*>(prm.value_hidden + (static_cast<size_t>(b) * cells + cell) * 4) = o;
							}
						}
						if (prm.trace and bi == 0 and et == 0)
							prm.trace[6 * l + 3] = clock64();
					}
				}
			}
			tc_fence_before();
			__syncthreads();
			cluster_sync(); // the peer may still be reading TMEM / receiving our remote arrivals
			if (warp == 1)
				tmem_dealloc_pair(tmem_base, tmem_cols);
		}

		// ---- value head: dense(4*cells -> D) + ReLU, dense(D -> 3), softmax (createValueHead, blocks.cpp:112-117) --------
		// 0.02 % of the network's FLOPs but 0.9 MB of fp32 weights: a CTA takes kValueBoards boards at a time so that every weight
		// it pulls from L2 feeds that many dot products (one warp per output neuron, lanes stride over the inputs)
		constexpr int kValueBoards = 8;
		__global__ void __launch_bounds__(256) value_head_kernel(const float *__restrict__ hidden, const float *__restrict__ wd1, const float *__restrict__ bd1,
				const float *__restrict__ wd2, const float *__restrict__ bd2, float *__restrict__ value, int n, int in_dim, int D, const int *__restrict__ n_dev,
				const int *__restrict__ gather, int slot_base)
		{
			if (n_dev != nullptr)
				n = *n_dev;
			extern __shared__ float sh[]; // [kValueBoards][in_dim] + [kValueBoards][D]
			float *sx = sh, *sd = sh + kValueBoards * in_dim;
			__shared__ int slot[kValueBoards];
			const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
			for (int b0 = blockIdx.x * kValueBoards; b0 < n; b0 += gridDim.x * kValueBoards)
			{
				const int nb = min(kValueBoards, n - b0);
				if (threadIdx.x < kValueBoards)
					slot[threadIdx.x] = (static_cast<int>(threadIdx.x) < nb) ? (gather ? gather[b0 + threadIdx.x] : b0 + static_cast<int>(threadIdx.x) + slot_base) : -1;
				__syncthreads();
				for (int k = 0; k < kValueBoards; k++)
				{
					const int b = slot[k];
					for (int i = threadIdx.x; i < in_dim; i += blockDim.x)
						sx[k * in_dim + i] = (b >= 0) ? hidden[static_cast<size_t>(b) * in_dim + i] : 0.0f;
				}
				__syncthreads();
				for (int o = warp; o < D; o += nwarps)
				{
					const float *w = wd1 + static_cast<size_t>(o) * in_dim;
					float s[kValueBoards];
#pragma unroll
					for (int k = 0; k < kValueBoards; k++)
						s[k] = 0.f;
					for (int i = lane; i < in_dim; i += 32)
					{
						const float wi = __ldg(w + i);
#pragma unroll
						for (int k = 0; k < kValueBoards; k++)
							s[k] += sx[k * in_dim + i] * wi;
					}
#pragma unroll
					for (int k = 0; k < kValueBoards; k++)
					{
						for (int j = 16; j > 0; j >>= 1)
							s[k] += __shfl_xor_sync(0xFFFFFFFFu, s[k], j);
						if (lane == 0)
							sd[k * D + o] = fmaxf(s[k] + bd1[o], 0.f);
					}
				}
				__syncthreads();
				if (warp < nb)
				{ // one warp per board for the last layer
					const int b = slot[warp];
					float z[3];
					for (int k = 0; k < 3; k++)
					{
						float acc = 0.f;
						for (int i = lane; i < D; i += 32)
							acc += sd[warp * D + i] * wd2[k * D + i];
						for (int o = 16; o > 0; o >>= 1)
							acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
						z[k] = acc + bd2[k];
					}
					if (lane == 0)
					{
						const float m = fmaxf(z[0], fmaxf(z[1], z[2]));
						const float e0 = expf(z[0] - m), e1 = expf(z[1] - m), e2 = expf(z[2] - m);
						const float inv = 1.0f / (e0 + e1 + e2);
						value[b * 3 + 0] = e0 * inv;
						value[b * 3 + 1] = e1 * inv;
						value[b * 3 + 2] = e2 * inv;
					}
				}
				__syncthreads();
			}
		}

		size_t blob_floats(const AgbConfig &c)
		{
			const size_t f = c.filters, cells = static_cast<size_t>(c.rows) * c.cols, d = std::min<size_t>(256, 2 * f);
			size_t n = f * 25 * 32 + f;
			n += static_cast<size_t>(c.blocks) * 2 * (f * 9 * f + f);
			n += f * 9 * f + f + f + 1;
			n += 4 * f + 4 + d * cells * 4 + d + 3 * d + 3;
			if (c.q_head)
				n += f * 9 * f + f + 3 * f + 3;
			return n;
		}
		// fp32 W[F][k][k][cin] -> bf16 weight slices [half][cin/16][tap][2][F/2][8]: 16 input channels at a time (what one part of the previous
		// layer's epilogue produces), all taps of those, each as the two 8-channel core-matrix columns of one K = 16 MMA step
		void append_conv_image(std::vector<uint16_t> &img, const float *w, int F, int k, int cin, int stream_taps)
		{
			const auto to_bf16 = [](float x)
			{
				uint32_t u;
				std::memcpy(&u, &x, 4);
				const uint32_t rounded = u + 0x7FFFu + ((u >> 16) & 1u); // round to nearest even
				return static_cast<uint16_t>(rounded >> 16);
			};
			// half-major: CTA `half` of a pair streams the slices for output channels half*F/2 ..
			for (int half = 0; half < 2; half++)
				for (int c16 = 0; c16 < cin / 16; c16++)
					for (int tap = 0; tap < stream_taps; tap++)
						for (int kc = 2 * c16; kc < 2 * c16 + 2; kc++)
							for (int co = half * (F / 2); co < (half + 1) * (F / 2); co++)
								for (int e = 0; e < 8; e++)
									img.push_back(tap < k * k ? to_bf16(w[(static_cast<size_t>(co) * k * k + tap) * cin + kc * 8 + e]) : static_cast<uint16_t>(0));
		}
	}

	int net_create(AgbEngine *e)
	{
		const AgbConfig &c = e->cfg;
		if (c.filters != 64 and c.filters != 128)
			return e->fail(AGB_EINVAL, "filters must be 64 or 128");
		if (c.rows != c.cols or c.rows > 20 or c.rows < 5)
			return e->fail(AGB_EINVAL, "network kernel supports square boards of 5..20 rows");
		if (1 + 2 * c.blocks + 2 > kMaxConvLayers)
			return e->fail(AGB_EINVAL, "too many blocks");
		e->net = new NetWeights();
		return AGB_OK;
	}
	void net_destroy(AgbEngine *e)
	{
		if (e->net == nullptr)
			return;
		NetWeights *n = e->net;
		void *ptrs[] = { n->d_w_images, n->d_small, n->d_wd1, n->d_bd1, n->d_wd2, n->d_bd2, n->d_value_hidden, n->d_policy, n->d_value, n->d_q, n->d_tickets };
		for (void *p : ptrs)
			if (p)
				cudaFree(p);
		delete n;
		e->net = nullptr;
	}

	int net_load(AgbEngine *e, const float *blob, size_t bytes)
	{
		NetWeights *n = e->net;
		if (n == nullptr)
			return e->fail(AGB_ESTATE, "engine was created without a network (blocks == 0)");
		const AgbConfig &c = e->cfg;
		if (bytes != blob_floats(c) * sizeof(float))
			return e->fail(AGB_EINVAL, "weight blob has the wrong size: expected " + std::to_string(blob_floats(c) * 4) + " bytes");
		const int F = c.filters, S = c.rows, cells = S * S, D = std::min(256, 2 * F);
		NetParams &p = n->params;
		p = NetParams { };
		p.S = S;
		p.P = S + 2;
		p.F = F;
		p.img_rows = ((256 + 2 * p.P + 2) + 7) / 8 * 8;
		p.in_img_rows = ((256 + 4 * p.P + 4) + 7) / 8 * 8;
		p.buf_bytes = (F / 8) * p.img_rows * 16;
		{ // weight ring geometry (tunable: AGB_NET_STAGE_SLICES / AGB_NET_STAGES). One stage = 18 slices = the nine taps of one 16-channel part
			// of a 3x3 layer's input (18 KiB per CTA at 128 filters), which is the unit the straight-line MMA schedule consumes.
			const char *sl = getenv("AGB_NET_STAGE_SLICES"), *ns = getenv("AGB_NET_STAGES");
			p.kc_per_stage = sl ? atoi(sl) : 18;
			p.stage_bytes = p.kc_per_stage * (F / 2) * 16;
			p.n_stages = ns ? atoi(ns) : ((F == 128) ? ((S * (S + 2) > 256) ? 3 : 4) : 6);
			if (p.kc_per_stage < 2 or p.kc_per_stage % 2 != 0 or p.n_stages < 2 or p.n_stages > kMaxStages)
				return e->fail(AGB_EINVAL, "bad weight ring geometry");
		}

		std::vector<uint16_t> images;
		std::vector<float> small; // per-conv biases, then head vectors
		const float *cur = blob;
		int nl = 0;
		const auto add_conv = [&](int k, int cin, int mode, bool last_trunk)
		{
			ConvDesc &L = p.layers[nl++];
			while (small.size() % 4 != 0)
				small.push_back(0.0f); // biases are read as float4
			L.w_offset = static_cast<uint32_t>(images.size() * 2);
			L.bias_offset = static_cast<uint32_t>(small.size());
			L.cin_chunks = static_cast<uint16_t>(cin / 8);
			L.radius = static_cast<uint8_t>(k / 2);
			L.mode = static_cast<uint8_t>(mode);
			L.real_taps = static_cast<uint8_t>(k * k);
			L.n_taps = static_cast<uint8_t>((k * k + 8) / 9 * 9); // 9, or 27 for the 5x5 stem: whole ring stages per 16-channel slice
			L.last_trunk = last_trunk;
			append_conv_image(images, cur, F, k, cin, L.n_taps);
			cur += static_cast<size_t>(F) * k * k * cin;
			small.insert(small.end(), cur, cur + F);
			cur += F;
		};
		add_conv(5, 32, MODE_STEM, c.blocks == 0);
		for (int i = 0; i < c.blocks; i++)
		{
			add_conv(3, F, MODE_CONV1, false);
			add_conv(3, F, MODE_CONV2, i + 1 == c.blocks);
		}
		add_conv(3, F, MODE_POLICY, false);
		const size_t policy_w1_off = small.size();
		small.insert(small.end(), cur, cur + F + 1);
		cur += F + 1;
		const size_t value_w_off = small.size();
		small.insert(small.end(), cur, cur + 4 * F + 4);
		cur += 4 * F + 4;
		const float *wd1 = cur;
		cur += static_cast<size_t>(D) * cells * 4;
		const float *bd1 = cur;
		cur += D;
		const float *wd2 = cur;
		cur += 3 * D;
		const float *bd2 = cur;
		cur += 3;
		size_t q_w1_off = 0;
		if (c.q_head)
		{
			add_conv(3, F, MODE_QHEAD, false);
			q_w1_off = small.size();
			small.insert(small.end(), cur, cur + 3 * F + 3);
			cur += 3 * F + 3;
		}
		p.n_layers = nl;

		const size_t cap = static_cast<size_t>(c.max_boards);
		const auto upload = [&](auto **dst, const void *src, size_t nbytes) -> cudaError_t
		{
			if (*dst)
				cudaFree(*dst);
			cudaError_t err = cudaMalloc(reinterpret_cast<void**>(dst), nbytes);
			if (err != cudaSuccess)
				return err;
			return src ? cudaMemcpyAsync(*dst, src, nbytes, cudaMemcpyHostToDevice, e->stream) : cudaSuccess;
		};
		AGB_CUDA_CHECK(e, upload(&n->d_w_images, images.data(), images.size() * 2));
		AGB_CUDA_CHECK(e, upload(&n->d_small, small.data(), small.size() * 4));
		AGB_CUDA_CHECK(e, upload(&n->d_wd1, wd1, static_cast<size_t>(D) * cells * 4 * 4));
		AGB_CUDA_CHECK(e, upload(&n->d_bd1, bd1, D * 4));
		AGB_CUDA_CHECK(e, upload(&n->d_wd2, wd2, 3 * D * 4));
		AGB_CUDA_CHECK(e, upload(&n->d_bd2, bd2, 3 * 4));
		AGB_CUDA_CHECK(e, upload(&n->d_value_hidden, nullptr, cap * cells * 4 * 4));
		AGB_CUDA_CHECK(e, upload(&n->d_policy, nullptr, cap * cells * 4));
		AGB_CUDA_CHECK(e, upload(&n->d_value, nullptr, cap * 3 * 4));
		AGB_CUDA_CHECK(e, upload(&n->d_q, nullptr, cap * cells * 3 * 4));
		AGB_CUDA_CHECK(e, upload(&n->d_tickets, nullptr, kTicketCounters * sizeof(int)));
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		p.w_images = n->d_w_images;
		p.bias = n->d_small;
		p.policy_w1 = n->d_small + policy_w1_off;
		p.value_w = n->d_small + value_w_off;
		p.q_w1 = c.q_head ? n->d_small + q_w1_off : nullptr;
		p.value_hidden = n->d_value_hidden;
		n->dense_width = D;
		n->smem_bytes = 2 * static_cast<size_t>(p.buf_bytes) + static_cast<size_t>(p.n_stages) * p.stage_bytes + (256 + 8 + 256 + 4 * 128 + 8 + 8) * 4 + 48 * 8;
		n->split = S * (S + 2) > 256; // the board does not fit one CTA's 256 accumulator rows: one board per CTA pair
		if (n->smem_bytes > 232448)
			return e->fail(AGB_EINVAL, "network kernel needs " + std::to_string(n->smem_bytes) + " bytes of shared memory per CTA, more than the device has");
		const void *kernel = n->split ? (F == 128 ? reinterpret_cast<const void*>(resnet_board_kernel<128, true>) : reinterpret_cast<const void*>(resnet_board_kernel<64, true>))
				: (F == 128 ? reinterpret_cast<const void*>(resnet_board_kernel<128, false>) : reinterpret_cast<const void*>(resnet_board_kernel<64, false>));
		AGB_CUDA_CHECK(e, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(n->smem_bytes)));
		AGB_CUDA_CHECK(e, cudaFuncSetAttribute(value_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kValueBoards * (cells * 4 + D) * 4));
		n->loaded = true;
		return AGB_OK;
	}

	int net_forward_impl(AgbEngine *e, const uint32_t *features_dev, int n_boards, const int *n_dev, const int *gather_dev, float *policy_dev, float *value_dev,
			float *q_dev, int slot_base = 0, cudaStream_t stream = nullptr, int max_sms = 0, cudaStream_t tail_stream = nullptr, cudaEvent_t trunk_done = nullptr)
	{
		if (stream == nullptr)
			stream = e->stream;
		NetWeights *n = e->net;
		if (n == nullptr or not n->loaded)
			return e->fail(AGB_ESTATE, "no weights loaded");
		NetParams p = n->params;
		p.features = features_dev;
		p.policy = policy_dev;
		p.q = e->cfg.q_head ? q_dev : nullptr;
		p.n_boards = n_boards;
		p.n_boards_dev = n_dev;
		p.gather = gather_dev;
		p.slot_base = slot_base;
		int ticket_slot = 0;
		while (ticket_slot < n->n_ticket_streams and n->ticket_streams[ticket_slot] != stream)
			ticket_slot++;
		if (ticket_slot == n->n_ticket_streams)
		{
			if (n->n_ticket_streams == kTicketCounters)
				return e->fail(AGB_ESTATE, "network kernel launched on more streams than it has board-ticket counters");
			n->ticket_streams[n->n_ticket_streams++] = stream;
		}
		p.ticket = n->d_tickets + ticket_slot;
		AGB_CUDA_CHECK(e, cudaMemsetAsync(p.ticket, 0, sizeof(int), stream));
		static long long *d_trace = nullptr;
		const bool trace = getenv("AGB_NET_TRACE") != nullptr;
		if (trace and d_trace == nullptr)
			cudaMalloc(&d_trace, kMaxConvLayers * 6 * sizeof(long long));
		p.trace = trace ? d_trace : nullptr;
		int sms = 148;
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->cfg.device);
		if (max_sms > 0)
			sms = std::min(sms, max_sms); // the other SMs are the solver's (AgbConfig::solver_sms)
		const int pairs = std::min(n->split ? n_boards : (n_boards + 1) / 2, sms / 2);
		const int grid = 2 * pairs; // clusters of two CTAs: one board each, or (large boards) one board per pair
		if (n->split)
		{
			if (p.F == 128)
				resnet_board_kernel<128, true><<<grid, kThreads, n->smem_bytes, stream>>>(p);
			else
				resnet_board_kernel<64, true><<<grid, kThreads, n->smem_bytes, stream>>>(p);
		}
		else
		{
			if (p.F == 128)
				resnet_board_kernel<128, false><<<grid, kThreads, n->smem_bytes, stream>>>(p);
			else
				resnet_board_kernel<64, false><<<grid, kThreads, n->smem_bytes, stream>>>(p);
		}
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		if (trace)
		{
			std::vector<long long> h(kMaxConvLayers * 6);
			cudaStreamSynchronize(stream);
			cudaMemcpy(h.data(), d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
			FILE *f = fopen(getenv("AGB_NET_TRACE"), "w");
			if (f)
			{
				for (int l = 0; l < p.n_layers; l++)
					fprintf(f, "%d %lld %lld %lld %lld\n", l, h[6 * l] - h[0], h[6 * l + 1] - h[0], h[6 * l + 2] - h[0], h[6 * l + 3] - h[0]);
				fclose(f);
			}
		}
		const int cells = e->cells, D = n->dense_width;
		if (tail_stream != nullptr and tail_stream != stream)
		{ // pipeline groups: the dense value layers go on the group's own stream, so that the next group's K4 follows this one at once
			AGB_CUDA_CHECK(e, cudaEventRecord(trunk_done, stream));
			AGB_CUDA_CHECK(e, cudaStreamWaitEvent(tail_stream, trunk_done, 0));
			stream = tail_stream;
		}
		value_head_kernel<<<std::min((n_boards + kValueBoards - 1) / kValueBoards, 4 * sms), 256, kValueBoards * (cells * 4 + D) * 4, stream>>>(n->d_value_hidden, n->d_wd1, n->d_bd1, n->d_wd2,
				n->d_bd2, value_dev, n_boards, cells * 4, D, n_dev, gather_dev, slot_base);
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		return AGB_OK;
	}
}

namespace agb
{
	int net_forward_dev(AgbEngine *e, const uint32_t *features_dev, int n_boards, float *policy_dev, float *value_dev, float *q_dev)
	{
		const int rc = net_forward_impl(e, features_dev, n_boards, nullptr, nullptr, policy_dev, value_dev, q_dev);
		if (const char *dump = getenv("AGB_NET_DUMP"))
		{ // diagnostics: every call's inputs and raw outputs appended to a file (n, features, policy, value)
			const size_t cells = e->cells;
			std::vector<uint32_t> f(n_boards * cells);
			std::vector<float> pol(n_boards * cells), val(n_boards * 3);
			cudaStreamSynchronize(e->stream);
			cudaMemcpy(f.data(), features_dev, f.size() * 4, cudaMemcpyDeviceToHost);
			cudaMemcpy(pol.data(), policy_dev, pol.size() * 4, cudaMemcpyDeviceToHost);
			cudaMemcpy(val.data(), value_dev, val.size() * 4, cudaMemcpyDeviceToHost);
			if (FILE *out = fopen(dump, "ab"))
			{
				fwrite(&n_boards, 4, 1, out);
				fwrite(f.data(), 4, f.size(), out);
				fwrite(pol.data(), 4, pol.size(), out);
				fwrite(val.data(), 4, val.size(), out);
				fclose(out);
			}
		}
		return rc;
	}
	int net_forward_dev_gather(AgbEngine *e, const uint32_t *features_dev, const int *count_dev, const int *gather_dev, int max_boards, float *policy_dev,
			float *value_dev, float *q_dev, int slot_base, cudaStream_t stream, int max_sms, cudaStream_t tail_stream, cudaEvent_t trunk_done)
	{
		return net_forward_impl(e, features_dev, max_boards, count_dev, gather_dev, policy_dev, value_dev, q_dev, slot_base, stream, max_sms, tail_stream, trunk_done);
	}
}

extern "C"
{
	size_t agb_weights_size(const AgbEngine *e)
	{
		return (e == nullptr or e->net == nullptr) ? 0 : agb::blob_floats(e->cfg) * sizeof(float);
	}
	int agb_load_weights(AgbEngine *e, const void *blob_host, size_t bytes)
	{
		const agb::DeviceGuard on_device(e);
		if (blob_host == nullptr)
			return e->fail(AGB_EINVAL, "null pointer");
		return agb::net_load(e, static_cast<const float*>(blob_host), bytes);
	}
	int agb_forward_dev(AgbEngine *e, const uint32_t *features_dev, int n, float *policy_dev, float *value_dev, float *q_dev)
	{
		const agb::DeviceGuard on_device(e);
		if (n < 0 or n > e->store.capacity)
			return e->fail(AGB_EINVAL, "n exceeds max_boards");
		if (features_dev == nullptr or policy_dev == nullptr or value_dev == nullptr)
			return e->fail(AGB_EINVAL, "null pointer");
		if (n == 0)
			return AGB_OK;
		return agb::net_forward_dev(e, features_dev, n, policy_dev, value_dev, q_dev);
	}
	int agb_evaluate(AgbEngine *e, const int8_t *boards_host, const int8_t *sign_to_move_host, const int8_t *symmetry_host, int n, float *policy_host,
			float *value_host, float *q_host)
	{
		const agb::DeviceGuard on_device(e);
		if (n < 0 or n > e->store.capacity)
			return e->fail(AGB_EINVAL, "n exceeds max_boards");
		if (boards_host == nullptr or sign_to_move_host == nullptr or policy_host == nullptr or value_host == nullptr)
			return e->fail(AGB_EINVAL, "null pointer");
		if (n == 0)
			return AGB_OK;
		agb::NetWeights *net = e->net;
		if (net == nullptr or not net->loaded)
			return e->fail(AGB_ESTATE, "no weights loaded");
		const size_t cells = e->cells;
		const bool want_q = (q_host != nullptr and e->cfg.q_head);
		const int rc_valid = agb::validate_boards(e, boards_host, sign_to_move_host, static_cast<size_t>(n));
		if (rc_valid != AGB_OK)
			return rc_valid;
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io8, boards_host, n * cells, cudaMemcpyHostToDevice, e->stream));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io8b, sign_to_move_host, n, cudaMemcpyHostToDevice, e->stream));
		int rc = agb::launch_set_boards(e, e->d_io8, e->d_io8b, n, e->d_features); // pack: K1 + K3
		if (rc != AGB_OK)
			return rc;
		const uint32_t *features = e->d_features;
		if (symmetry_host != nullptr)
		{ // NNEvaluator::pack_to_network: features.augment(symmetry)
			for (int i = 0; i < n; i++)
				if (symmetry_host[i] < 0 or symmetry_host[i] > 7)
					return e->fail(AGB_EINVAL, "symmetry must be in 0..7");
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io8b, symmetry_host, n, cudaMemcpyHostToDevice, e->stream));
			rc = agb::launch_augment(e, e->d_features, e->d_features2, e->d_io8b, n);
			if (rc != AGB_OK)
				return rc;
			features = e->d_features2;
		}
		rc = agb::net_forward_dev(e, features, n, net->d_policy, net->d_value, net->d_q); // forward: K4
		if (rc != AGB_OK)
			return rc;
		const float *policy = net->d_policy, *q = net->d_q;
		if (symmetry_host != nullptr)
		{ // unpack_from_network: inverse symmetry on policy and action values
			float *tmp_policy = reinterpret_cast<float*>(e->d_features); // features are dead now: reuse as scratch
			rc = agb::launch_symmetry_f32(e, net->d_policy, tmp_policy, e->d_io8b, n, 1, true);
			if (rc != AGB_OK)
				return rc;
			policy = tmp_policy;
			if (want_q)
			{
				rc = agb::launch_symmetry_f32(e, net->d_q, net->d_value_hidden, e->d_io8b, n, 3, true); // value_hidden [n][cells*4] is free again
				if (rc != AGB_OK)
					return rc;
				q = net->d_value_hidden;
			}
		}
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(policy_host, policy, n * cells * 4, cudaMemcpyDeviceToHost, e->stream));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(value_host, net->d_value, n * 3 * 4, cudaMemcpyDeviceToHost, e->stream));
		if (want_q)
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(q_host, q, n * cells * 3 * 4, cudaMemcpyDeviceToHost, e->stream));
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		if (const char *dump = getenv("AGB_EVAL_DUMP"))
		{ // diagnostics: what the caller gave and got (n, boards, sides to move, symmetries, policy, value)
			if (FILE *out = fopen(dump, "ab"))
			{
				std::vector<int8_t> zeros(n, 0);
				fwrite(&n, 4, 1, out);
				fwrite(boards_host, 1, n * cells, out);
				fwrite(sign_to_move_host, 1, n, out);
				fwrite(symmetry_host != nullptr ? symmetry_host : zeros.data(), 1, n, out);
				fwrite(policy_host, 4, n * cells, out);
				fwrite(value_host, 4, n * 3, out);
				fclose(out);
			}
		}
		return AGB_OK;
	}
	int agb_evaluate_features(AgbEngine *e, const uint32_t *features_host, const int8_t *symmetry_host, int n, float *policy_host, float *value_host, float *q_host)
	{ // NNEvaluator::pack_to_network's first branch (NNEvaluator.cpp:246-251): the caller's feature words go to the network as they are
		const agb::DeviceGuard on_device(e);
		if (n < 0 or n > e->store.capacity)
			return e->fail(AGB_EINVAL, "n exceeds max_boards");
		if (features_host == nullptr or policy_host == nullptr or value_host == nullptr)
			return e->fail(AGB_EINVAL, "null pointer");
		if (n == 0)
			return AGB_OK;
		agb::NetWeights *net = e->net;
		if (net == nullptr or not net->loaded)
			return e->fail(AGB_ESTATE, "no weights loaded");
		const size_t cells = e->cells;
		const bool want_q = (q_host != nullptr and e->cfg.q_head);
		if (symmetry_host != nullptr)
			for (int i = 0; i < n; i++)
				if (symmetry_host[i] < 0 or symmetry_host[i] > 7)
					return e->fail(AGB_EINVAL, "symmetry must be in 0..7");
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_features2, features_host, n * cells * 4, cudaMemcpyHostToDevice, e->stream));
		int rc = agb::net_forward_dev(e, e->d_features2, n, net->d_policy, net->d_value, net->d_q);
		if (rc != AGB_OK)
			return rc;
		const float *policy = net->d_policy, *q = net->d_q;
		if (symmetry_host != nullptr)
		{ // unpack_from_network: inverse symmetry on policy and action values
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io8b, symmetry_host, n, cudaMemcpyHostToDevice, e->stream));
			float *tmp_policy = reinterpret_cast<float*>(e->d_features); // scratch
			rc = agb::launch_symmetry_f32(e, net->d_policy, tmp_policy, e->d_io8b, n, 1, true);
			if (rc != AGB_OK)
				return rc;
			policy = tmp_policy;
			if (want_q)
			{
				rc = agb::launch_symmetry_f32(e, net->d_q, net->d_value_hidden, e->d_io8b, n, 3, true); // value_hidden [n][cells*4] is free again
				if (rc != AGB_OK)
					return rc;
				q = net->d_value_hidden;
			}
		}
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(policy_host, policy, n * cells * 4, cudaMemcpyDeviceToHost, e->stream));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(value_host, net->d_value, n * 3 * 4, cudaMemcpyDeviceToHost, e->stream));
		if (want_q)
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(q_host, q, n * cells * 3 * 4, cudaMemcpyDeviceToHost, e->stream));
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		return AGB_OK;
	}
	int agb_forward(AgbEngine *e, const uint32_t *features_host, int n, float *policy_host, float *value_host, float *q_host)
	{
		const agb::DeviceGuard on_device(e);
		if (n < 0 or n > e->store.capacity)
			return e->fail(AGB_EINVAL, "n exceeds max_boards");
		if (features_host == nullptr or policy_host == nullptr or value_host == nullptr)
			return e->fail(AGB_EINVAL, "null pointer");
		if (n == 0)
			return AGB_OK;
		agb::NetWeights *net = e->net;
		if (net == nullptr or not net->loaded)
			return e->fail(AGB_ESTATE, "no weights loaded");
		const size_t cells = e->cells;
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_features, features_host, n * cells * 4, cudaMemcpyHostToDevice, e->stream));
		const int rc = agb::net_forward_dev(e, e->d_features, n, net->d_policy, net->d_value, net->d_q);
		if (rc != AGB_OK)
			return rc;
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(policy_host, net->d_policy, n * cells * 4, cudaMemcpyDeviceToHost, e->stream));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(value_host, net->d_value, n * 3 * 4, cudaMemcpyDeviceToHost, e->stream));
		if (q_host != nullptr and e->cfg.q_head)
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(q_host, net->d_q, n * cells * 3 * 4, cudaMemcpyDeviceToHost, e->stream));
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		return AGB_OK;
	}
}
