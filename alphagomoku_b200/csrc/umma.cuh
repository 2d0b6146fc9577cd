// Thin inline-PTX wrappers for the Blackwell (sm_100a) tensor-core path: mbarrier, bulk async copy, TMEM allocation,
// tcgen05.mma (kind::f16, operands from shared memory, fp32 accumulators in TMEM) and tcgen05.ld.
//
// Shared-memory operand layout used throughout this repo: K-major, no swizzle ("interleave"), i.e. 8x8-element core
// matrices of 128 contiguous bytes (8 rows x 16 bytes). In units of bytes an operand tile is
//     element (row, k)  ->  base + (k / 8) * LBO + (row / 8) * SBO + (row % 8) * 16 + (k % 8) * 2
// With SBO = 128 the rows of one 8-channel chunk are contiguous (row * 16), so a tile may start at ANY row of a larger
// [chunk][row][8] image: shifting the start address by 16 bytes shifts the operand by one row. resnet.cu relies on this
// to read the nine shifted views of a padded board image without im2col.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>

namespace agb
{
	namespace umma
	{
		__device__ __forceinline__ uint32_t smem_u32(const void *p)
		{
			return static_cast<uint32_t>(__cvta_generic_to_shared(p));
		}
		// ---- mbarrier -----------------------------------------------------------------------------------------
		__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
		{
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
		}
		__device__ __forceinline__ void fence_mbar_init()
		{
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
		{
			asm volatile("{\n\t.reg .b64 state;\n\tmbarrier.arrive.shared::cta.b64 state, [%0];\n\t}" :: "r"(smem_u32(bar)) : "memory");
		}
		__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
		{
			asm volatile("{\n\t.reg .b64 state;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 state, [%0], %1;\n\t}" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
		}
		__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
		{
			uint32_t ok;
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
					: "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
			return ok != 0;
		}
		__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
		{
			while (not mbar_try_wait(bar, parity))
			{
			}
		}
		// for roles that wait a long time next to the MMA-issuing lane: back off so the spin does not take issue slots from it
		__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity, unsigned ns = 64)
		{
			while (not mbar_try_wait(bar, parity))
				__nanosleep(ns);
		}
		// generic-proxy shared-memory writes -> visible to the async proxy (TMA / tensor core reads)
		__device__ __forceinline__ void fence_proxy_async()
		{
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		}
		// ---- bulk async copy global -> shared, completion on an mbarrier (bytes multiple of 16, 16-byte aligned) ----
		__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
		{
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
					:: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
		}
		// ---- TMEM -------------------------------------------------------------------------------------------------
		__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result, uint32_t columns)
		{ // one full warp; columns: power of two >= 32
			asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_result)), "r"(columns) : "memory");
		}
		__device__ __forceinline__ void tmem_relinquish()
		{
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
		}
		__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr, uint32_t columns)
		{
			asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_addr), "r"(columns) : "memory");
		}
		__device__ __forceinline__ void tc_fence_before()
		{
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
		}
		__device__ __forceinline__ void tc_fence_after()
		{
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		}
		// ---- descriptors ------------------------------------------------------------------------------------------
		// shared-memory matrix descriptor, no swizzle; lbo = byte stride between core matrices adjacent in K,
		// sbo = byte stride between core matrices adjacent in M/N (cute::UMMA::SmemDescriptor, version 1)
		__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
		{
			uint64_t d = 0;
			d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
			d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
			d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
			d |= 1ull << 46; // descriptor version (Blackwell)
			return d; // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
		}
		// instruction descriptor: bf16 x bf16 -> f32, both operands K-major (cute::UMMA::InstrDescriptor)
		__host__ __device__ constexpr uint32_t idesc_bf16_f32(int m, int n)
		{
			return (1u << 4) /* D = f32 */ | (1u << 7) /* A = bf16 */ | (1u << 10) /* B = bf16 */ | (static_cast<uint32_t>(n >> 3) << 17)
					| (static_cast<uint32_t>(m >> 4) << 24);
		}
		// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues for the CTA
		__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
		{
			asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
					:: "r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate)) : "memory");
		}
		// make all previously issued MMAs arrive on an mbarrier when they complete (implies fence::before_thread_sync)
		__device__ __forceinline__ void mma_commit(uint64_t *bar)
		{
			asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
		}
		// ---- TMEM -> registers: warp w of a warpgroup reads lanes 32*(w%4).., thread i gets lane i, 16 consecutive columns ---
		__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
		{
			asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
					: "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
					  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
					: "r"(taddr) : "memory");
		}
		__device__ __forceinline__ void tmem_ld_wait()
		{
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
		}
	}
}

// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster share one MMA. Each CTA supplies 128 rows of A and half of
// the rows of B from its own shared memory (same offsets in both), the leader CTA issues, and each CTA finds its 128 rows
// of D in its own TMEM. Validated on hardware by tools/umma_probe2.cu (64 cycles per M=256,N=128,K=16 instruction).
namespace agb
{
	namespace umma
	{
		// one lane of a converged warp (elect.sync); used to issue single-thread instructions from warp-uniform code so that
		// descriptors stay in uniform registers
		__device__ __forceinline__ bool elect_one()
		{
			uint32_t pred = 0;
			asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
			return pred != 0;
		}
		__device__ __forceinline__ uint32_t cluster_ctarank()
		{
			uint32_t r;
			asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
			return r;
		}
		__device__ __forceinline__ void cluster_sync()
		{
			asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
		}
		__device__ __forceinline__ void tmem_alloc_pair(uint32_t *smem_result, uint32_t columns)
		{
			asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_result)), "r"(columns) : "memory");
		}
		__device__ __forceinline__ void tmem_relinquish_pair()
		{
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
		}
		__device__ __forceinline__ void tmem_dealloc_pair(uint32_t tmem_addr, uint32_t columns)
		{
			asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_addr), "r"(columns) : "memory");
		}
		__device__ __forceinline__ void mma_pair_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
		{
			asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
					:: "r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate)) : "memory");
		}
		// arrive on the mbarrier at this shared-memory offset in every CTA of `cta_mask` once the issued MMAs have completed
		__device__ __forceinline__ void mma_pair_commit(uint64_t *bar, uint16_t cta_mask)
		{
			asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
					:: "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
		}
		// arrive (release, cluster scope) on the mbarrier that sits at the same offset as `bar` in CTA `target_rank` of the cluster
		__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t target_rank)
		{
			asm volatile("{\n\t.reg .b32 raddr;\n\tmapa.shared::cluster.u32 raddr, %0, %1;\n\tmbarrier.arrive.release.cluster.shared::cluster.b64 _, [raddr];\n\t}"
					:: "r"(smem_u32(bar)), "r"(target_rank) : "memory");
		}
		// wait with cluster-scope acquire: pairs with remote arrivals from the peer CTA
		// ---- distributed shared memory of the CTA pair (board split over two CTAs) ----
		__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t target_rank)
		{
			uint32_t r;
			asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(target_rank));
			return r;
		}
		__device__ __forceinline__ void st_peer_v4(uint32_t cluster_addr, uint4 v)
		{
			asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
		}
		__device__ __forceinline__ void st_peer_u32(uint32_t cluster_addr, uint32_t v)
		{
			asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(cluster_addr), "r"(v) : "memory");
		}
		__device__ __forceinline__ void st_peer_f32(uint32_t cluster_addr, float v)
		{
			asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(cluster_addr), "f"(v) : "memory");
		}
		__device__ __forceinline__ void fence_proxy_async_all()
		{ // generic-proxy writes (also those into the peer's shared memory) before async-proxy reads (tcgen05.mma operands)
			asm volatile("fence.proxy.async;" ::: "memory");
		}
		__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar)
		{ // local barrier, cluster-scope release: publishes this thread's writes into the peer's shared memory as well
			asm volatile("{\n\t.reg .b64 state;\n\tmbarrier.arrive.release.cluster.shared::cta.b64 state, [%0];\n\t}" :: "r"(smem_u32(bar)) : "memory");
		}
		__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
		{
			uint32_t ok = 0;
			while (not ok)
				asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
						: "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
		}
	}
}
