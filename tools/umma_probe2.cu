// Hardware probe for the CTA-pair tensor-core path (tcgen05 cta_group::2): a cluster of two CTAs computes
// D[256 x n] = A[256 x K] * B[n x K]^T where CTA r holds rows 128r..128r+127 of A and rows (n/2)r.. of B in its own shared
// memory (K-major, no swizzle), the leader CTA issues the MMAs, and each CTA reads its 128 rows of D from its own TMEM.
//   umma_probe2 <n 64|128|256> <timing_reps>
#include "../alphagomoku_b200/csrc/umma.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace agb::umma;

constexpr int K = 64, A_ROWS = 160;

__device__ __forceinline__ uint32_t cluster_rank()
{
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *smem_result, uint32_t columns)
{
	asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_result)), "r"(columns) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2()
{
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t columns)
{
	asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(columns) : "memory");
}
__device__ __forceinline__ void mma2_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
			:: "r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate)) : "memory");
}
__device__ __forceinline__ void mma2_commit_multicast(uint64_t *bar, uint16_t cta_mask)
{
	asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
			:: "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe2_kernel(const __nv_bfloat16 *a_img, const __nv_bfloat16 *b_img, float *d, int n,
		int timing_reps, long long *cycles, int mode)
{
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint64_t bar_mma, bar_time;
	__shared__ uint32_t tmem_base;
	const uint32_t rank = cluster_rank();
	const int half_n = n / 2;
	uint8_t *sa = smem; // [K/8][A_ROWS][8]
	uint8_t *sb = smem + (K / 8) * A_ROWS * 16; // [K/8][n/2][8]
	const uint32_t a_bytes = (K / 8) * A_ROWS * 16, b_bytes = (K / 8) * half_n * 16;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0)
	{
		mbar_init(&bar_mma, 1);
		mbar_init(&bar_time, 1);
		fence_mbar_init();
	}
	if (warp == 0)
	{
		tmem_alloc2(&tmem_base, 512);
		tmem_relinquish2();
	}
	for (uint32_t i = threadIdx.x; i < a_bytes / 16; i += blockDim.x)
		reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a_img)[rank * (a_bytes / 16) + i];
	for (uint32_t i = threadIdx.x; i < b_bytes / 16; i += blockDim.x)
		reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_img)[rank * (b_bytes / 16) + i];
	fence_proxy_async();
	tc_fence_before();
	cluster_sync_all();
	tc_fence_after();
	const uint32_t tmem = tmem_base;

	if (rank == 0 and threadIdx.x == 0)
	{
		const uint32_t idesc = idesc_bf16_f32(256, n);
		for (int k = 0; k < K / 16; k++)
		{
			const uint64_t ad = smem_desc(smem_u32(sa) + k * 2 * A_ROWS * 16, A_ROWS * 16, 128);
			const uint64_t bd = smem_desc(smem_u32(sb) + k * 2 * half_n * 16, half_n * 16, 128);
			mma2_bf16(tmem, ad, bd, idesc, k > 0);
		}
		mma2_commit_multicast(&bar_mma, 3);
		if (timing_reps > 0)
		{
			mbar_wait(&bar_mma, 0);
			const long long t0 = clock64();
			if (mode == 0)
			{
				for (int rep = 0; rep < timing_reps; rep++)
					for (int k = 0; k < K / 16; k++)
					{
						const uint64_t ad = smem_desc(smem_u32(sa) + ((rep & 7) * 3) * 16 + k * 2 * A_ROWS * 16, A_ROWS * 16, 128);
						const uint64_t bd = smem_desc(smem_u32(sb) + k * 2 * half_n * 16, half_n * 16, 128);
						mma2_bf16(tmem + 256, ad, bd, idesc, true);
					}
			}
			else
			{ // like resnet.cu: two accumulators per K-step, descriptors advanced incrementally
				const uint64_t a0 = smem_desc(smem_u32(sa), A_ROWS * 16, 128), b0 = smem_desc(smem_u32(sb), half_n * 16, 128);
				for (int rep = 0; rep < timing_reps / 2; rep++)
				{
					uint64_t bd = b0;
					int kin = 0;
					for (int j = 0; j < K / 8; j += 2)
					{
						const uint64_t ad = a0 + (rep & 15) + kin * (A_ROWS);
						mma2_bf16(tmem + 256, ad, bd, idesc, true);
						mma2_bf16(tmem + 256 + (mode == 2 ? 0 : n), ad + 16, bd, idesc, true);
						bd += 2 * half_n;
						kin += 2;
					}
				}
			}
			mma2_commit_multicast(&bar_time, 1);
			mbar_wait(&bar_time, 0);
			if (blockIdx.x == 0)
				cycles[0] = clock64() - t0;
		}
	}
	mbar_wait(&bar_mma, 0);
	tc_fence_after();
	for (int c0 = 0; c0 < n; c0 += 16)
	{
		uint32_t v[16];
		tmem_ld16(tmem + ((warp * 32u) << 16) + c0, v);
		tmem_ld_wait();
		for (int j = 0; j < 16; j++)
			d[(rank * 128 + warp * 32 + lane) * n + c0 + j] = __uint_as_float(v[j]);
	}
	tc_fence_before();
	cluster_sync_all();
	if (warp == 0)
		tmem_dealloc2(tmem, 512);
}

int main(int argc, char **argv)
{
	const int n = argc > 1 ? atoi(argv[1]) : 128, timing_reps = argc > 2 ? atoi(argv[2]) : 0, mode = argc > 3 ? atoi(argv[3]) : 0;
	const int grid = argc > 4 ? atoi(argv[4]) : 2; // > 2: every cluster repeats the same work (chip-wide throughput / power behaviour)
	const int half_n = n / 2;
	std::vector<float> a(2 * A_ROWS * K), b(n * K);
	srand(2);
	for (auto &x : a) x = (rand() % 17 - 8) / 8.0f;
	for (auto &x : b) x = (rand() % 13 - 6) / 4.0f;
	std::vector<__nv_bfloat16> a_img(2 * (K / 8) * A_ROWS * 8), b_img(2 * (K / 8) * half_n * 8);
	for (int r = 0; r < 2; r++)
		for (int row = 0; row < A_ROWS; row++)
			for (int k = 0; k < K; k++)
				a_img[r * (K / 8) * A_ROWS * 8 + ((k / 8) * A_ROWS + row) * 8 + k % 8] = __float2bfloat16(a[(r * A_ROWS + row) * K + k]);
	for (int r = 0; r < 2; r++)
		for (int row = 0; row < half_n; row++)
			for (int k = 0; k < K; k++)
				b_img[r * (K / 8) * half_n * 8 + ((k / 8) * half_n + row) * 8 + k % 8] = __float2bfloat16(b[(r * half_n + row) * K + k]);
	__nv_bfloat16 *da, *db;
	float *dd;
	long long *dc;
	cudaMalloc(&da, a_img.size() * 2);
	cudaMalloc(&db, b_img.size() * 2);
	cudaMalloc(&dd, 256 * n * 4);
	cudaMalloc(&dc, 8);
	cudaMemcpy(da, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice);
	cudaMemcpy(db, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice);
	cudaMemset(dd, 0, 256 * n * 4);
	cudaMemset(dc, 0, 8);
	const int smem_bytes = (K / 8) * (A_ROWS + half_n) * 16;
	cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
	probe2_kernel<<<grid, 128, smem_bytes>>>(da, db, dd, n, timing_reps, dc, mode);
	cudaError_t err = cudaDeviceSynchronize();
	if (err != cudaSuccess)
	{
		printf("2cta n=%d : CUDA ERROR %s\n", n, cudaGetErrorString(err));
		return 1;
	}
	std::vector<float> d(256 * n);
	cudaMemcpy(d.data(), dd, 256 * n * 4, cudaMemcpyDeviceToHost);
	double max_err = 0;
	for (int m = 0; m < 256; m++)
		for (int j = 0; j < n; j++)
		{
			double ref = 0;
			const int r = m / 128, row = m % 128;
			for (int k = 0; k < K; k++)
				ref += (double) a[(r * A_ROWS + row) * K + k] * b[j * K + k];
			max_err = fmax(max_err, fabs(ref - d[m * n + j]));
		}
	printf("2cta n=%d : max_err=%g %s\n", n, max_err, max_err < 1e-3 ? "PASS" : "FAIL");
	if (timing_reps > 0)
	{
		long long c = 0;
		cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
		printf("  timing: %d MMAs (M=256 N=%d K=16, cta_group::2) in %lld cycles = %.1f cycles/MMA\n", timing_reps * 4, n, c, (double) c / (timing_reps * 4));
	}
	return 0;
}
