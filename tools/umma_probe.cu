// Hardware probe for alphagomoku_b200/csrc/umma.cuh: one CTA computes D = A[shift:shift+128] * B^T with tcgen05.mma from
// the no-swizzle K-major layout and prints the max error against a host computation. Run on a B200:
//   umma_probe <swap_lbo_sbo 0|1> <row_shift> <n 64|128|256> <use_bulk_copy 0|1>
#include "../alphagomoku_b200/csrc/umma.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace agb::umma;

constexpr int M = 128, K = 64, A_ROWS = 192;

__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16 *a_img, const __nv_bfloat16 *b_img, float *d, int n, int swap, int shift, int bulk, int timing_reps, long long *cycles)
{
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint64_t bar_load, bar_mma;
	__shared__ uint32_t tmem_base;
	uint8_t *sa = smem; // [K/8][A_ROWS][8] bf16
	uint8_t *sb = smem + (K / 8) * A_ROWS * 16; // [K/8][n][8] bf16
	const uint32_t a_bytes = (K / 8) * A_ROWS * 16, b_bytes = (K / 8) * n * 16;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	if (threadIdx.x == 0)
	{
		mbar_init(&bar_load, 1);
		mbar_init(&bar_mma, 1);
		fence_mbar_init();
	}
	if (warp == 0)
	{
		tmem_alloc(&tmem_base, 256);
		tmem_relinquish();
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = tmem_base;

	if (bulk)
	{
		if (threadIdx.x == 0)
		{
			mbar_arrive_expect_tx(&bar_load, a_bytes + b_bytes);
			bulk_g2s(sa, a_img, a_bytes, &bar_load);
			bulk_g2s(sb, b_img, b_bytes, &bar_load);
		}
		mbar_wait(&bar_load, 0);
	}
	else
	{
		for (uint32_t i = threadIdx.x; i < a_bytes / 16; i += blockDim.x)
			reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a_img)[i];
		for (uint32_t i = threadIdx.x; i < b_bytes / 16; i += blockDim.x)
			reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_img)[i];
		fence_proxy_async();
		__syncthreads();
	}

	if (threadIdx.x == 0)
	{
		tc_fence_after();
		const uint32_t idesc = idesc_bf16_f32(M, n);
		uint32_t a_lbo = A_ROWS * 16, a_sbo = 128, b_lbo = n * 16, b_sbo = 128;
		if (swap)
		{
			uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t;
			t = b_lbo; b_lbo = b_sbo; b_sbo = t;
		}
		for (int k = 0; k < K / 16; k++)
		{
			const uint64_t ad = smem_desc(smem_u32(sa) + shift * 16 + k * 2 * A_ROWS * 16, a_lbo, a_sbo);
			const uint64_t bd = smem_desc(smem_u32(sb) + k * 2 * n * 16, b_lbo, b_sbo);
			mma_bf16(tmem, ad, bd, idesc, k > 0);
		}
		mma_commit(&bar_mma);
		if (timing_reps > 0)
		{ // throughput: timing_reps x (2 tiles x 4 K-steps) back-to-back MMAs on resident operands, second accumulator
			mbar_wait(&bar_mma, 0);
			const long long t0 = clock64();
			for (int rep = 0; rep < timing_reps; rep++)
				for (int k = 0; k < K / 16; k++)
				{
					const uint64_t ad = smem_desc(smem_u32(sa) + ((rep & 7) * 3) * 16 + k * 2 * A_ROWS * 16, a_lbo, a_sbo);
					const uint64_t bd = smem_desc(smem_u32(sb) + k * 2 * n * 16, b_lbo, b_sbo);
					mma_bf16(tmem + 128, ad, bd, idesc, true);
				}
			mma_commit(&bar_load);
			mbar_wait(&bar_load, bulk ? 1 : 0);
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
			d[(warp * 32 + lane) * n + c0 + j] = __uint_as_float(v[j]);
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0)
		tmem_dealloc(tmem, 256);
}

int main(int argc, char **argv)
{
	const int swap = argc > 1 ? atoi(argv[1]) : 0, shift = argc > 2 ? atoi(argv[2]) : 0, n = argc > 3 ? atoi(argv[3]) : 128, bulk = argc > 4 ? atoi(argv[4]) : 0;
	const int timing_reps = argc > 5 ? atoi(argv[5]) : 0;
	long long *dcycles;
	cudaMalloc(&dcycles, 8);
	cudaMemset(dcycles, 0, 8);
	std::vector<float> a(A_ROWS * K), b(n * K);
	srand(1);
	for (auto &x : a) x = (rand() % 17 - 8) / 8.0f;
	for (auto &x : b) x = (rand() % 13 - 6) / 4.0f;
	std::vector<__nv_bfloat16> a_img((K / 8) * A_ROWS * 8), b_img((K / 8) * n * 8);
	for (int r = 0; r < A_ROWS; r++)
		for (int k = 0; k < K; k++)
			a_img[((k / 8) * A_ROWS + r) * 8 + k % 8] = __float2bfloat16(a[r * K + k]);
	for (int r = 0; r < n; r++)
		for (int k = 0; k < K; k++)
			b_img[((k / 8) * n + r) * 8 + k % 8] = __float2bfloat16(b[r * K + k]);
	__nv_bfloat16 *da, *db;
	float *dd;
	cudaMalloc(&da, a_img.size() * 2);
	cudaMalloc(&db, b_img.size() * 2);
	cudaMalloc(&dd, M * n * 4);
	cudaMemcpy(da, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice);
	cudaMemcpy(db, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice);
	cudaMemset(dd, 0, M * n * 4);
	const int smem_bytes = (K / 8) * (A_ROWS + n) * 16;
	cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
	probe_kernel<<<1, 128, smem_bytes>>>(da, db, dd, n, swap, shift, bulk, timing_reps, dcycles);
	cudaError_t err = cudaDeviceSynchronize();
	if (err != cudaSuccess)
	{
		printf("swap=%d shift=%d n=%d bulk=%d : CUDA ERROR %s\n", swap, shift, n, bulk, cudaGetErrorString(err));
		return 1;
	}
	std::vector<float> d(M * n);
	cudaMemcpy(d.data(), dd, M * n * 4, cudaMemcpyDeviceToHost);
	double max_err = 0;
	for (int m = 0; m < M; m++)
		for (int j = 0; j < n; j++)
		{
			double ref = 0;
			for (int k = 0; k < K; k++)
				ref += (double) a[(m + shift) * K + k] * b[j * K + k];
			max_err = fmax(max_err, fabs(ref - d[m * n + j]));
		}
	printf("swap=%d shift=%d n=%d bulk=%d : max_err=%g %s\n", swap, shift, n, bulk, max_err, max_err < 1e-3 ? "PASS" : "FAIL");
	if (timing_reps > 0)
	{
		long long c = 0;
		cudaMemcpy(&c, dcycles, 8, cudaMemcpyDeviceToHost);
		printf("  timing: %d MMAs (M=128 N=%d K=16) in %lld cycles = %.1f cycles/MMA\n", timing_reps * 4, n, c, (double) c / (timing_reps * 4));
	}
	return 0;
}
