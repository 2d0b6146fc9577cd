// Probe: do green contexts (CUDA driver API, fetched through cudaGetDriverEntryPoint -- no link against libcuda) give two disjoint SM
// partitions whose streams accept runtime-API launches, events and cluster kernels? Prints the SM ids each partition's blocks ran on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/green_probe tools/green_probe.cu && tools/green_probe 56
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

#define DRV(name) decltype(&name) p_##name = nullptr; { void *f = nullptr; cudaDriverEntryPointQueryResult q; \
	if (cudaGetDriverEntryPoint(#name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("no %s\n", #name); return 1; } \
	p_##name = reinterpret_cast<decltype(&name)>(f); }
#define CK(x) do { CUresult r = (x); if (r != CUDA_SUCCESS) { printf("%s failed: %d\n", #x, (int) r); return 1; } } while (0)
#define RT(x) do { cudaError_t r = (x); if (r != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(r)); return 1; } } while (0)

__global__ void spin_kernel(int *smids, long long cycles)
{
	unsigned smid;
	asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
	if (threadIdx.x == 0)
		smids[blockIdx.x] = smid;
	const long long t0 = clock64();
	while (clock64() - t0 < cycles)
	{
	}
}
__global__ void __cluster_dims__(2, 1, 1) cluster_kernel(int *smids, long long cycles)
{
	extern __shared__ char big[];
	unsigned smid;
	asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
	if (threadIdx.x == 0)
		smids[blockIdx.x] = smid;
	big[threadIdx.x] = 1;
	const long long t0 = clock64();
	while (clock64() - t0 < cycles)
	{
	}
}

int main(int argc, char **argv)
{
	const unsigned want = argc > 1 ? atoi(argv[1]) : 56;
	RT(cudaSetDevice(0));
	RT(cudaFree(0));
	DRV(cuDeviceGet) DRV(cuDeviceGetDevResource) DRV(cuDevSmResourceSplitByCount) DRV(cuDevResourceGenerateDesc) DRV(cuGreenCtxCreate)
	DRV(cuGreenCtxStreamCreate) DRV(cuGreenCtxDestroy)
	CUdevice dev;
	CK(p_cuDeviceGet(&dev, 0));
	CUdevResource all, part, rest;
	CK(p_cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
	unsigned groups = 1;
	CK(p_cuDevSmResourceSplitByCount(&part, &groups, &all, &rest, 0, want));
	printf("device SMs %u: partition %u + remaining %u (asked %u)\n", all.sm.smCount, part.sm.smCount, rest.sm.smCount, want);
	CUdevResourceDesc d1, d2;
	CK(p_cuDevResourceGenerateDesc(&d1, &part, 1));
	CK(p_cuDevResourceGenerateDesc(&d2, &rest, 1));
	CUgreenCtx g1, g2;
	CK(p_cuGreenCtxCreate(&g1, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM));
	CK(p_cuGreenCtxCreate(&g2, d2, dev, CU_GREEN_CTX_DEFAULT_STREAM));
	CUstream s1a, s1b, s2;
	CK(p_cuGreenCtxStreamCreate(&s1a, g1, CU_STREAM_NON_BLOCKING, 0));
	CK(p_cuGreenCtxStreamCreate(&s1b, g1, CU_STREAM_NON_BLOCKING, 0));
	CK(p_cuGreenCtxStreamCreate(&s2, g2, CU_STREAM_NON_BLOCKING, 0));
	int *a, *b, *c;
	const int n = 1024;
	RT(cudaMalloc(&a, n * 4)); RT(cudaMalloc(&b, n * 4)); RT(cudaMalloc(&c, n * 4));
	RT(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	cudaEvent_t e0, e1, e2, e3;
	RT(cudaEventCreate(&e0)); RT(cudaEventCreate(&e1)); RT(cudaEventCreate(&e2)); RT(cudaEventCreate(&e3));
	RT(cudaEventRecord(e0, s1a));
	spin_kernel<<<n, 64, 0, s1a>>>(a, 2000000);   // many small blocks in partition 1, stream a
	spin_kernel<<<n, 896, 0, s1b>>>(b, 2000000);  // whole-SM blocks in partition 1, stream b (co-resident launches)
	RT(cudaEventRecord(e2, s2));
	cluster_kernel<<<rest.sm.smCount & ~1u, 192, 200 * 1024, s2>>>(c, 200000000); // one 200 KB CTA per SM, pairs, in partition 2
	RT(cudaGetLastError());
	RT(cudaEventRecord(e1, s1a));
	RT(cudaEventRecord(e3, s2));
	RT(cudaStreamWaitEvent(s1b, e3, 0)); // cross-partition event dependency
	RT(cudaDeviceSynchronize());
	float ms1, ms2;
	RT(cudaEventElapsedTime(&ms1, e0, e1)); RT(cudaEventElapsedTime(&ms2, e2, e3));
	std::vector<int> ha(n), hb(n), hc(n);
	RT(cudaMemcpy(ha.data(), a, n * 4, cudaMemcpyDeviceToHost)); RT(cudaMemcpy(hb.data(), b, n * 4, cudaMemcpyDeviceToHost)); RT(cudaMemcpy(hc.data(), c, n * 4, cudaMemcpyDeviceToHost));
	std::set<int> sa(ha.begin(), ha.end()), sb(hb.begin(), hb.end()), sc(hc.begin(), hc.begin() + (rest.sm.smCount & ~1u));
	int overlap = 0;
	for (int x : sc) overlap += sa.count(x) + sb.count(x);
	printf("partition 1 stream a ran on %zu SMs, stream b on %zu SMs; partition 2 cluster kernel on %zu SMs; SM ids shared between the partitions: %d\n", sa.size(), sb.size(), sc.size(), overlap);
	printf("stream a spin kernel %.2f ms (1024 blocks x 1.0 ms spin), partition 2 kernel %.2f ms (one 100 ms wave expected)\n", ms1, ms2);
	printf("partition 1 SM ids:");
	for (int x : sa) printf(" %d", x);
	printf("\n");
	return 0;
}
