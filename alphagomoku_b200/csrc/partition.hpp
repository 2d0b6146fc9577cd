// SM partition between the solver side (K5 and the small tree kernels) and the network kernel (K4) by CUDA green contexts: two disjoint sets
// of SMs, each with its own streams. Kernels launched on a partition's stream only ever run on its SMs, so any number of solver launches
// (one per pipeline group, any block shape) can overlap each other without taking SMs from the network kernel, whose persistent CTAs would
// otherwise be scheduled late. The driver entry points are fetched through cudaGetDriverEntryPoint: the library does not link libcuda.
#pragma once
#include <cuda_runtime.h>

struct AgbEngine;
namespace agb
{
	struct SmPartition
	{
			void *solver_ctx = nullptr, *net_ctx = nullptr, *tree_ctx = nullptr; // CUgreenCtx
			int solver_sms = 0, net_sms = 0, tree_sms = 0; // as provisioned (multiples of 8 on sm_90+, the remainder goes to the network side)
	};
	// splits the device's SMs: `want_tree_sms` for the small tree kernels (select, set-board, expand, make-move: they must never queue behind
	// long-running solver warps), about `want_solver_sms` for the solver kernel, the rest for the network. Returns false (and leaves *out empty)
	// when the driver or the device cannot do it; the caller then uses the register-filling block scheme instead.
	bool partition_create(int device, int want_solver_sms, int want_tree_sms, SmPartition *out, std::string *why);
	bool partition_stream(void *green_ctx, cudaStream_t *stream);
	void partition_destroy(SmPartition *p);
}
