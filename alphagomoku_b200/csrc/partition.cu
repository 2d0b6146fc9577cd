#include <string>
#include "partition.hpp"

#include <cuda.h>

#include <algorithm>
#include <vector>

namespace agb
{
	namespace
	{
		struct Driver
		{
				decltype(&cuDeviceGet) device_get = nullptr;
				decltype(&cuDeviceGetDevResource) get_resource = nullptr;
				decltype(&cuDevSmResourceSplitByCount) split = nullptr;
				decltype(&cuDevResourceGenerateDesc) make_desc = nullptr;
				decltype(&cuGreenCtxCreate) ctx_create = nullptr;
				decltype(&cuGreenCtxDestroy) ctx_destroy = nullptr;
				decltype(&cuGreenCtxStreamCreate) stream_create = nullptr;
				bool ok = false;
		};
		template<typename F>
		bool fetch(const char *name, F *out)
		{
			void *f = nullptr;
			cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
			if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess or q != cudaDriverEntryPointSuccess or f == nullptr)
				return false;
			*out = reinterpret_cast<F>(f);
			return true;
		}
		const Driver& driver()
		{
			static const Driver d = []
			{
				Driver r;
				r.ok = fetch("cuDeviceGet", &r.device_get) and fetch("cuDeviceGetDevResource", &r.get_resource) and fetch("cuDevSmResourceSplitByCount", &r.split)
						and fetch("cuDevResourceGenerateDesc", &r.make_desc) and fetch("cuGreenCtxCreate", &r.ctx_create) and fetch("cuGreenCtxDestroy", &r.ctx_destroy)
						and fetch("cuGreenCtxStreamCreate", &r.stream_create);
				cudaGetLastError();
				return r;
			}();
			return d;
		}
	}
	bool partition_create(int device, int want_solver_sms, int want_tree_sms, SmPartition *out, std::string *why)
	{
		*out = SmPartition { };
		const Driver &d = driver();
		if (not d.ok)
		{
			*why = "the CUDA driver has no green-context entry points";
			return false;
		}
		CUdevice dev;
		CUdevResource all = { }; // zeroed: the driver reads reserved fields of these structs
		if (d.device_get(&dev, device) != CUDA_SUCCESS or d.get_resource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS)
		{
			*why = "cuDeviceGetDevResource failed";
			return false;
		}
		// One split into groups of 8 SMs (the granularity of sm_90+; the results of a split cannot be split again, but a descriptor may combine
		// several of them): group 0 serves the tree kernels, the next ones the solver, the others and the remainder the network.
		constexpr unsigned kMaxParts = 40;
		CUdevResource parts8[kMaxParts] = { }, remainder = { };
		unsigned n_parts = kMaxParts;
		const int rc = d.split(parts8, &n_parts, &all, &remainder, 0, 8u);
		const int tree_parts = (want_tree_sms + 7) / 8, solver_parts = std::max(1, (want_solver_sms + 4) / 8);
		if (rc != CUDA_SUCCESS or static_cast<int>(n_parts) < tree_parts + solver_parts + 1)
		{
			*why = "cuDevSmResourceSplitByCount(8) gave " + std::to_string(n_parts) + " groups (driver result " + std::to_string(rc) + "), not enough for "
					+ std::to_string(want_tree_sms) + " + " + std::to_string(want_solver_sms) + " SMs and a network partition";
			return false;
		}
		std::vector<CUdevResource> tree_res(parts8, parts8 + tree_parts), solver_res(parts8 + tree_parts, parts8 + tree_parts + solver_parts),
				net_res(parts8 + tree_parts + solver_parts, parts8 + n_parts);
		if (remainder.sm.smCount > 0)
			net_res.push_back(remainder);
		const auto count = [](const std::vector<CUdevResource> &v)
		{
			unsigned n = 0;
			for (const CUdevResource &r : v)
				n += r.sm.smCount;
			return static_cast<int>(n);
		};
		std::vector<CUdevResource> *parts[3] = { &solver_res, &net_res, &tree_res };
		CUgreenCtx ctx[3] = { nullptr, nullptr, nullptr };
		for (int i = 0; i < 3; i++)
		{
			CUdevResourceDesc desc;
			int r1 = d.make_desc(&desc, parts[i]->data(), static_cast<unsigned>(parts[i]->size())), r2 = 0;
			if (r1 != CUDA_SUCCESS and i == 1 and remainder.sm.smCount > 0)
			{ // the remainder could not be combined with whole groups: leave those few SMs unused
				net_res.pop_back();
				r1 = d.make_desc(&desc, net_res.data(), static_cast<unsigned>(net_res.size()));
			}
			if (r1 != CUDA_SUCCESS or (r2 = d.ctx_create(&ctx[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM)) != CUDA_SUCCESS)
			{
				for (int j = 0; j < i; j++)
					d.ctx_destroy(ctx[j]);
				*why = "green context " + std::to_string(i) + " of " + std::to_string(parts[i]->size()) + " SM groups could not be created (driver results "
						+ std::to_string(r1) + ", " + std::to_string(r2) + ")";
				return false;
			}
		}
		out->solver_ctx = ctx[0];
		out->net_ctx = ctx[1];
		out->tree_ctx = ctx[2];
		out->solver_sms = count(solver_res);
		out->net_sms = count(net_res);
		out->tree_sms = count(tree_res);
		return true;
	}
	bool partition_stream(void *green_ctx, cudaStream_t *stream)
	{
		CUstream s = nullptr;
		if (not driver().ok or driver().stream_create(&s, static_cast<CUgreenCtx>(green_ctx), CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS)
			return false;
		*stream = s;
		return true;
	}
	void partition_destroy(SmPartition *p)
	{
		if (p->solver_ctx != nullptr)
			driver().ctx_destroy(static_cast<CUgreenCtx>(p->solver_ctx));
		if (p->net_ctx != nullptr)
			driver().ctx_destroy(static_cast<CUgreenCtx>(p->net_ctx));
		if (p->tree_ctx != nullptr)
			driver().ctx_destroy(static_cast<CUgreenCtx>(p->tree_ctx));
		*p = SmPartition { };
	}
}
