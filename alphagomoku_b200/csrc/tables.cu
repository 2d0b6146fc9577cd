// Static classification tables, built on the device at engine creation.
//
// Replaces PatternTable::get / ThreatTable::get of the reference (src/patterns/PatternTable.cpp:110-192,
// src/patterns/ThreatTable.cpp:52-96, rule strings of src/patterns/PatternClassifier.cpp:182-327). The reference
// enumerates all 4^10 line windows on one CPU core (~0.9 s per rule); here one thread classifies one window
// (1 Mi threads, a few hundred microseconds) against shape rules held in constant memory.
#include "engine.hpp"
#include "tables_logic.cuh"

#include <cstring>
#include <string>
#include <vector>

namespace agb
{
	using namespace tables_logic;
	namespace
	{
		__constant__ ShapeRule c_rules[2][kMaxRules]; // [colour - 1][priority order]
		__constant__ int c_rule_count[2];

		__global__ void build_pattern_table_kernel(uint8_t *table)
		{
			const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
			if (i < (1u << 20))
				table[i] = pattern_table_entry(i, c_rules[0], c_rule_count[0], c_rules[1], c_rule_count[1]);
		}

	}

	int build_tables(AgbEngine *e)
	{
		for (int colour = CROSS; colour <= CIRCLE; colour++)
		{
			RuleBuilder rb { e->cfg.rules, colour, { } };
			rb.build();
			if (rb.out.size() > kMaxRules)
				return e->fail(AGB_EINVAL, "too many shape rules");
			const int count = static_cast<int>(rb.out.size());
			AGB_CUDA_CHECK(e, cudaMemcpyToSymbol(c_rules, rb.out.data(), rb.out.size() * sizeof(ShapeRule), (colour - 1) * kMaxRules * sizeof(ShapeRule)));
			AGB_CUDA_CHECK(e, cudaMemcpyToSymbol(c_rule_count, &count, sizeof(int), (colour - 1) * sizeof(int)));
		}
		AGB_CUDA_CHECK(e, cudaMalloc(&e->d_pattern, 1u << 20));
		AGB_CUDA_CHECK(e, cudaMalloc(&e->d_threat, 4096));
		build_pattern_table_kernel<<<(1u << 20) / 256, 256, 0, e->stream>>>(e->d_pattern);
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());

		std::vector<uint8_t> threat(4096);
		for (int idx = 0; idx < 4096; idx++)
		{
			const int t[4] = { idx & 7, (idx >> 3) & 7, (idx >> 6) & 7, (idx >> 9) & 7 };
			int cross, circle;
			threat_of(t, e->cfg.rules, cross, circle);
			threat[idx] = static_cast<uint8_t>(cross | (circle << 4));
		}
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_threat, threat.data(), 4096, cudaMemcpyHostToDevice, e->stream));
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		e->tables.pattern = e->d_pattern;
		e->tables.threat = e->d_threat;
		return AGB_OK;
	}
}
