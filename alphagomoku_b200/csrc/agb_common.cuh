// Shared device/host definitions for libagb200 (sm_100a only).
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define AGB_HD __host__ __device__
#ifdef __CUDA_ARCH__
#define AGB_HD_NOINLINE __host__ __device__ __noinline__ // device pass: keeps the solver's code small enough for the instruction cache
#else
#define AGB_HD_NOINLINE __host__ __device__
#endif
#else
#define AGB_HD
#define AGB_HD_NOINLINE
#endif

namespace agb
{
	constexpr int kMaxSize = 20; // RawPatternCalculator.hpp:48 (a 20-cell line + 2x6 padding cells = 64 bits)
	constexpr int kMaxCells = kMaxSize * kMaxSize;
	constexpr int kMaxLines = 6 * kMaxSize - 2;
	constexpr int kCellPitch = 512; // cells per board slot in every per-cell array (>= 400, keeps slots 512-element aligned)
	constexpr int kLinePitch = 128; // uint64 words per board slot in the line store (>= 118)
	constexpr int kHistTypes = 10; // ThreatType values, ThreatTable.hpp:18-30

	// Sign, Move.hpp:17-23
	enum : int { NONE = 0, CROSS = 1, CIRCLE = 2, ILLEGAL = 3 };
	// PatternType, PatternTable.hpp:22-32
	enum : int { PT_NONE = 0, PT_HALF_OPEN_3, PT_OPEN_3, PT_HALF_OPEN_4, PT_OPEN_4, PT_DOUBLE_4, PT_FIVE, PT_OVERLINE };
	// ThreatType, ThreatTable.hpp:18-30
	enum : int { TT_NONE = 0, TT_HALF_OPEN_3, TT_OPEN_3, TT_FORK_3x3, TT_HALF_OPEN_4, TT_FORK_4x3, TT_FORK_4x4, TT_OPEN_4, TT_FIVE, TT_OVERLINE };
	enum : int { RULE_FREESTYLE = 0, RULE_STANDARD, RULE_RENJU, RULE_CARO5, RULE_CARO6 };
	// bit of the device status word: a kernel met input it cannot use (a cell that is not a Sign, a move onto an occupied or off-board cell,
	// an undo of a stone that is not there); the offending slot is left alone and the host call fails with AGB_EINVAL
	constexpr uint32_t kStatusBadInput = 1u << 12;

	// Device views of the static tables (built once per engine by tables.cu)
	struct Tables
	{
			const uint8_t *pattern; // [1<<20] bits 0-2 cross type, bit 3 cross half-open-3, bits 4-6 circle type, bit 7 circle half-open-3
			const uint8_t *threat; // [4096] low nibble cross ThreatType, high nibble circle ThreatType
	};

	// Structure-of-arrays pattern store for `capacity` board slots (all in HBM)
	struct BoardStore
	{
			int capacity;
			int8_t *board; // [capacity][kCellPitch] Sign per cell
			int8_t *sign_to_move; // [capacity]
			uint64_t *lines; // [capacity][kLinePitch] 2 bits/cell line words, see patterns.cu
			uint32_t *ptypes; // [capacity][kCellPitch] 4 direction bytes (H,V,D,A), each low nibble cross / high nibble circle PatternType
			uint8_t *threats; // [capacity][kCellPitch] low nibble cross / high nibble circle ThreatType
			uint8_t *forbidden; // [capacity][kCellPitch] renju black forbidden flag (valid after set/encode)
			int32_t *hist_count; // [capacity][2][kHistTypes]
			uint16_t *hist_cells; // [capacity][2][kHistTypes][kCellPitch] Location::toShort (row | col << 8) in list order
	};

	AGB_HD inline uint32_t narrow_window(uint32_t w22)
	{ // drop the two centre bits of a 22-bit (11 cell) window -> 20-bit table index (PatternTable.hpp:135-138)
		return (w22 & 1023u) | ((w22 >> 2) & 0xFFC00u);
	}
}

#define AGB_CUDA_CHECK(engine, expr)                                                                  \
	do                                                                                                \
	{                                                                                                 \
		cudaError_t err__ = (expr);                                                                   \
		if (err__ != cudaSuccess)                                                                     \
			return (engine)->fail(AGB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); \
	} while (0)
