// K8: format-201 self-play records written on the device.
//
// Reference: SearchDataPack(rootNode, board) (src/dataset/data_packs.cpp:24-43), SearchDataStorage_v201::loadFrom / serialize
// (src/dataset/SearchDataStorage.cpp:326-374, 410-419), the LowFP<S,E,M,B> mini-float codecs
// (include/alphagomoku/utils/low_precision.hpp:20-156; visit <0,3,5,-8>, policy/value <0,4,4,-16>, score <1,3,2,-8>,
// fp16 <0,5,11,-16>, SearchDataStorage.cpp:22,161-164) and GameDataStorage::serialize (src/dataset/GameDataStorage.cpp:217-251).
// The float operations are kept in the reference's order so the quantised bytes are identical.
#pragma once
#include "agb_common.cuh"

namespace agb
{
	namespace records
	{
		template<int S, int E, int M, int B>
		struct LowFP
		{
				static constexpr int bits = S + E + M;
				static constexpr int max_exponent = (1 << E) - 1 + B;
				static constexpr int min_exponent = B;
				static constexpr uint32_t max_mantissa = (1u << M) - 1u;
				AGB_HD static float scale(int e)
				{ // ldexp(1.0f, -e)
					union { uint32_t u; float f; } v;
					v.u = static_cast<uint32_t>(127 - e) << 23;
					return v.f;
				}
				AGB_HD static uint32_t to_lowp(float x)
				{
					union { float f; uint32_t u; } v;
					v.f = x;
					const uint32_t sign = (S == 1) ? ((v.u & 0x80000000u) >> (32u - bits)) : 0u;
					int exponent = static_cast<int>((v.u & 0x7F800000u) >> 23u) - 127;
					exponent = exponent < min_exponent ? min_exponent : (exponent > max_exponent ? max_exponent : exponent);
					const int is_subnormal = (exponent == min_exponent) ? 1 : 0;
					const float base = ((sign == 0) ? x : (-x)) * scale(exponent + is_subnormal) + is_subnormal - 1;
					const uint32_t m = static_cast<uint32_t>(base * (1 << M) + 0.5f);
					const uint32_t mantissa = m < max_mantissa ? m : max_mantissa;
					return sign | (static_cast<uint32_t>(exponent - B) << M) | mantissa;
				}
				AGB_HD static float to_fp32(uint32_t x)
				{
					const uint32_t sign_mask = (S == 1) ? (1u << (E + M)) : 0u;
					const uint32_t sign = x & sign_mask;
					const int exponent = static_cast<int>((x & (((1u << E) - 1u) << M)) >> M) + B;
					const float base = static_cast<float>(x & max_mantissa) / (1 << M);
					const int is_subnormal = (exponent == min_exponent) ? 1 : 0;
					return ((sign == 0) ? 1.0f : -1.0f) * (1 - is_subnormal + base) / scale(exponent + is_subnormal);
				}
				AGB_HD static float max()
				{
					return to_fp32((S == 0) ? ((1u << bits) - 1u) : ((1u << (bits - 1)) - 1u));
				}
		};
		using visit_format = LowFP<0, 3, 5, -8>;
		using policy_format = LowFP<0, 4, 4, -16>;
		using value_format = LowFP<0, 4, 4, -16>;
		using score_format = LowFP<1, 3, 2, -8>;
		using fp16_format = LowFP<0, 5, 11, -16>;

		AGB_HD inline uint8_t score_to_int8(uint16_t s)
		{ // SearchDataStorage.cpp:24-31
			const int pv = (s >> 13) & 3;
			const bool infinite = (s == 0x0000 or s == 0xFFFF);
			const int eval = (s & 8191) - 4000;
			if (pv != 2 and not infinite)
			{
				int distance = (pv == 3) ? -eval : eval;
				distance = distance < 0 ? 0 : (distance > 63 ? 63 : distance);
				return static_cast<uint8_t>((pv << 6) | distance);
			}
			return static_cast<uint8_t>((pv << 6) | score_format::to_lowp(eval / 1000.0f));
		}
		AGB_HD inline void put16(uint8_t *dst, size_t &off, uint16_t v)
		{
			dst[off++] = static_cast<uint8_t>(v & 0xFF);
			dst[off++] = static_cast<uint8_t>(v >> 8);
		}
		AGB_HD inline void put32(uint8_t *dst, size_t &off, uint32_t v)
		{
			for (int k = 0; k < 4; k++)
				dst[off++] = static_cast<uint8_t>((v >> (8 * k)) & 0xFF);
		}
		// one ply: dense per-cell arrays of the root's edges -> bytes; returns the number of bytes written
		AGB_HD inline size_t serialize_sample_v201(uint8_t *dst, int cells, const int8_t *board, const int32_t *visits, const float *prior, const float *win,
				const float *draw, const uint16_t *scores, uint16_t minimax_score, uint16_t flags)
		{
			int move_number = 0;
			uint32_t entries = 0;
			float policy_scale = 0.0f, value_scale = 0.0f, visit_scale = 1.0f;
			int last_idx = 0;
			for (int i = 0; i < cells; i++)
			{
				const uint16_t sc = scores[i];
				const bool proven = (((sc >> 13) & 3) != 2) and not (sc == 0x0000 or sc == 0xFFFF);
				if (visits[i] > 0 or proven or (i - last_idx) >= 255)
				{
					entries++;
					last_idx = i;
				}
				move_number += (board[i] != NONE);
				policy_scale = policy_scale > prior[i] ? policy_scale : prior[i];
				const float wd = win[i] > draw[i] ? win[i] : draw[i];
				value_scale = value_scale > wd ? value_scale : wd;
				const float vf = static_cast<float>(visits[i]);
				visit_scale = visit_scale > vf ? visit_scale : vf;
			}
			policy_scale = (policy_scale == 0.0f) ? 1.0f : (policy_scale / policy_format::max());
			value_scale = (value_scale == 0.0f) ? 1.0f : (value_scale / policy_format::max());
			visit_scale /= visit_format::max();
			size_t off = 0;
			put16(dst, off, static_cast<uint16_t>(fp16_format::to_lowp(value_scale)));
			put16(dst, off, static_cast<uint16_t>(fp16_format::to_lowp(policy_scale)));
			put16(dst, off, static_cast<uint16_t>(fp16_format::to_lowp(visit_scale)));
			put16(dst, off, minimax_score);
			put16(dst, off, static_cast<uint16_t>(move_number));
			put16(dst, off, flags);
			put32(dst, off, entries);
			last_idx = 0;
			for (int i = 0; i < cells; i++)
			{
				const uint16_t sc = scores[i];
				const bool proven = (((sc >> 13) & 3) != 2) and not (sc == 0x0000 or sc == 0xFFFF);
				if (visits[i] > 0 or proven or (i - last_idx) >= 255)
				{
					dst[off++] = static_cast<uint8_t>(i - last_idx);
					dst[off++] = static_cast<uint8_t>(visit_format::to_lowp(visits[i] / visit_scale));
					dst[off++] = static_cast<uint8_t>(policy_format::to_lowp(prior[i] / policy_scale));
					dst[off++] = score_to_int8(sc);
					dst[off++] = static_cast<uint8_t>(value_format::to_lowp(win[i] / value_scale));
					dst[off++] = static_cast<uint8_t>(value_format::to_lowp(draw[i] / value_scale));
					last_idx = i;
				}
			}
			return off;
		}
		AGB_HD inline size_t max_sample_bytes(int cells)
		{
			return 16 + 6 * static_cast<size_t>(cells);
		}
	}
}
