// The trainer's batch loader on the engine: the reference's torch_api (include/alphagomoku/dataset/torch_api.h:14-41,
// src/dataset/torch_api.cpp:130-281) over GameDataBuffer files (src/dataset/GameDataBuffer.cpp:96-128).
//
// load_batch spends its time in PatternCalculator::setBoard + NNInputFeatures::encode per sample (torch_api.cpp:232-233, ~13.5 us each on a
// host core); here the samples are decoded and augmented on the host (a few hundred bytes each) and the boards of the whole batch go through
// K1 + K3 in one launch, followed by a kernel that expands the 32 feature bits of every cell into the float input tensor.
// The record decoding restates GameDataStorage::getSample / SearchDataStorage_v201::storeTo (GameDataStorage.cpp:109-148,
// SearchDataStorage.cpp:237-274) in float32 step by step (the same arithmetic as alphagomoku_b200/dataset.py, which is checked bit for bit
// against the reference's readers).
#include "engine.hpp"
#include "json_reader.hpp"
#include "patterns_logic.cuh"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace agb
{
	struct DatasetFragment
	{
			int rules = 0, rows = 0, cols = 0;
			std::vector<std::string> games; // GameDataStorage::serialize blobs (format 201)
			std::vector<std::vector<size_t>> sample_offsets; // per game: byte offset of every sample
			std::vector<size_t> moves_offset; // per game: byte offset of its [u32 n_moves][moves]
	};
	struct DatasetStore
	{
			std::map<int, DatasetFragment> fragments;
			float *d_input = nullptr;
			size_t d_input_capacity = 0;
	};
	namespace
	{
		float lowfp(uint32_t x, int exp_bits, int man_bits, int bias)
		{ // LowFP<0, E, M, B>::convert_to_fp32 (include/alphagomoku/utils/low_precision.hpp:141-148)
			const int exponent = static_cast<int>((x >> man_bits) & ((1u << exp_bits) - 1u)) + bias;
			const float base = static_cast<float>(x & ((1u << man_bits) - 1u)) / static_cast<float>(1u << man_bits);
			const int subnormal = (exponent == bias) ? 1 : 0;
			return (1.0f - static_cast<float>(subnormal) + base) * std::ldexp(1.0f, exponent + subnormal);
		}
		uint16_t int8_to_score(uint32_t x)
		{ // int8_to_score (SearchDataStorage.cpp:32-49) -> Score::to_short
			const uint32_t pv = x >> 6, ev = x & 63u;
			if (pv == 0)
				return static_cast<uint16_t>((0u << 13) | (4000u + ev));
			if (pv == 1)
				return static_cast<uint16_t>((1u << 13) | (4000u + ev));
			if (pv == 3)
				return static_cast<uint16_t>((3u << 13) | (4000u - ev));
			// unknown: a signed LowFP<1, 3, 2, -8> evaluation in thousandths
			const uint32_t m = ev & 31u;
			float v = lowfp(m, 3, 2, -8);
			if (ev & 32u)
				v = -v;
			const int e = static_cast<int>(1000.0f * v + 0.5f);
			return static_cast<uint16_t>((2u << 13) | static_cast<uint32_t>(4000 + e));
		}
		bool score_is_proven(uint16_t s)
		{
			return ((s >> 13) & 3) != 2 and s != 0x0000 and s != 0xFFFF;
		}
		uint32_t rd32(const std::string &s, size_t off)
		{
			uint32_t v;
			std::memcpy(&v, s.data() + off, 4);
			return v;
		}
		uint16_t rd16(const std::string &s, size_t off)
		{
			uint16_t v;
			std::memcpy(&v, s.data() + off, 2);
			return v;
		}
		// a decoded sample: the fields of SearchDataPack that load_batch uses
		struct Sample
		{
				std::vector<int8_t> board;
				std::vector<int32_t> visits;
				std::vector<float> win, draw;
				std::vector<uint16_t> scores;
				int outcome = 0, moves_left = 0;
				uint16_t played_move = 0;
		};
		void decode(const DatasetFragment &f, int game, int index, Sample &out)
		{
			const std::string &rec = f.games.at(game);
			const int cells = f.rows * f.cols;
			out.board.assign(cells, 0);
			out.visits.assign(cells, 0);
			out.win.assign(cells, 0.0f);
			out.draw.assign(cells, 0.0f);
			out.scores.assign(cells, static_cast<uint16_t>((2u << 13) | 4000u));
			size_t off = f.sample_offsets.at(game).at(index);
			const float value_scale = lowfp(rd16(rec, off), 5, 11, -16), visit_scale = lowfp(rd16(rec, off + 4), 5, 11, -16);
			const int move_number = rd16(rec, off + 8);
			const uint32_t n_entries = rd32(rec, off + 12);
			off += 16;
			const size_t moves_at = f.moves_offset.at(game);
			const int n_moves = static_cast<int>(rd32(rec, moves_at));
			for (int i = 0; i < move_number and i < n_moves; i++)
			{
				const uint16_t mv = rd16(rec, moves_at + 4 + 2 * i);
				out.board[((mv >> 2) & 127) * f.cols + ((mv >> 9) & 127)] = static_cast<int8_t>(mv & 3);
			}
			int idx = 0;
			for (uint32_t k = 0; k < n_entries; k++, off += 6)
			{
				const uint8_t *e = reinterpret_cast<const uint8_t*>(rec.data() + off);
				idx += e[0];
				const float vf = lowfp(e[1], 3, 5, -8) * visit_scale + 0.5f;
				out.visits[idx] = static_cast<int>(vf);
				float win = lowfp(e[4], 4, 4, -16) * value_scale, draw = lowfp(e[5], 4, 4, -16) * value_scale;
				const float total = win + draw;
				if (total > 1.0f)
				{ // get_valid_value (SearchDataStorage.cpp)
					win = win / total;
					draw = draw / total;
				}
				out.win[idx] = win;
				out.draw[idx] = draw;
				out.scores[idx] = int8_to_score(e[3]);
			}
			out.outcome = static_cast<int>(rd32(rec, moves_at + 4 + 2 * n_moves));
			out.moves_left = n_moves - move_number;
			out.played_move = rd16(rec, moves_at + 4 + 2 * move_number);
		}
		template<typename T>
		void apply_symmetry(std::vector<T> &a, int mode, int S)
		{ // ag::apply_symmetry_in_place (include/alphagomoku/utils/augmentations.hpp): dst(r, c) = src(symmetry_source(r, c))
			if (mode == 0)
				return;
			std::vector<T> tmp(a.size());
			for (int r = 0; r < S; r++)
				for (int c = 0; c < S; c++)
				{
					int sr, sc;
					plogic::symmetry_source(mode, S, r, c, sr, sc);
					tmp[r * S + c] = a[sr * S + sc];
				}
			a.swap(tmp);
		}
		__global__ void unpack_features_kernel(const uint32_t *__restrict__ features, float *__restrict__ input, size_t n_cells)
		{ // one thread per (cell, channel): input[cell][channel] = bit `channel` of the feature word (torch_api.cpp:243-249)
			const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
			if (i < n_cells * 32)
				input[i] = ((features[i >> 5] >> (i & 31)) & 1u) ? 1.0f : 0.0f;
		}
		bool inflate_all(const std::vector<char> &in, std::string &out)
		{
			z_stream zs { };
			if (inflateInit(&zs) != Z_OK)
				return false;
			zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(in.data()));
			zs.avail_in = static_cast<uInt>(in.size());
			char chunk[1 << 16];
			int rc = Z_OK;
			while (rc == Z_OK)
			{
				zs.next_out = reinterpret_cast<Bytef*>(chunk);
				zs.avail_out = sizeof(chunk);
				rc = inflate(&zs, Z_NO_FLUSH);
				out.append(chunk, sizeof(chunk) - zs.avail_out);
			}
			inflateEnd(&zs);
			return rc == Z_STREAM_END;
		}
		DatasetStore* store_of(AgbEngine *e)
		{
			if (e->dataset == nullptr)
				e->dataset = new DatasetStore();
			return e->dataset;
		}
	}
	void dataset_destroy(AgbEngine *e)
	{
		if (e->dataset != nullptr)
		{
			if (e->dataset->d_input)
				cudaFree(e->dataset->d_input);
			delete e->dataset;
			e->dataset = nullptr;
		}
	}
}

extern "C"
{
	using namespace agb;

	int agb_dataset_load_fragment(AgbEngine *e, int index, const char *path)
	{ // load_dataset_fragment -> Dataset::load -> GameDataBuffer::load (GameDataBuffer.cpp:113-128; FileLoader with uncompress)
		if (path == nullptr)
			return e->fail(AGB_EINVAL, "null pointer");
		FILE *f = std::fopen(path, "rb");
		if (f == nullptr)
			return e->fail(AGB_EINVAL, std::string("File '") + path + "' does not exist");
		std::vector<char> raw;
		char chunk[1 << 16];
		size_t n;
		while ((n = std::fread(chunk, 1, sizeof(chunk), f)) > 0)
			raw.insert(raw.end(), chunk, chunk + n);
		std::fclose(f);
		std::string data;
		if (not inflate_all(raw, data))
			return e->fail(AGB_EINVAL, std::string(path) + ": not a zlib-compressed GameDataBuffer file");
		try
		{
			// FileLoader::find_split_point (file_util.cpp:102-118)
			size_t split = data.size();
			int opened = 0;
			for (size_t i = 0; i < data.size(); i++)
			{
				if (data[i] == '{' or data[i] == '[')
					opened++;
				if (data[i] == '}' or data[i] == ']')
					opened--;
				if (opened == 0)
				{
					split = std::min(data.size(), i + 2);
					break;
				}
			}
			const std::string header = data.substr(0, split);
			json::Parser parser(header.c_str());
			const json::Value root = parser.parse();
			const json::Value *format = root.find("format"), *config = root.find("config"), *offsets = root.find("offsets");
			if (format == nullptr or config == nullptr or offsets == nullptr or static_cast<int>(format->number) != 201)
				return e->fail(AGB_EINVAL, std::string(path) + ": not a format-201 GameDataBuffer file");
			DatasetFragment frag;
			const char *rule_names[5] = { "FREESTYLE", "STANDARD", "RENJU", "CARO5", "CARO6" };
			const json::Value *rules = config->find("rules"), *rows = config->find("rows"), *cols = config->find("cols");
			if (rules == nullptr or rows == nullptr or cols == nullptr)
				return e->fail(AGB_EINVAL, std::string(path) + ": header without a GameConfig");
			for (int i = 0; i < 5; i++)
				if (rules->string == rule_names[i])
					frag.rules = i;
			frag.rows = static_cast<int>(rows->number);
			frag.cols = static_cast<int>(cols->number);
			const std::string binary = data.substr(split);
			for (size_t g = 0; g < offsets->array.size(); g++)
			{
				const size_t begin = static_cast<size_t>(offsets->array[g].number);
				// walk one GameDataStorage::serialize: [u32 samples][samples][u32 moves][moves][outcome][rows][cols]
				size_t off = begin;
				const uint32_t n_samples = rd32(binary, off);
				off += 4;
				std::vector<size_t> sample_at;
				for (uint32_t k = 0; k < n_samples; k++)
				{
					sample_at.push_back(off - begin);
					off += 16 + 6 * static_cast<size_t>(rd32(binary, off + 12));
				}
				const size_t moves_at = off - begin;
				off += 4 + 2 * static_cast<size_t>(rd32(binary, off)) + 12;
				if (off > binary.size())
					return e->fail(AGB_EINVAL, std::string(path) + ": truncated record");
				frag.games.push_back(binary.substr(begin, off - begin));
				frag.sample_offsets.push_back(sample_at);
				frag.moves_offset.push_back(moves_at);
			}
			store_of(e)->fragments[index] = std::move(frag);
		}
		catch (std::exception &ex)
		{
			return e->fail(AGB_EINVAL, std::string(path) + ": " + ex.what());
		}
		return AGB_OK;
	}
	int agb_dataset_unload_fragment(AgbEngine *e, int index)
	{
		store_of(e)->fragments.erase(index);
		return AGB_OK;
	}
	int agb_dataset_size(AgbEngine *e, int *n_games, int32_t *sizes_host)
	{ // get_dataset_size (torch_api.cpp:139-165): per game (fragment, game, samples, available symmetries)
		if (n_games == nullptr)
			return e->fail(AGB_EINVAL, "null pointer");
		int count = 0;
		for (const auto &kv : store_of(e)->fragments)
			for (size_t g = 0; g < kv.second.games.size(); g++, count++)
				if (sizes_host != nullptr)
				{
					sizes_host[4 * count + 0] = kv.first;
					sizes_host[4 * count + 1] = static_cast<int32_t>(g);
					sizes_host[4 * count + 2] = static_cast<int32_t>(kv.second.sample_offsets[g].size());
					sizes_host[4 * count + 3] = (kv.second.rows == kv.second.cols) ? 8 : 4;
				}
		*n_games = count;
		return AGB_OK;
	}
	int agb_load_batch(AgbEngine *e, int batch_size, const AgbSample *samples, float *input_host, float *policy_target_host, float *value_target_host,
			float *moves_left_target_host, float *action_values_target_host)
	{
		const agb::DeviceGuard on_device(e);
		if (batch_size <= 0)
			return AGB_OK;
		if (samples == nullptr or input_host == nullptr or policy_target_host == nullptr or value_target_host == nullptr or moves_left_target_host == nullptr
				or action_values_target_host == nullptr)
			return e->fail(AGB_EINVAL, "null pointer");
		DatasetStore *store = store_of(e);
		const auto first = store->fragments.find(samples[0].buffer_index);
		if (first == store->fragments.end())
			return e->fail(AGB_EINVAL, "sample 0 names a dataset fragment that is not loaded");
		const int rows = first->second.rows, cols = first->second.cols, cells = rows * cols;
		if (rows != e->cfg.rows or cols != e->cfg.cols or first->second.rules != e->cfg.rules)
			return e->fail(AGB_EINVAL, "the dataset's GameConfig differs from the engine's");
		std::vector<int8_t> boards(static_cast<size_t>(batch_size) * cells), stm(batch_size);
		Sample s;
		float *policy = policy_target_host, *value = value_target_host, *moves_left = moves_left_target_host;
		for (int b = 0; b < batch_size; b++)
		{
			const auto it = store->fragments.find(samples[b].buffer_index);
			if (it == store->fragments.end() or it->second.rows != rows or it->second.cols != cols or it->second.rules != first->second.rules)
				return e->fail(AGB_EINVAL, "GameConfig mismatch at " + std::to_string(b)); // torch_api.cpp:209-214
			const DatasetFragment &f = it->second;
			if (samples[b].game_index < 0 or samples[b].game_index >= static_cast<int>(f.games.size()) or samples[b].sample_index < 0
					or samples[b].sample_index >= static_cast<int>(f.sample_offsets[samples[b].game_index].size()) or samples[b].augmentation < 0
					or samples[b].augmentation > 7)
				return e->fail(AGB_EINVAL, "sample " + std::to_string(b) + " is out of range");
			decode(f, samples[b].game_index, samples[b].sample_index, s);
			const int mode = samples[b].augmentation;
			apply_symmetry(s.board, mode, rows);
			apply_symmetry(s.visits, mode, rows);
			apply_symmetry(s.win, mode, rows);
			apply_symmetry(s.draw, mode, rows);
			apply_symmetry(s.scores, mode, rows);
			const int sign = s.played_move & 3;
			std::memcpy(boards.data() + static_cast<size_t>(b) * cells, s.board.data(), cells);
			stm[b] = static_cast<int8_t>(sign);
			// convertOutcome (src/search/Value.cpp:16-35): the game's outcome seen by the side to move
			const bool draw = s.outcome == AGB_OUTCOME_DRAW;
			const bool won = (s.outcome == AGB_OUTCOME_CROSS_WIN and sign == CROSS) or (s.outcome == AGB_OUTCOME_CIRCLE_WIN and sign == CIRCLE);
			const float win_rate = (not draw and won) ? 1.0f : 0.0f, draw_rate = draw ? 1.0f : 0.0f;
			value[0] = win_rate;
			value[1] = draw_rate;
			value[2] = 1.0f - (win_rate + draw_rate);
			moves_left[0] = static_cast<float>(s.moves_left);
			float policy_sum = 0.0f;
			for (int i = 0; i < cells; i++)
			{
				const uint16_t score = s.scores[i];
				const int pv = (score >> 13) & 3;
				float w = s.win[i], d = s.draw[i];
				if (score_is_proven(score))
				{ // Score::convertToValue
					w = (pv == 3) ? 1.0f : 0.0f;
					d = (pv == 1) ? 1.0f : 0.0f;
				}
				// like the reference, the action-value target pointer is NOT advanced from sample to sample (torch_api.cpp:251-254, 277-280):
				// every sample writes the first board's slot
				action_values_target_host[i * 3 + 0] = w;
				action_values_target_host[i * 3 + 1] = d;
				action_values_target_host[i * 3 + 2] = 1.0f - (w + d);
				switch (pv)
				{
					case 0:
						policy[i] = 1.0e-6f;
						break;
					case 1:
						policy[i] = static_cast<float>(std::max(1, s.visits[i]));
						break;
					default:
					case 2:
						policy[i] = static_cast<float>(s.visits[i]);
						break;
					case 3:
						policy[i] = 1.0e+6f;
						break;
				}
				policy_sum += policy[i];
			}
			const float tmp = 1.0f / policy_sum;
			for (int i = 0; i < cells; i++)
				policy[i] *= tmp;
			policy += cells;
			value += 3;
			moves_left += 1;
		}
		// the input tensor: K1 + K3 on the whole batch, then 32 floats per cell
		const int capacity = e->store.capacity;
		const size_t chunk_floats = static_cast<size_t>(std::min(batch_size, capacity)) * cells * 32;
		if (store->d_input_capacity < chunk_floats)
		{
			if (store->d_input)
				cudaFree(store->d_input);
			AGB_CUDA_CHECK(e, cudaMalloc(&store->d_input, chunk_floats * sizeof(float)));
			store->d_input_capacity = chunk_floats;
		}
		for (int begin = 0; begin < batch_size; begin += capacity)
		{
			const int count = std::min(capacity, batch_size - begin);
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io8, boards.data() + static_cast<size_t>(begin) * cells, static_cast<size_t>(count) * cells, cudaMemcpyHostToDevice, e->stream));
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io8b, stm.data() + begin, count, cudaMemcpyHostToDevice, e->stream));
			const int rc = launch_set_boards(e, e->d_io8, e->d_io8b, count, e->d_features);
			if (rc != AGB_OK)
				return rc;
			const size_t n_cells = static_cast<size_t>(count) * cells;
			unpack_features_kernel<<<static_cast<unsigned>((n_cells * 32 + 255) / 256), 256, 0, e->stream>>>(e->d_features, store->d_input, n_cells);
			e->launches++;
			AGB_CUDA_CHECK(e, cudaGetLastError());
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(input_host + static_cast<size_t>(begin) * cells * 32, store->d_input, n_cells * 32 * sizeof(float), cudaMemcpyDeviceToHost,
					e->stream));
			AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		}
		uint32_t status = 0;
		const int rc = take_status(e, &status);
		if (rc != AGB_OK)
			return rc;
		if (status != 0)
			return e->fail(AGB_EOVERFLOW, "device-side overflow, flags=" + std::to_string(status));
		return AGB_OK;
	}
}
