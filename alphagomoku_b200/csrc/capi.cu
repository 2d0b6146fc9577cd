// The C ABI (include/agb200.h): argument checking, staging of host buffers, and dispatch to the kernels.
// No compute happens on the host: every entry point launches device kernels and fails if CUDA is unavailable.
#include "engine.hpp"
#include "patterns_logic.cuh"

#include <cstring>
#include <new>
#include <string>

namespace
{
	std::string g_create_error;

	template<typename T>
	cudaError_t dev_alloc(T **ptr, size_t count)
	{
		return cudaMalloc(reinterpret_cast<void**>(ptr), count * sizeof(T));
	}
	int check_status(AgbEngine *e)
	{ // surfaces device-side overflows of bounded structures (never silent)
		uint32_t status = 0;
		const int rc = agb::take_status(e, &status);
		if (rc != AGB_OK)
			return rc;
		if (status & agb::kStatusBadInput)
			return e->fail(AGB_EINVAL, "invalid input reached the device: a board cell outside 0..2, a sign to move outside 1..2, or a move that does not fit "
					"its position (occupied or off-board cell, undo of an absent stone); the offending slots were left unchanged");
		if (status != 0)
			return e->fail(AGB_EOVERFLOW, "device-side overflow, flags=" + std::to_string(status));
		return AGB_OK;
	}
}

namespace agb
{
	// host-side validation of caller buffers (the reference asserts on these; here a bad value would index device arrays out of range)
	int validate_boards(AgbEngine *e, const int8_t *boards, const int8_t *sign_to_move, size_t n)
	{
		const size_t cells = static_cast<size_t>(e->cells);
		unsigned bad = 0;
		for (size_t i = 0; i < n * cells; i++)
			bad |= static_cast<unsigned>(static_cast<uint8_t>(boards[i]) > 2u);
		if (bad)
			return e->fail(AGB_EINVAL, "board cells must be 0 (empty), 1 (cross) or 2 (circle)");
		if (sign_to_move != nullptr)
		{
			for (size_t i = 0; i < n; i++)
				bad |= static_cast<unsigned>(sign_to_move[i] != 1 and sign_to_move[i] != 2);
			if (bad)
				return e->fail(AGB_EINVAL, "sign_to_move must be 1 (cross) or 2 (circle)");
		}
		return AGB_OK;
	}
}

#define AGB_REQUIRE(e, cond, msg)            \
	do                                       \
	{                                        \
		if (!(cond))                         \
			return (e)->fail(AGB_EINVAL, msg); \
	} while (0)
#define AGB_TRY(expr)           \
	do                          \
	{                           \
		const int rc__ = (expr); \
		if (rc__ != AGB_OK)     \
			return rc__;        \
	} while (0)

extern "C"
{
	const char* agb_version(void)
	{
		return "agb200 0.1 (sm_100a)";
	}
	const char* agb_last_error(const AgbEngine *engine)
	{
		return engine ? engine->error.c_str() : g_create_error.c_str();
	}
	int agb_create(const AgbConfig *config, AgbEngine **engine)
	{
		if (config == nullptr or engine == nullptr)
		{
			g_create_error = "null argument";
			return AGB_EINVAL;
		}
		*engine = nullptr;
		if (config->rows != config->cols or config->rows < 5 or config->rows > agb::kMaxSize or config->rules < 0 or config->rules > 4
				or config->max_boards <= 0)
		{
			g_create_error = "unsupported game configuration (square boards of 5..20 cells, rules 0..4, max_boards > 0)";
			return AGB_EINVAL;
		}
		int device_count = 0;
		cudaError_t err = cudaGetDeviceCount(&device_count);
		if (err != cudaSuccess or device_count == 0)
		{ // there is no CPU path: the engine only exists on a GPU
			g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(err);
			return AGB_ECUDA;
		}
		if ((err = cudaSetDevice(config->device)) != cudaSuccess)
		{
			g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(err);
			return AGB_ECUDA;
		}
		AgbEngine *e = new (std::nothrow) AgbEngine();
		if (e == nullptr)
			return AGB_ENOMEM;
		e->cfg = *config;
		e->cells = config->rows * config->cols;
		const auto bail = [&](int rc)
		{
			g_create_error = e->error;
			agb_destroy(e);
			return rc;
		};
		if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess)
			return bail(e->fail(AGB_ECUDA, "cudaStreamCreate failed"));
		cudaDeviceSetLimit(cudaLimitStackSize, 8192); // renju forbidden-move recursion (patterns_logic.cuh)

		const size_t cap = static_cast<size_t>(config->max_boards);
		agb::BoardStore &s = e->store;
		s.capacity = config->max_boards;
		bool ok = true;
		ok = ok and dev_alloc(&s.board, cap * agb::kCellPitch) == cudaSuccess;
		ok = ok and dev_alloc(&s.sign_to_move, cap) == cudaSuccess;
		ok = ok and dev_alloc(&s.lines, cap * agb::kLinePitch) == cudaSuccess;
		ok = ok and dev_alloc(&s.ptypes, cap * agb::kCellPitch) == cudaSuccess;
		ok = ok and dev_alloc(&s.threats, cap * agb::kCellPitch) == cudaSuccess;
		ok = ok and dev_alloc(&s.forbidden, cap * agb::kCellPitch) == cudaSuccess;
		ok = ok and dev_alloc(&s.hist_count, cap * 2 * agb::kHistTypes) == cudaSuccess;
		ok = ok and dev_alloc(&s.hist_cells, cap * 2 * agb::kHistTypes * agb::kCellPitch) == cudaSuccess;
		ok = ok and dev_alloc(&e->d_features, cap * e->cells) == cudaSuccess;
		ok = ok and dev_alloc(&e->d_features2, cap * e->cells) == cudaSuccess;
		ok = ok and dev_alloc(&e->d_io8, cap * e->cells) == cudaSuccess;
		ok = ok and dev_alloc(&e->d_io8b, cap) == cudaSuccess;
		ok = ok and dev_alloc(&e->d_io16, cap) == cudaSuccess;
		ok = ok and dev_alloc(&e->d_status, 4) == cudaSuccess;
		if (not ok)
			return bail(e->fail(AGB_ENOMEM, std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError())));
		cudaMemsetAsync(e->d_status, 0, 16, e->stream);
		cudaMemsetAsync(s.hist_count, 0, cap * 2 * agb::kHistTypes * sizeof(int32_t), e->stream);
		int rc = agb::build_tables(e);
		if (rc != AGB_OK)
			return bail(rc);
		rc = agb::solver_create(e);
		if (rc != AGB_OK)
			return bail(rc);
		if (config->blocks > 0)
		{
			rc = agb::net_create(e);
			if (rc != AGB_OK)
				return bail(rc);
		}
		if (config->games > 0)
		{
			rc = agb::selfplay_create(e);
			if (rc != AGB_OK)
				return bail(rc);
		}
		*engine = e;
		return AGB_OK;
	}
	void agb_destroy(AgbEngine *e)
	{
		if (e == nullptr)
			return;
		const agb::DeviceGuard on_device(e);
		if (e->stream)
			cudaStreamSynchronize(e->stream);
		agb::selfplay_destroy(e);
		agb::solve_scratch_destroy(e);
		agb::openings_destroy(e);
		agb::dataset_destroy(e);
		agb::net_destroy(e);
		agb::BoardStore &s = e->store;
		void *ptrs[] = { s.board, s.sign_to_move, s.lines, s.ptypes, s.threats, s.forbidden, s.hist_count, s.hist_cells, e->d_features, e->d_features2,
				e->d_io8, e->d_io8b, e->d_io16, e->d_status, e->d_pattern, e->d_threat, e->d_def_table };
		for (void *p : ptrs)
			if (p)
				cudaFree(p);
		for (cudaEvent_t ev : e->events)
			cudaEventDestroy(ev);
		if (e->stream)
			cudaStreamDestroy(e->stream);
		delete e;
	}
	int agb_get_config(const AgbEngine *engine, AgbConfig *config)
	{
		if (engine == nullptr or config == nullptr)
			return AGB_EINVAL;
		*config = engine->cfg;
		return AGB_OK;
	}
	int agb_synchronize(AgbEngine *e)
	{
		const agb::DeviceGuard on_device(e);
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		return AGB_OK;
	}
	void* agb_stream(AgbEngine *e)
	{
		return e->stream;
	}

	int agb_get_tables(AgbEngine *e, uint8_t *pattern_types_host, uint8_t *half_open_3_host, uint8_t *threats_host)
	{
		const agb::DeviceGuard on_device(e);
		if (pattern_types_host != nullptr or half_open_3_host != nullptr)
		{
			std::string raw(1u << 20, '\0');
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(&raw[0], e->d_pattern, 1u << 20, cudaMemcpyDeviceToHost, e->stream));
			AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
			for (size_t i = 0; i < (1u << 20); i++)
			{ // split our packed byte back into the reference's two tables
				const uint8_t v = static_cast<uint8_t>(raw[i]);
				if (pattern_types_host)
					pattern_types_host[i] = v & 0x77;
				if (half_open_3_host)
					half_open_3_host[i] = ((v >> 3) & 1) | ((v >> 6) & 2);
			}
		}
		if (threats_host != nullptr)
		{
			uint8_t raw[4096];
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(raw, e->d_threat, 4096, cudaMemcpyDeviceToHost, e->stream));
			AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
			for (int i = 0; i < 4096; i++)
			{
				threats_host[2 * i + 0] = raw[i] & 15;
				threats_host[2 * i + 1] = raw[i] >> 4;
			}
		}
		return AGB_OK;
	}

	int agb_set_boards_dev(AgbEngine *e, const int8_t *boards_dev, const int8_t *sign_to_move_dev, int n, uint32_t *features_dev)
	{
		const agb::DeviceGuard on_device(e);
		AGB_REQUIRE(e, n >= 0 and n <= e->store.capacity, "n exceeds max_boards");
		AGB_REQUIRE(e, boards_dev and sign_to_move_dev and features_dev, "null pointer");
		if (n == 0)
			return AGB_OK;
		return agb::launch_set_boards(e, boards_dev, sign_to_move_dev, n, features_dev);
	}
	int agb_set_boards(AgbEngine *e, const int8_t *boards_host, const int8_t *sign_to_move_host, int n, uint32_t *features_host)
	{
		const agb::DeviceGuard on_device(e);
		AGB_REQUIRE(e, n >= 0 and n <= e->store.capacity, "n exceeds max_boards");
		AGB_REQUIRE(e, boards_host and sign_to_move_host and features_host, "null pointer");
		if (n == 0)
			return AGB_OK;
		const size_t cells = e->cells;
		AGB_TRY(agb::validate_boards(e, boards_host, sign_to_move_host, n));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io8, boards_host, n * cells, cudaMemcpyHostToDevice, e->stream));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io8b, sign_to_move_host, n, cudaMemcpyHostToDevice, e->stream));
		AGB_TRY(agb::launch_set_boards(e, e->d_io8, e->d_io8b, n, e->d_features));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(features_host, e->d_features, n * cells * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
		return check_status(e);
	}
	static int add_undo(AgbEngine *e, const uint16_t *moves_host, int n, bool undo)
	{
		AGB_REQUIRE(e, n >= 0 and n <= e->store.capacity, "n exceeds max_boards");
		AGB_REQUIRE(e, moves_host, "null pointer");
		if (n == 0)
			return AGB_OK;
		for (int i = 0; i < n; i++)
		{ // Move::toShort: sign | row << 2 | col << 9; sign 0 skips the slot
			const int sign = moves_host[i] & 3, row = (moves_host[i] >> 2) & 127, col = (moves_host[i] >> 9) & 127;
			if (sign == 3 or (sign != 0 and (row >= e->cfg.rows or col >= e->cfg.cols)))
				return e->fail(AGB_EINVAL, "move " + std::to_string(i) + " is not on the board (sign " + std::to_string(sign) + ", row " + std::to_string(row)
						+ ", col " + std::to_string(col) + ")");
		}
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io16, moves_host, n * sizeof(uint16_t), cudaMemcpyHostToDevice, e->stream));
		AGB_TRY(agb::launch_add_undo(e, e->d_io16, n, undo));
		return check_status(e); // occupied cell / absent stone is found on the device
	}
	int agb_add_moves(AgbEngine *e, const uint16_t *moves_host, int n)
	{
		const agb::DeviceGuard on_device(e);
		return add_undo(e, moves_host, n, false);
	}
	int agb_undo_moves(AgbEngine *e, const uint16_t *moves_host, int n)
	{
		const agb::DeviceGuard on_device(e);
		return add_undo(e, moves_host, n, true);
	}
	int agb_encode(AgbEngine *e, int n, uint32_t *features_host)
	{
		const agb::DeviceGuard on_device(e);
		AGB_REQUIRE(e, n >= 0 and n <= e->store.capacity, "n exceeds max_boards");
		AGB_REQUIRE(e, features_host, "null pointer");
		if (n == 0)
			return AGB_OK;
		AGB_TRY(agb::launch_encode(e, n, e->d_features));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(features_host, e->d_features, static_cast<size_t>(n) * e->cells * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
		return check_status(e);
	}
	int agb_get_state(AgbEngine *e, int n, uint8_t *pattern_types_host, uint8_t *threats_host, uint8_t *legal_host, uint8_t *forbidden_host,
			int32_t *hist_counts_host, uint16_t *hist_cells_host)
	{
		const agb::DeviceGuard on_device(e);
		AGB_REQUIRE(e, n >= 0 and n <= e->store.capacity, "n exceeds max_boards");
		if (n == 0)
			return AGB_OK;
		const int cells = e->cells;
		const size_t padded = static_cast<size_t>(n) * agb::kCellPitch;
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		if (pattern_types_host)
		{
			std::vector<uint32_t> tmp(padded);
			AGB_CUDA_CHECK(e, cudaMemcpy(tmp.data(), e->store.ptypes, padded * sizeof(uint32_t), cudaMemcpyDeviceToHost));
			for (int b = 0; b < n; b++)
				for (int i = 0; i < cells; i++)
					for (int d = 0; d < 4; d++)
						pattern_types_host[(static_cast<size_t>(b) * cells + i) * 4 + d] = (tmp[static_cast<size_t>(b) * agb::kCellPitch + i] >> (8 * d)) & 0x77;
		}
		if (threats_host)
		{
			std::vector<uint8_t> tmp(padded);
			AGB_CUDA_CHECK(e, cudaMemcpy(tmp.data(), e->store.threats, padded, cudaMemcpyDeviceToHost));
			for (int b = 0; b < n; b++)
				for (int i = 0; i < cells; i++)
				{
					threats_host[(static_cast<size_t>(b) * cells + i) * 2 + 0] = tmp[static_cast<size_t>(b) * agb::kCellPitch + i] & 15;
					threats_host[(static_cast<size_t>(b) * cells + i) * 2 + 1] = tmp[static_cast<size_t>(b) * agb::kCellPitch + i] >> 4;
				}
		}
		if (legal_host)
		{
			std::vector<int8_t> tmp(padded);
			AGB_CUDA_CHECK(e, cudaMemcpy(tmp.data(), e->store.board, padded, cudaMemcpyDeviceToHost));
			for (int b = 0; b < n; b++)
				for (int i = 0; i < cells; i++)
					legal_host[static_cast<size_t>(b) * cells + i] = (tmp[static_cast<size_t>(b) * agb::kCellPitch + i] == agb::NONE);
		}
		if (forbidden_host)
		{
			std::vector<uint8_t> tmp(padded);
			AGB_CUDA_CHECK(e, cudaMemcpy(tmp.data(), e->store.forbidden, padded, cudaMemcpyDeviceToHost));
			for (int b = 0; b < n; b++)
				std::memcpy(forbidden_host + static_cast<size_t>(b) * cells, tmp.data() + static_cast<size_t>(b) * agb::kCellPitch, cells);
		}
		if (hist_counts_host)
			AGB_CUDA_CHECK(e, cudaMemcpy(hist_counts_host, e->store.hist_count, static_cast<size_t>(n) * 2 * agb::kHistTypes * sizeof(int32_t), cudaMemcpyDeviceToHost));
		if (hist_cells_host)
		{
			std::vector<uint16_t> tmp(static_cast<size_t>(n) * 2 * agb::kHistTypes * agb::kCellPitch);
			AGB_CUDA_CHECK(e, cudaMemcpy(tmp.data(), e->store.hist_cells, tmp.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost));
			for (size_t list = 0; list < static_cast<size_t>(n) * 2 * agb::kHistTypes; list++)
				std::memcpy(hist_cells_host + list * cells, tmp.data() + list * agb::kCellPitch, cells * sizeof(uint16_t));
		}
		return AGB_OK;
	}
	int agb_augment(AgbEngine *e, uint32_t *features_host, const int8_t *symmetry_host, int n)
	{
		const agb::DeviceGuard on_device(e);
		AGB_REQUIRE(e, n >= 0 and n <= e->store.capacity, "n exceeds max_boards");
		AGB_REQUIRE(e, features_host and symmetry_host, "null pointer");
		if (n == 0)
			return AGB_OK;
		for (int i = 0; i < n; i++)
			AGB_REQUIRE(e, symmetry_host[i] >= 0 and symmetry_host[i] < 8, "symmetry must be in 0..7");
		const size_t bytes = static_cast<size_t>(n) * e->cells * sizeof(uint32_t);
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_features, features_host, bytes, cudaMemcpyHostToDevice, e->stream));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io8b, symmetry_host, n, cudaMemcpyHostToDevice, e->stream));
		AGB_TRY(agb::launch_augment(e, e->d_features, e->d_features2, e->d_io8b, n));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(features_host, e->d_features2, bytes, cudaMemcpyDeviceToHost, e->stream));
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		return AGB_OK;
	}
	int agb_get_outcomes(AgbEngine *e, const int8_t *boards_host, const uint16_t *last_moves_host, int n, int8_t *outcomes_host)
	{
		const agb::DeviceGuard on_device(e);
		AGB_REQUIRE(e, n >= 0 and n <= e->store.capacity, "n exceeds max_boards");
		AGB_REQUIRE(e, boards_host and last_moves_host and outcomes_host, "null pointer");
		if (n == 0)
			return AGB_OK;
		AGB_TRY(agb::validate_boards(e, boards_host, nullptr, n));
		for (int i = 0; i < n; i++)
			AGB_REQUIRE(e, ((last_moves_host[i] >> 2) & 127) < e->cfg.rows and ((last_moves_host[i] >> 9) & 127) < e->cfg.cols, "last move is not on the board");
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io8, boards_host, static_cast<size_t>(n) * e->cells, cudaMemcpyHostToDevice, e->stream));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(e->d_io16, last_moves_host, n * sizeof(uint16_t), cudaMemcpyHostToDevice, e->stream));
		AGB_TRY(agb::launch_outcomes(e, e->d_io8, e->d_io16, n, e->d_io8b));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(outcomes_host, e->d_io8b, n, cudaMemcpyDeviceToHost, e->stream));
		return check_status(e);
	}
}
