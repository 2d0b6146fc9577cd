// Host-side engine object behind the C ABI (include/agb200.h).
#pragma once
#include "../../include/agb200.h"
#include "agb_common.cuh"

#include <string>
#include <vector>

namespace agb
{
	struct NetWeights; // resnet.cu
	struct SelfplayState; // tree.cu
}

namespace agb
{
	struct SolveScratch;
	struct DatasetStore; // dataset_api.cu: GameDataBuffer fragments loaded for agb_load_batch
}
struct AgbEngine
{
		AgbConfig cfg { };
		int cells = 0;
		cudaStream_t stream = nullptr;
		std::string error;
		uint64_t launches = 0;
		uint64_t nn_kernel_ns = 0, nn_kernel_launches = 0, nn_positions = 0; // device timing of K4 inside agb_step
		uint64_t solver_kernel_ns = 0; // device timing of K5 inside agb_step
		std::vector<cudaEvent_t> events; // reusable event pool for that timing

		// static tables
		uint8_t *d_pattern = nullptr; // [1<<20]
		uint8_t *d_threat = nullptr; // [4096]
		uint16_t *d_def_table = nullptr; // [15][256][2] defensive move masks (solver_logic.cuh)
		agb::Tables tables { };

		// pattern store
		agb::BoardStore store { };
		uint32_t *d_features = nullptr; // [max_boards][cells] staging for host-pointer entry points
		uint32_t *d_features2 = nullptr; // [max_boards][cells]
		int8_t *d_io8 = nullptr; // [max_boards][cells] staging
		int8_t *d_io8b = nullptr; // [max_boards]
		uint16_t *d_io16 = nullptr; // [max_boards]
		uint32_t *d_status = nullptr; // device overflow / error word (cleared every time it is reported, see take_status)
		uint32_t overflow_seen = 0; // every flag reported since creation or the last agb_selfplay_reset / agb_load_games (AgbStats::overflow_flags)

		agb::NetWeights *net = nullptr;
		agb::SelfplayState *selfplay = nullptr;
		agb::SolveScratch *solve_scratch = nullptr; // agb_solve: solver memory for positions outside the lockstep engine
		std::vector<uint64_t> solver_keys_host; // Zobrist words given through agb_set_solver_keys (applied to every solver state)
		agb::DatasetStore *dataset = nullptr;
		void *opening_rng = nullptr; // std::mt19937 of the opening generator (openings.cu)
		bool think_mode = false; // agb_think in progress: games stop at their decision

		int fail(int code, const std::string &msg)
		{
			error = msg;
			return code;
		}
};

namespace agb
{
	// Every entry point that takes an engine runs with the engine's device current and restores the caller's afterwards: one process may hold
	// engines on several GPUs (one GeneratorThread per DeviceConfig in the reference; two engines of an arena on different devices).
	struct DeviceGuard
	{
			int previous = -1;
			explicit DeviceGuard(const AgbEngine *e)
			{
				if (e != nullptr and cudaGetDevice(&previous) == cudaSuccess and previous != e->cfg.device)
					cudaSetDevice(e->cfg.device);
				else
					previous = -1;
			}
			~DeviceGuard()
			{
				if (previous >= 0)
					cudaSetDevice(previous);
			}
			DeviceGuard(const DeviceGuard&) = delete;
			DeviceGuard& operator=(const DeviceGuard&) = delete;
	};
	// The evaluator's randInt(8) per scheduled task as a counter-based stream keyed by (seed, global game id, evaluations drawn so far), so that
	// results do not depend on how games are sharded; or, for replaying a reference run, a caller-supplied table indexed by that counter.
	struct SymmetryStream
	{
			int8_t *task_sym = nullptr; // [slots]
			uint32_t *counter = nullptr; // [games]
			unsigned long long seed = 0;
			int first_game_id = 0;
			const int8_t *table = nullptr; // agb_set_symmetry_table
			int table_size = 0;
#ifdef __CUDACC__
			__device__ int draw(int game) const
			{
				const uint32_t i = counter[game]++;
				if (table != nullptr)
					return table[i % static_cast<uint32_t>(table_size)];
				unsigned long long z = seed ^ (static_cast<unsigned long long>(first_game_id + game) << 32) ^ i;
				z += 0x9E3779B97F4A7C15ull;
				z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
				z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
				return static_cast<int>((z ^ (z >> 31)) & 7ull);
			}
#endif
	};
	// Reads the device status word on the engine's stream and clears it, so that one overflow is reported once and the engine stays
	// usable afterwards (the flags stay visible in AgbStats::overflow_flags). Synchronises the stream.
	inline int take_status(AgbEngine *e, uint32_t *status)
	{
		*status = 0;
		cudaError_t err = cudaMemcpyAsync(status, e->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream);
		if (err == cudaSuccess)
			err = cudaMemsetAsync(e->d_status, 0, sizeof(uint32_t), e->stream);
		if (err == cudaSuccess)
			err = cudaStreamSynchronize(e->stream);
		if (err != cudaSuccess)
			return e->fail(AGB_ECUDA, std::string("status word: ") + cudaGetErrorString(err));
		e->overflow_seen |= *status;
		return AGB_OK;
	}
	// per-slot outputs of the solver (K5)
	struct SolverOutputs
	{
			uint16_t *moves; // [slots][pitch] Move::toShort, in action-list order
			uint16_t *scores; // [slots][pitch]
			int32_t *n_actions;
			uint16_t *score; // position score
			uint8_t *must_defend;
			int32_t *nodes; // positions visited by the search (1: solved statically)
			int pitch;
	};
	// per-game memory of the solver: transposition table, action stack, frames, and the slots of the current batch in task order
	struct SolverState
	{
			static constexpr int kFrameBytes = 24;
			int games = 0, batch = 0;
			size_t table_entries = 0; // per game, power of two
			uint64_t *table = nullptr; // [games][table_entries][2]
			uint64_t *keys = nullptr; // [key sets][2 * cells][2] Zobrist words (low, high) per (cell, colour)
			size_t keys_stride = 0; // words between the key sets of consecutive games (0: all games share one set)
			int32_t *generation = nullptr; // [games] SharedHashTable::m_base_generation
			int32_t *game_slots = nullptr; // [games][batch]
			int32_t *game_slot_count = nullptr; // [games]
			uint16_t *stack_moves = nullptr, *stack_scores = nullptr; // [games][stack_capacity]
			int stack_capacity = 0;
			void *frames = nullptr; // [games][kMaxFrames] solver::Frame
			void *children = nullptr; // [games][cells] solver::ChildInfo (previews of the root actions)
			unsigned long long *game_cycles = nullptr; // [games][2] SM clocks and positions visited by the last launch (load-balance diagnostics)
			uint32_t *game_work = nullptr; // [games] estimated cost of the game in the last launch (next launch's priorities)
			int32_t *order = nullptr; // [games] games of a launch, most expensive (by the previous launch's clocks) first
			int32_t *next = nullptr; // [games] work-queue heads, one per launch range (indexed by its first game)
			const uint16_t *def_table = nullptr;
			// SelfplayConfig::use_symmetries with the solver on: the evaluator draws a symmetry for the tasks that are actually scheduled to the
			// network, in task order (Search::scheduleToNN -> NNEvaluator::addToQueue, Search.cpp:184-198, NNEvaluator.cpp:134-139) -- which
			// only the solver kernel knows. Unset (task_sym == nullptr): no symmetries.
			SymmetryStream sym { };
	};
	// capi.cu: caller-supplied boards hold only Sign values 0..2 and sign_to_move (may be NULL) only 1..2, else AGB_EINVAL
	int validate_boards(AgbEngine *e, const int8_t *boards, const int8_t *sign_to_move, size_t n);
	// solver.cu
	int solver_create(AgbEngine *e);
	int solver_state_create(AgbEngine *e, int games, int batch, SolverState *st);
	int solver_state_reset(AgbEngine *e, SolverState *st);
	void solver_state_destroy(SolverState *st);
	void solve_scratch_destroy(AgbEngine *e);
	void openings_destroy(AgbEngine *e);
	void dataset_destroy(AgbEngine *e);
	// solver_sms > 0: run on that many SMs only (side by side with the network kernel, see AgbConfig::solver_sms). green: the stream belongs to a
	// green context of that many SMs, any block shape stays inside it; otherwise the launch uses blocks that take whole SMs
	int launch_solve_games(AgbEngine *e, const SolverState &st, int game_begin, int game_count, const SolverOutputs &out, const uint8_t *slot_is_root, int *nn_list,
			int *nn_count, cudaStream_t stream, int solver_sms = 0, bool green = false);
	// tables.cu
	int build_tables(AgbEngine *e);
	// patterns.cu
	int launch_set_boards(AgbEngine *e, const int8_t *boards_dev, const int8_t *stm_dev, int n, uint32_t *features_dev);
	int launch_set_boards_counted(AgbEngine *e, const int8_t *boards_dev, const int8_t *stm_dev, const int *n_dev, int max_n, uint32_t *features_dev,
			int slot_base = 0, cudaStream_t stream = nullptr);
	int launch_add_undo(AgbEngine *e, const uint16_t *moves_dev, int n, bool undo);
	int launch_encode(AgbEngine *e, int n, uint32_t *features_dev);
	int launch_symmetry_f32(AgbEngine *e, const float *src_dev, float *dst_dev, const int8_t *sym_dev, int n, int channels, bool inverse,
			cudaStream_t stream = nullptr);
	int launch_augment(AgbEngine *e, const uint32_t *src_dev, uint32_t *dst_dev, const int8_t *sym_dev, int n, cudaStream_t stream = nullptr);
	int launch_outcomes(AgbEngine *e, const int8_t *boards_dev, const uint16_t *moves_dev, int n, int8_t *out_dev);
	// resnet.cu
	int net_create(AgbEngine *e);
	void net_destroy(AgbEngine *e);
	// selfplay.cu
	int selfplay_create(AgbEngine *e);
	void selfplay_destroy(AgbEngine *e);
}
