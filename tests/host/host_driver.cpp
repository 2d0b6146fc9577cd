// TEST INFRASTRUCTURE -- runs the reference's own host code on top of the B200 boundary, on the GPU box.
//
//   agb_host_b200 generator <weights.f32> <out buffer> <games> : the UNMODIFIED reference GeneratorManager / GeneratorThread::run /
//        GameGenerator / Search / Tree / AlphaBetaSearch (from oracle/_ref/libagref.so) with integration/NNEvaluator_b200.cpp linked in
//        place of the reference's NNEvaluator.cpp: every NNEvaluator call of the reference -- addToQueue, isQueueFull,
//        asyncEvaluateGraphLaunch / Join, useSymmetries, getStats, loadGraph / unloadGraph -- lands in libagb200.so.
//   agb_host_shadow generator ...                              : the same program with the reference's NNEvaluator.cpp and the oracle's
//        shadow AGNetwork, whose forward is agb_forward on the features the REFERENCE computed. Same seeds => the two programs must write
//        byte-identical game buffers (tests/test_host_gpu.py), which pins the whole evaluator drop-in incl. the double-buffered schedule.
//   agb_host_b200 device <weights.f32> <out dir> <games>       : GeneratorManagerB200 (integration/agb200_shims.hpp) -- the device lockstep
//        engine behind GeneratorManager's surface: generate, getGameBuffer, saveState, loadState.
#include <alphagomoku/selfplay/GeneratorManager.hpp>
#include <alphagomoku/selfplay/GameGenerator.hpp>
#include <alphagomoku/selfplay/NetworkLoader.hpp>
#include <alphagomoku/networks/AGNetwork.hpp>
#include <alphagomoku/utils/configs.hpp>

#include <agb200.h>
#include "../../integration/agb200_shims.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <sys/stat.h>

using namespace ag;

namespace agb200
{
	void register_network(const ag::GameConfig &game, int blocks, int filters, bool q_head, const float *blob, size_t count, int device);
}
#ifdef AGB_HOST_SHADOW
namespace agref
{
	extern GameConfig g_game_config;
	extern agref_eval_fn g_eval_fn;
	extern void *g_eval_ctx;
}
namespace
{
	AgbEngine *g_forward_engine = nullptr;
	void device_forward(void*, const uint32_t *features, int batch, int rows, int cols, float *policy, float *value, float *action_values, float *moves_left)
	{ // the shadow network's forward: the device network on the reference's own feature words
		if (agb_forward(g_forward_engine, features, batch, policy, value, nullptr) != AGB_OK)
		{
			std::fprintf(stderr, "agb_forward: %s\n", agb_last_error(g_forward_engine));
			std::exit(3);
		}
		std::memset(action_values, 0, sizeof(float) * batch * rows * cols * 3);
		std::memset(moves_left, 0, sizeof(float) * batch);
	}
}
#endif

namespace
{
	constexpr int kBlocks = 4, kFilters = 64, kSize = 15;
	std::vector<float> read_floats(const char *path)
	{
		FILE *f = std::fopen(path, "rb");
		if (f == nullptr)
		{
			std::fprintf(stderr, "cannot read %s\n", path);
			std::exit(2);
		}
		std::vector<float> out;
		float chunk[16384];
		size_t n;
		while ((n = std::fread(chunk, sizeof(float), 16384, f)) > 0)
			out.insert(out.end(), chunk, chunk + n);
		std::fclose(f);
		return out;
	}
	SelfplayConfig selfplay_config()
	{
		SelfplayConfig sc;
		sc.use_opening = true;
		sc.use_symmetries = true;
		sc.games_per_thread = 8;
		sc.constraints.max_simulations = 100;
		sc.final_selector.policy = "max_visit";
		sc.device_config = { DeviceConfig() };
		sc.device_config[0].batch_size = 64;
		sc.search_config.max_batch_size = 8;
		sc.search_config.mcts_config.edge_selector_config.policy = "puct";
		sc.search_config.mcts_config.edge_selector_config.init_to = "parent";
		sc.search_config.mcts_config.edge_selector_config.exploration_constant = 1.25f;
		sc.search_config.tss_config.max_positions = 100;
		return sc;
	}
}

int main(int argc, char **argv)
{
	if (argc < 5)
	{
		std::fprintf(stderr, "usage: %s generator|device <weights.f32> <out> <games>\n", argv[0]);
		return 2;
	}
	const std::string mode = argv[1];
	const std::vector<float> blob = read_floats(argv[2]);
	const int games = std::atoi(argv[4]);
	const GameConfig gc(GameRules::STANDARD, kSize, kSize);
	try
	{
		if (mode == "generator")
		{
#ifdef AGB_HOST_SHADOW
			AgbConfig c { };
			c.rules = static_cast<int>(gc.rules);
			c.rows = c.cols = kSize;
			c.max_boards = 64;
			c.blocks = kBlocks;
			c.filters = kFilters;
			if (agb_create(&c, &g_forward_engine) != AGB_OK or agb_load_weights(g_forward_engine, blob.data(), blob.size() * sizeof(float)) != AGB_OK)
			{
				std::fprintf(stderr, "engine: %s\n", agb_last_error(g_forward_engine));
				return 3;
			}
			agref::g_game_config = gc;
			agref::g_eval_fn = device_forward;
			agref::g_eval_ctx = nullptr;
#else
			agb200::register_network(gc, kBlocks, kFilters, false, blob.data(), blob.size(), 0);
#endif
			GeneratorManager manager(gc, selfplay_config());
			manager.generate(NetworkLoader(""), games); // the reference's own loop: GeneratorThread::run on a worker thread until enough games
			manager.getGameBuffer().save(argv[3]);
			std::printf("games %d samples %d\n", manager.getGameBuffer().numberOfGames(), manager.getGameBuffer().numberOfSamples());
			return 0;
		}
#ifndef AGB_HOST_SHADOW
		if (mode == "device")
		{
			const std::string dir = argv[3];
			mkdir(dir.c_str(), 0755);
			mkdir((dir + "/saved_state").c_str(), 0755);
			AgbConfig c { };
			c.rules = static_cast<int>(gc.rules);
			c.rows = c.cols = kSize;
			c.blocks = kBlocks;
			c.filters = kFilters;
			c.games = 64;
			c.max_batch_size = 8;
			c.max_boards = c.games * c.max_batch_size;
			c.max_simulations = 100;
			c.init_to = 1;
			c.exploration_constant = 1.25f;
			c.information_leak_threshold = 0.01f;
			c.policy_expansion_threshold = 1.0e-4f;
			c.solver_max_positions = 100;
			c.use_symmetries = 1;
			c.seed = 7;
			{
				GeneratorManagerB200 manager(gc, c, blob.data(), blob.size() * sizeof(float));
				manager.setWorkingDirectory(dir);
				manager.resetGames();
				manager.generate(games / 2);
				manager.saveState(true); // buffer.bin + the games in flight
				std::printf("first half: games %d samples %d\n", manager.getGameBuffer().numberOfGames(), manager.getGameBuffer().numberOfSamples());
			}
			GeneratorManagerB200 resumed(gc, c, blob.data(), blob.size() * sizeof(float));
			resumed.setWorkingDirectory(dir);
			resumed.resetGames();
			resumed.loadState();
			const int loaded = resumed.getGameBuffer().numberOfGames();
			resumed.generate(games);
			resumed.getGameBuffer().save(dir + "/buffer_final.bin");
			std::printf("resumed with %d games, finished with games %d samples %d\n", loaded, resumed.getGameBuffer().numberOfGames(),
					resumed.getGameBuffer().numberOfSamples());
			return 0;
		}
#endif
		std::fprintf(stderr, "unknown mode %s\n", mode.c_str());
		return 2;
	}
	catch (std::exception &e)
	{
		std::fprintf(stderr, "host driver: %s\n", e.what());
		return 1;
	}
}
