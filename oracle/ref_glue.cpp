// TEST INFRASTRUCTURE (oracle build only) -- NOT reference code and NOT product code.
// Bodies for the shadow AGNetwork (oracle/ref_shadow) and for ag::NetworkLoader, whose reference bodies
// (src/networks/AGNetwork.cpp, src/selfplay/NetworkLoader.cpp) need MinML. Semantics follow
// src/networks/NetworkDataPack.cpp:131-141 (pack) and :200-235 (unpack).
#include <alphagomoku/networks/AGNetwork.hpp>
#include <alphagomoku/selfplay/NetworkLoader.hpp>
#include <alphagomoku/search/alpha_beta/AlphaBetaSearch.hpp>

#include <cstring>
#include <stdexcept>

namespace agref
{
	ag::GameConfig g_game_config;
	agref_eval_fn g_eval_fn = nullptr;
	void *g_eval_ctx = nullptr;
}

namespace ag
{
	AGNetwork::AGNetwork(const GameConfig &cfg, agref_eval_fn fn, void *ctx) :
			game_config(cfg),
			eval_fn(fn),
			eval_ctx(ctx),
			input_features(cfg.rows, cfg.cols)
	{
		setBatchSize(1);
	}
	void AGNetwork::packInputData(int index, const matrix<Sign> &board, Sign signToMove)
	{
		if (calculator == nullptr)
			calculator = std::make_unique<PatternCalculator>(game_config);
		calculator->setBoard(board, signToMove);
		input_features.encode(*calculator);
		packInputData(index, input_features);
	}
	void AGNetwork::packInputData(int index, const NNInputFeatures &features)
	{
		const size_t cells = static_cast<size_t>(game_config.rows) * game_config.cols;
		std::memcpy(input.data() + index * cells, features.data(), cells * sizeof(uint32_t));
	}
	void AGNetwork::unpackOutput(int index, matrix<float> &p, matrix<Value> &actionValues, Value &v, float &movesLeft) const
	{
		const size_t cells = static_cast<size_t>(game_config.rows) * game_config.cols;
		std::memcpy(p.data(), policy.data() + index * cells, cells * sizeof(float));
		v = Value(value[3 * index + 0], value[3 * index + 1]);
		movesLeft = moves_left[index];
		for (size_t i = 0; i < cells; i++)
			actionValues[i] = Value(action_values[(index * cells + i) * 3 + 0], action_values[(index * cells + i) * 3 + 1]);
	}
	void AGNetwork::asyncForwardLaunch(int n)
	{
		forward(n);
	}
	void AGNetwork::asyncForwardJoin()
	{
	}
	void AGNetwork::forward(int n)
	{
		if (eval_fn == nullptr)
			throw std::logic_error("oracle AGNetwork: no evaluation callback installed");
		eval_fn(eval_ctx, input.data(), n, game_config.rows, game_config.cols, policy.data(), value.data(), action_values.data(), moves_left.data());
	}
	void AGNetwork::optimize(int)
	{
	}
	void AGNetwork::convertToHalfFloats()
	{
	}
	void AGNetwork::unloadGraph()
	{
	}
	bool AGNetwork::isLoaded() const noexcept
	{
		return eval_fn != nullptr;
	}
	void AGNetwork::synchronize()
	{
	}
	void AGNetwork::moveTo(ml::Device)
	{
	}
	int AGNetwork::getBatchSize() const noexcept
	{
		return batch_size;
	}
	void AGNetwork::setBatchSize(int batchSize)
	{
		batch_size = batchSize;
		const size_t cells = static_cast<size_t>(game_config.rows) * game_config.cols;
		input.assign(batch_size * cells, 0u);
		policy.assign(batch_size * cells, 0.0f);
		value.assign(batch_size * 3, 0.0f);
		action_values.assign(batch_size * cells * 3, 0.0f);
		moves_left.assign(batch_size, 0.0f);
	}
	GameConfig AGNetwork::getGameConfig() const noexcept
	{
		return game_config;
	}
	ml::Event AGNetwork::addEvent() const
	{
		return ml::Event::now();
	}

	NetworkLoader::NetworkLoader(const char *path) :
			NetworkLoader(std::string(path))
	{
	}
	NetworkLoader::NetworkLoader(const std::string &path) :
			paths( { path })
	{
	}
	NetworkLoader::NetworkLoader(const std::vector<std::string> &path) :
			paths(path)
	{
	}
	std::unique_ptr<AGNetwork> NetworkLoader::get(bool) const
	{
		return std::make_unique<AGNetwork>(agref::g_game_config, agref::g_eval_fn, agref::g_eval_ctx);
	}
}

// NNUE is dead on the self-play path (all call sites in AlphaBetaSearch.cpp are commented out) but AlphaBetaSearch
// still has an InferenceNNUE member; these are the few symbols the linker wants.
#include <alphagomoku/networks/NNUE.hpp>
namespace ag
{
	namespace nnue
	{
		NNUEStats::NNUEStats() :
				refresh("refresh"),
				update("update "),
				forward("forward")
		{
		}
		std::string NNUEStats::toString() const
		{
			return "";
		}
		TrainingNNUE_policy::TrainingNNUE_policy() :
				game_config(GameConfig()),
				calculator(game_config)
		{
		}
		InferenceNNUE::InferenceNNUE(GameConfig gameConfig, const NNUEWeights &) :
				game_config(gameConfig)
		{
		}
		void InferenceNNUE::print_stats() const
		{
		}
	}
}

// Player (src/evaluation/Player.cpp) holds a MovesLeftEstimator, whose body lives in src/player/TimeManager.cpp next to the tournament engine's
// settings (out of scope); the two members Player.cpp needs, restated from TimeManager.cpp:65-76. Only time-controlled searches read it.
#include <alphagomoku/player/TimeManager.hpp>
namespace ag
{
	MovesLeftEstimator::MovesLeftEstimator(const std::vector<std::pair<int, float>> &c0, const std::vector<std::pair<int, float>> &c2) :
			c0(c0, "linear"),
			c2(c2, "linear")
	{
	}
	double MovesLeftEstimator::get(int moveNumber, Value eval) const noexcept
	{
		const double x = std::abs(eval.getExpectation() - 0.5);
		return std::max(1.0, static_cast<double>(c0.getValue(moveNumber)) - static_cast<double>(c2.getValue(moveNumber)) * x * x);
	}
}
