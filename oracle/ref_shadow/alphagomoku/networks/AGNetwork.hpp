// TEST INFRASTRUCTURE (oracle build only) -- NOT reference code and NOT product code.
//
// Shadows the reference header include/alphagomoku/networks/AGNetwork.hpp so that the reference's own
// NNEvaluator.cpp / Search.cpp / GameGenerator.cpp / OpeningGenerator.cpp compile unchanged without MinML.
// It keeps the public shape those files use (NNEvaluator.cpp:105-286) and replaces the MinML graph by a
// caller-supplied evaluation callback, so a test can inject any policy/value function (the fp32 restatement
// of the ResNet, or a replay of device outputs) and compare the search bit-for-bit.
#pragma once
#include <alphagomoku/game/Move.hpp>
#include <alphagomoku/networks/NNInputFeatures.hpp>
#include <alphagomoku/patterns/PatternCalculator.hpp>
#include <alphagomoku/search/Value.hpp>
#include <alphagomoku/utils/configs.hpp>
#include <alphagomoku/utils/matrix.hpp>

#include <minml/core/Device.hpp>
#include <minml/core/Event.hpp>
#include <minml/graph/Graph.hpp>

#include <memory>
#include <string>
#include <vector>

extern "C"
{
	// features: uint32 [batch, rows, cols]; policy: f32 [batch, rows*cols]; value: f32 [batch, 3] (win, draw, loss);
	// action_values: f32 [batch, rows*cols, 3] or untouched; moves_left: f32 [batch]
	typedef void (*agref_eval_fn)(void *ctx, const uint32_t *features, int batch, int rows, int cols, float *policy, float *value,
			float *action_values, float *moves_left);
}

namespace ag
{
	class AGNetwork
	{
			GameConfig game_config;
			int batch_size = 1;
			agref_eval_fn eval_fn = nullptr;
			void *eval_ctx = nullptr;
			ml::Graph graph;
			std::unique_ptr<PatternCalculator> calculator;
			NNInputFeatures input_features;
			std::vector<uint32_t> input;
			std::vector<float> policy, value, action_values, moves_left;
		public:
			AGNetwork(const GameConfig &cfg, agref_eval_fn fn, void *ctx);
			void packInputData(int index, const matrix<Sign> &board, Sign signToMove);
			void packInputData(int index, const NNInputFeatures &features);
			void unpackOutput(int index, matrix<float> &policy, matrix<Value> &actionValues, Value &value, float &movesLeft) const;
			void asyncForwardLaunch(int batch_size);
			void asyncForwardJoin();
			void forward(int batch_size);
			void optimize(int level = 1);
			void convertToHalfFloats();
			void unloadGraph();
			bool isLoaded() const noexcept;
			void synchronize();
			void moveTo(ml::Device device);
			int getBatchSize() const noexcept;
			void setBatchSize(int batchSize);
			GameConfig getGameConfig() const noexcept;
			ml::Event addEvent() const;
			ml::Graph& get_graph() { return graph; }
	};
}
