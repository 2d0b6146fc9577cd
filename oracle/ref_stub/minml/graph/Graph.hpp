// TEST INFRASTRUCTURE (oracle build only) -- opaque stand-in for <minml/graph/Graph.hpp>.
// Only the members named by src/search/monte_carlo/EdgeSelector.cpp:626-645 (LearnablePolicySelector weights; never
// constructed on the self-play path) exist, and they throw if reached.
#pragma once
#include <minml/core/Context.hpp>
#include <minml/core/Tensor.hpp>
#include <minml/core/math.hpp>
#include <minml/utils/json.hpp>
#include <minml/utils/serialization.hpp>
#include <stdexcept>
namespace ml
{
	class Parameter
	{
			Tensor m_param;
		public:
			Tensor& getParam() { return m_param; }
	};
	class Layer
	{
			Parameter m_w, m_b;
		public:
			Parameter& getWeights() { return m_w; }
			Parameter& getBias() { return m_b; }
	};
	class GraphNode
	{
			Layer m_layer;
		public:
			Layer& getLayer() { return m_layer; }
	};
	class Graph
	{
			GraphNode m_node;
		public:
			Graph() = default;
			void load(const Json&, const SerializedObject&) { throw std::logic_error("ml::Graph stub: load() is not available in the oracle build"); }
			int numberOfNodes() const { return 0; }
			GraphNode& getNode(int) { return m_node; }
	};
}
