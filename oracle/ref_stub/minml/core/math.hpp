// TEST INFRASTRUCTURE (oracle build only) -- stand-in for <minml/core/math.hpp>; no-ops (dead branch in EdgeSelector.cpp:818-820).
#pragma once
#include <minml/core/Context.hpp>
#include <minml/core/Tensor.hpp>
namespace ml
{
	enum class ActivationType { LINEAR, SIGMOID, TANH, RELU, SOFTMAX };
	inline void gemm_ex(const Context&, Tensor&, float, char, const Tensor&, char, const Tensor&, float, const Tensor&, const Tensor&, ActivationType) {}
	inline void gemm(const Context&, char, char, Tensor&, const Tensor&, const Tensor&, float, float) {}
}
