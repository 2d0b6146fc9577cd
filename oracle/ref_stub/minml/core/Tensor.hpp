// TEST INFRASTRUCTURE (oracle build only) -- opaque stand-in for <minml/core/Tensor.hpp>.
// Only what src/search/monte_carlo/EdgeSelector.cpp (LearnablePolicySelector, dead in self-play) needs to compile.
#pragma once
#include <minml/core/Device.hpp>
#include <initializer_list>
#include <string>
#include <vector>
namespace ml
{
	enum class DataType { UNKNOWN, FLOAT16, FLOAT32, INT32 };
	class Shape
	{
			std::vector<int> m_dims;
		public:
			Shape() = default;
			Shape(std::initializer_list<int> d) : m_dims(d) {}
			int rank() const { return static_cast<int>(m_dims.size()); }
			int operator[](int i) const { return m_dims.at(i); }
			int volume() const { int v = 1; for (int d : m_dims) v *= d; return v; }
			int lastDim() const { return m_dims.back(); }
			int firstDim() const { return m_dims.front(); }
	};
	class Tensor
	{
			Shape m_shape;
			std::vector<float> m_data;
		public:
			Tensor() = default;
			Tensor(const Shape &s, const std::string& = "float32", Device = Device::cpu()) : m_shape(s), m_data(s.volume()) {}
			Tensor(const Shape &s, DataType, Device = Device::cpu()) : m_shape(s), m_data(s.volume()) {}
			Tensor(std::initializer_list<int> d) : m_shape(d), m_data(m_shape.volume()) {}
			const Shape& shape() const { return m_shape; }
			int dim(int i) const { return m_shape[i]; }
			int volume() const { return m_shape.volume(); }
			void* data() { return m_data.data(); }
			const void* data() const { return m_data.data(); }
			float get(std::initializer_list<int>) const { return 0.0f; }
			void set(float, std::initializer_list<int>) {}
			void zeroall() {}
			Tensor view(std::initializer_list<int> d) const { return Tensor(Shape(d)); }
			Tensor flatten(std::initializer_list<int>) const { return *this; }
			int lastDim() const { return m_shape.lastDim(); }
			int firstDim() const { return m_shape.firstDim(); }
	};
}
