// TEST INFRASTRUCTURE (oracle build only) -- opaque stand-in for <minml/core/Context.hpp>.
#pragma once
#include <minml/core/Device.hpp>
namespace ml
{
	class Context
	{
		public:
			Context() = default;
			Context(Device) {}
			Device device() const { return Device::cpu(); }
			void synchronize() const {}
	};
}
