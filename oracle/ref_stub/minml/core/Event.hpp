// TEST INFRASTRUCTURE (oracle build only) -- stand-in for <minml/core/Event.hpp>; wall-clock events.
#pragma once
#include <chrono>
namespace ml
{
	class Event
	{
			double m_time = -1.0;
		public:
			Event() = default;
			static Event now()
			{
				Event e;
				e.m_time = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
				return e;
			}
			bool isReady() const noexcept { return m_time >= 0.0; }
			void synchronize() const noexcept {}
			static double getElapsedTime(const Event &start, const Event &end) noexcept { return end.m_time - start.m_time; }
	};
}
