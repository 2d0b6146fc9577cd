// TEST INFRASTRUCTURE (oracle build only) -- stand-in for <minml/core/Device.hpp>; CPU device only.
#pragma once
#include <string>
namespace ml
{
	enum class DeviceType { CPU, CUDA, OPENCL };
	class Device
	{
			DeviceType m_type = DeviceType::CPU;
			int m_index = 0;
		public:
			Device() = default;
			static Device cpu() { return Device(); }
			static Device cuda(int) { return Device(); }
			static Device fromString(const std::string &) { return Device(); }
			std::string toString() const { return "CPU"; }
			bool isCPU() const { return true; }
			bool isCUDA() const { return false; }
			bool isOPENCL() const { return false; }
			DeviceType type() const { return m_type; }
			int index() const { return m_index; }
			std::string info() const { return "stub CPU"; }
			static void setNumberOfThreads(int) {}
			static int numberOfCudaDevices() { return 0; }
			static int numberOfOpenCLDevices() { return 0; }
			static int cpuCores() { return 1; }
			static std::string hardwareInfo() { return "stub"; }
			friend bool operator==(const Device &a, const Device &b) { return a.m_type == b.m_type and a.m_index == b.m_index; }
			friend bool operator!=(const Device &a, const Device &b) { return not (a == b); }
	};
}
