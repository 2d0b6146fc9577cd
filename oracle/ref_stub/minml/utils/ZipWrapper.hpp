// TEST INFRASTRUCTURE (oracle build only) -- stand-in for <minml/utils/ZipWrapper.hpp> over system zlib.
#pragma once
#include <vector>
class ZipWrapper
{
	public:
		static std::vector<char> compress(const std::vector<char> &data, int level = -1);
		static std::vector<char> uncompress(const std::vector<char> &data);
};
