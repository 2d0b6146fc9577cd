// TEST INFRASTRUCTURE (oracle build only) -- stand-in for <minml/utils/serialization.hpp>.
// A growable byte buffer with POD save/load, as used by src/dataset/*.cpp and src/utils/file_util.cpp.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <type_traits>
#include <vector>

class SerializedObject
{
		std::vector<uint8_t> m_data;
	public:
		SerializedObject() = default;
		size_t size() const noexcept { return m_data.size(); }
		size_t capacity() const noexcept { return m_data.capacity(); }
		void clear() noexcept { m_data.clear(); }
		const uint8_t* data() const noexcept { return m_data.data(); }
		uint8_t* data() noexcept { return m_data.data(); }
		void save(const void *src, size_t bytes)
		{
			const uint8_t *p = static_cast<const uint8_t*>(src);
			m_data.insert(m_data.end(), p, p + bytes);
		}
		void load(void *dst, size_t offset, size_t bytes) const
		{
			if (offset + bytes > m_data.size()) throw std::out_of_range("SerializedObject::load");
			std::memcpy(dst, m_data.data() + offset, bytes);
		}
		template<typename T>
		void save(const T &value)
		{
			static_assert(std::is_trivially_copyable<T>::value, "");
			save(&value, sizeof(T));
		}
		template<typename T>
		T load(size_t offset) const
		{ // does not advance: every reference call site adds sizeof(T) itself (e.g. SearchDataStorage.cpp:298-316)
			static_assert(std::is_trivially_copyable<T>::value, "");
			T result;
			load(&result, offset, sizeof(T));
			return result;
		}
};
