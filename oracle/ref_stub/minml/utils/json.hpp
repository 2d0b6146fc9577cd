// TEST INFRASTRUCTURE (oracle build only) -- not product code.
// Minimal stand-in for the un-vendored MinML header <minml/utils/json.hpp>, written from the
// call sites in the reference (src/utils/configs.cpp, src/utils/Parameter.hpp, src/dataset/GameDataBuffer.cpp).
// Only the members those call sites use are provided.
#pragma once
#include <cstdint>
#include <initializer_list>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

enum class JsonType { Null, Bool, Number, String, Array, Object };

class Json
{
		JsonType m_type = JsonType::Null;
		bool m_bool = false;
		double m_number = 0.0;
		bool m_is_integer = false;
		std::string m_string;
		std::vector<Json> m_array;
		std::vector<std::pair<std::string, Json>> m_object; // insertion ordered
	public:
		Json() = default;
		Json(JsonType t) : m_type(t) {}
		Json(bool b) : m_type(JsonType::Bool), m_bool(b) {}
		Json(int i) : m_type(JsonType::Number), m_number(i), m_is_integer(true) {}
		Json(unsigned i) : m_type(JsonType::Number), m_number(i), m_is_integer(true) {}
		Json(long i) : m_type(JsonType::Number), m_number(static_cast<double>(i)), m_is_integer(true) {}
		Json(long long i) : m_type(JsonType::Number), m_number(static_cast<double>(i)), m_is_integer(true) {}
		Json(unsigned long i) : m_type(JsonType::Number), m_number(static_cast<double>(i)), m_is_integer(true) {}
		Json(unsigned long long i) : m_type(JsonType::Number), m_number(static_cast<double>(i)), m_is_integer(true) {}
		Json(float f) : m_type(JsonType::Number), m_number(f) {}
		Json(double d) : m_type(JsonType::Number), m_number(d) {}
		Json(const char *s) : m_type(JsonType::String), m_string(s) {}
		Json(const std::string &s) : m_type(JsonType::String), m_string(s) {}
		Json(std::initializer_list<std::pair<std::string, Json>> list) : m_type(JsonType::Object)
		{
			for (const auto &kv : list)
				(*this)[kv.first] = kv.second;
		}

		bool isNull() const { return m_type == JsonType::Null; }
		bool isBool() const { return m_type == JsonType::Bool; }
		bool isNumber() const { return m_type == JsonType::Number; }
		bool isString() const { return m_type == JsonType::String; }
		bool isArray() const { return m_type == JsonType::Array; }
		bool isObject() const { return m_type == JsonType::Object; }

		bool getBool() const { check(JsonType::Bool); return m_bool; }
		int getInt() const { check(JsonType::Number); return static_cast<int>(m_number); }
		int64_t getLong() const { check(JsonType::Number); return static_cast<int64_t>(m_number); }
		double getDouble() const { check(JsonType::Number); return m_number; }
		const std::string& getString() const { check(JsonType::String); return m_string; }
		operator bool() const { return getBool(); }
		operator int() const { return getInt(); }
		operator int64_t() const { return getLong(); }
		operator float() const { return static_cast<float>(getDouble()); }
		operator double() const { return getDouble(); }
		operator std::string() const { return getString(); }

		int size() const
		{
			if (m_type == JsonType::Array) return static_cast<int>(m_array.size());
			if (m_type == JsonType::Object) return static_cast<int>(m_object.size());
			return 0;
		}
		bool hasKey(const std::string &key) const
		{
			for (const auto &kv : m_object)
				if (kv.first == key) return true;
			return false;
		}
		Json& operator[](const std::string &key)
		{
			if (m_type == JsonType::Null) m_type = JsonType::Object;
			check(JsonType::Object);
			for (auto &kv : m_object)
				if (kv.first == key) return kv.second;
			m_object.emplace_back(key, Json());
			return m_object.back().second;
		}
		const Json& operator[](const std::string &key) const
		{
			check(JsonType::Object);
			for (const auto &kv : m_object)
				if (kv.first == key) return kv.second;
			throw std::out_of_range("Json: no key '" + key + "'");
		}
		Json& operator[](const char *key) { return (*this)[std::string(key)]; }
		const Json& operator[](const char *key) const { return (*this)[std::string(key)]; }
		Json& operator[](int idx)
		{
			if (m_type == JsonType::Null) m_type = JsonType::Array;
			check(JsonType::Array);
			if (idx >= static_cast<int>(m_array.size())) m_array.resize(idx + 1);
			return m_array[idx];
		}
		const Json& operator[](int idx) const { check(JsonType::Array); return m_array.at(idx); }
		Json& operator[](size_t idx) { return (*this)[static_cast<int>(idx)]; }
		const Json& operator[](size_t idx) const { return (*this)[static_cast<int>(idx)]; }

		std::string dump(int indent = -1) const;
		static Json load(const std::string &str);
	private:
		void check(JsonType t) const
		{
			if (m_type != t) throw std::logic_error("Json: wrong type access");
		}
		void dump_impl(std::string &out, int indent, int level) const;
};
