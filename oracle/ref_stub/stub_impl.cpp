// TEST INFRASTRUCTURE (oracle build only) -- bodies for the MinML stand-ins in oracle/ref_stub/minml.
#include <minml/utils/json.hpp>
#include <minml/utils/ZipWrapper.hpp>

#include <zlib.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace
{
	void escape_into(std::string &out, const std::string &s)
	{
		out += '"';
		for (char c : s)
		{
			switch (c)
			{
				case '"': out += "\\\""; break;
				case '\\': out += "\\\\"; break;
				case '\n': out += "\\n"; break;
				case '\t': out += "\\t"; break;
				case '\r': out += "\\r"; break;
				default: out += c;
			}
		}
		out += '"';
	}
	struct Parser
	{
			const std::string &s;
			size_t p = 0;
			explicit Parser(const std::string &str) : s(str) {}
			void ws() { while (p < s.size() and (s[p] == ' ' or s[p] == '\n' or s[p] == '\t' or s[p] == '\r')) p++; }
			std::string str()
			{
				std::string r;
				p++; // opening quote
				while (p < s.size() and s[p] != '"')
				{
					if (s[p] == '\\' and p + 1 < s.size())
					{
						p++;
						switch (s[p])
						{
							case 'n': r += '\n'; break;
							case 't': r += '\t'; break;
							case 'r': r += '\r'; break;
							default: r += s[p];
						}
					}
					else
						r += s[p];
					p++;
				}
				p++; // closing quote
				return r;
			}
			Json value()
			{
				ws();
				if (p >= s.size()) return Json();
				const char c = s[p];
				if (c == '{')
				{
					Json r(JsonType::Object);
					p++; ws();
					if (s[p] == '}') { p++; return r; }
					while (true)
					{
						ws();
						const std::string key = str();
						ws(); p++; // ':'
						r[key] = value();
						ws();
						if (s[p] == ',') { p++; continue; }
						p++; // '}'
						break;
					}
					return r;
				}
				if (c == '[')
				{
					Json r(JsonType::Array);
					p++; ws();
					if (s[p] == ']') { p++; return r; }
					int i = 0;
					while (true)
					{
						r[i++] = value();
						ws();
						if (s[p] == ',') { p++; continue; }
						p++; // ']'
						break;
					}
					return r;
				}
				if (c == '"') return Json(str());
				if (s.compare(p, 4, "true") == 0) { p += 4; return Json(true); }
				if (s.compare(p, 5, "false") == 0) { p += 5; return Json(false); }
				if (s.compare(p, 4, "null") == 0) { p += 4; return Json(); }
				const size_t start = p;
				bool is_int = true;
				while (p < s.size() and (std::isdigit(static_cast<unsigned char>(s[p])) or s[p] == '-' or s[p] == '+' or s[p] == '.' or s[p] == 'e' or s[p] == 'E'))
				{
					if (s[p] == '.' or s[p] == 'e' or s[p] == 'E') is_int = false;
					p++;
				}
				const std::string num = s.substr(start, p - start);
				if (is_int) return Json(static_cast<long long>(std::strtoll(num.c_str(), nullptr, 10)));
				return Json(std::strtod(num.c_str(), nullptr));
			}
	};
}

void Json::dump_impl(std::string &out, int indent, int level) const
{
	const auto newline = [&](int lvl)
	{
		if (indent >= 0)
		{
			out += '\n';
			out.append(static_cast<size_t>(indent * lvl), ' ');
		}
	};
	switch (m_type)
	{
		case JsonType::Null: out += "null"; break;
		case JsonType::Bool: out += m_bool ? "true" : "false"; break;
		case JsonType::Number:
		{
			char buf[64];
			if (m_is_integer or (std::floor(m_number) == m_number and std::fabs(m_number) < 1e15))
				std::snprintf(buf, sizeof(buf), m_is_integer ? "%lld" : "%lld.0", static_cast<long long>(m_number));
			else
				std::snprintf(buf, sizeof(buf), "%.9g", m_number);
			out += buf;
			break;
		}
		case JsonType::String: escape_into(out, m_string); break;
		case JsonType::Array:
			out += '[';
			for (size_t i = 0; i < m_array.size(); i++)
			{
				if (i) out += (indent >= 0) ? ", " : ",";
				m_array[i].dump_impl(out, -1, level + 1);
			}
			out += ']';
			break;
		case JsonType::Object:
			out += '{';
			for (size_t i = 0; i < m_object.size(); i++)
			{
				if (i) out += ',';
				newline(level + 1);
				escape_into(out, m_object[i].first);
				out += (indent >= 0) ? ": " : ":";
				m_object[i].second.dump_impl(out, indent, level + 1);
			}
			if (not m_object.empty()) newline(level);
			out += '}';
			break;
	}
}
std::string Json::dump(int indent) const
{
	std::string out;
	dump_impl(out, indent, 0);
	return out;
}
Json Json::load(const std::string &str)
{
	Parser parser(str);
	return parser.value();
}

std::vector<char> ZipWrapper::compress(const std::vector<char> &data, int level)
{
	uLongf bound = compressBound(data.size());
	std::vector<char> out(bound);
	if (compress2(reinterpret_cast<Bytef*>(out.data()), &bound, reinterpret_cast<const Bytef*>(data.data()), data.size(), level) != Z_OK)
		throw std::runtime_error("zlib compress failed");
	out.resize(bound);
	return out;
}
std::vector<char> ZipWrapper::uncompress(const std::vector<char> &data)
{
	std::vector<char> out(std::max<size_t>(1024, data.size() * 4));
	while (true)
	{
		uLongf len = out.size();
		const int rc = ::uncompress(reinterpret_cast<Bytef*>(out.data()), &len, reinterpret_cast<const Bytef*>(data.data()), data.size());
		if (rc == Z_OK) { out.resize(len); return out; }
		if (rc != Z_BUF_ERROR) throw std::runtime_error("zlib uncompress failed");
		out.resize(out.size() * 2);
	}
}
