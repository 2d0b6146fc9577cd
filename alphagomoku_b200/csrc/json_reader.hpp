// A small recursive-descent JSON reader, enough for the reference's files (config.json, the headers of GameDataBuffer files): objects,
// arrays, strings with escapes, numbers, true / false / null. Host-only.
#pragma once
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace agb
{
	namespace json
	{
	struct Value
		{
				enum Type { Null, Bool, Number, String, Array, Object } type = Null;
				bool boolean = false;
				double number = 0.0;
				std::string string;
				std::vector<Value> array;
				std::vector<std::pair<std::string, Value>> object;

				const Value* find(const std::string &key) const
				{
					if (type != Object)
						return nullptr;
					for (const auto &kv : object)
						if (kv.first == key)
							return &kv.second;
					return nullptr;
				}
		};
		struct Parser
		{
				const char *p, *end;
				explicit Parser(const char *text) : p(text), end(text + std::strlen(text)) {}
				void skip()
				{
					while (p < end and (*p == ' ' or *p == '\t' or *p == '\n' or *p == '\r'))
						p++;
				}
				Value parse()
				{
					skip();
					if (p >= end)
						throw std::runtime_error("JSON: unexpected end of text");
					Value v;
					if (*p == '{')
					{
						v.type = Value::Object;
						p++;
						skip();
						if (p < end and *p == '}')
						{
							p++;
							return v;
						}
						for (;;)
						{
							skip();
							if (p >= end or *p != '"')
								throw std::runtime_error("JSON: expected a key string");
							std::string key = parse_string();
							skip();
							if (p >= end or *p != ':')
								throw std::runtime_error("JSON: expected ':' after key \"" + key + "\"");
							p++;
							v.object.emplace_back(key, parse());
							skip();
							if (p < end and *p == ',')
							{
								p++;
								continue;
							}
							if (p < end and *p == '}')
							{
								p++;
								return v;
							}
							throw std::runtime_error("JSON: expected ',' or '}' after the value of \"" + key + "\"");
						}
					}
					if (*p == '[')
					{
						v.type = Value::Array;
						p++;
						skip();
						if (p < end and *p == ']')
						{
							p++;
							return v;
						}
						for (;;)
						{
							v.array.push_back(parse());
							skip();
							if (p < end and *p == ',')
							{
								p++;
								continue;
							}
							if (p < end and *p == ']')
							{
								p++;
								return v;
							}
							throw std::runtime_error("JSON: expected ',' or ']' in an array");
						}
					}
					if (*p == '"')
					{
						v.type = Value::String;
						v.string = parse_string();
						return v;
					}
					if (std::strncmp(p, "true", 4) == 0)
					{
						p += 4;
						v.type = Value::Bool;
						v.boolean = true;
						return v;
					}
					if (std::strncmp(p, "false", 5) == 0)
					{
						p += 5;
						v.type = Value::Bool;
						return v;
					}
					if (std::strncmp(p, "null", 4) == 0)
					{
						p += 4;
						return v;
					}
					char *after = nullptr;
					v.number = std::strtod(p, &after);
					if (after == p)
						throw std::runtime_error(std::string("JSON: unexpected character '") + *p + "'");
					p = after;
					v.type = Value::Number;
					return v;
				}
				std::string parse_string()
				{
					std::string out;
					p++; // opening quote
					while (p < end and *p != '"')
					{
						if (*p == '\\' and p + 1 < end)
						{
							p++;
							switch (*p)
							{
								case 'n': out += '\n'; break;
								case 't': out += '\t'; break;
								case 'r': out += '\r'; break;
								case 'b': out += '\b'; break;
								case 'f': out += '\f'; break;
								case 'u': // the reference's files are ASCII: keep the escape's low byte
									if (p + 4 < end)
									{
										out += static_cast<char>(std::strtol(std::string(p + 1, p + 5).c_str(), nullptr, 16) & 0x7F);
										p += 4;
									}
									break;
								default: out += *p; break;
							}
							p++;
						}
						else
							out += *p++;
					}
					if (p >= end)
						throw std::runtime_error("JSON: unterminated string");
					p++;
					return out;
				}
		};

	}
}
