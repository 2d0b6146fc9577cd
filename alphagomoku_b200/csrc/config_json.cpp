// agb_config_from_json: the reference's config.json -> AgbConfig.
//
// The reference reads MasterLearningConfig / SelfplayConfig / SearchConfig / ... through their (const Json&) constructors
// (src/utils/configs.cpp:33-306). This is the same mapping for the fields the device engine honours, with the same key names, the same
// required / optional split (a missing required key is an error there: get_value throws) and the same defaults -- including the one
// place where the JSON default differs from the struct's: EdgeSelectorConfig::init_to is "parent" when the key is absent
// (configs.cpp:71) although a default-constructed EdgeSelectorConfig says "q_head" (configs.hpp:78).
// Host-only code (no CUDA): a small recursive-descent JSON reader, enough for the reference's files (objects, arrays, strings with
// escapes, numbers, true / false / null).
#include "../../include/agb200.h"
#include "json_reader.hpp"

#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace
{
	using agb::json::Value;
	using agb::json::Parser;

	// get_value of configs.cpp:17-28: required (throws) or with a default
	const Value& required(const Value &obj, const char *key, const char *where)
	{
		const Value *v = obj.find(key);
		if (v == nullptr)
			throw std::runtime_error(std::string("missing key \"") + key + "\" in " + where);
		return *v;
	}
	double number_of(const Value &v, const char *key)
	{
		if (v.type != Value::Number)
			throw std::runtime_error(std::string("\"") + key + "\" must be a number");
		return v.number;
	}
	bool bool_of(const Value &v, const char *key)
	{
		if (v.type != Value::Bool)
			throw std::runtime_error(std::string("\"") + key + "\" must be true or false");
		return v.boolean;
	}
	const std::string& string_of(const Value &v, const char *key)
	{
		if (v.type != Value::String)
			throw std::runtime_error(std::string("\"") + key + "\" must be a string");
		return v.string;
	}
	double number_or(const Value &obj, const char *key, double fallback)
	{
		const Value *v = obj.find(key);
		return v == nullptr ? fallback : number_of(*v, key);
	}
	std::string string_or(const Value &obj, const char *key, const char *fallback)
	{
		const Value *v = obj.find(key);
		return v == nullptr ? std::string(fallback) : string_of(*v, key);
	}
	int one_of(const std::string &value, const char *key, std::initializer_list<const char*> names)
	{
		int i = 0;
		for (const char *n : names)
		{
			if (value == n)
				return i;
			i++;
		}
		throw std::runtime_error(std::string("\"") + key + "\": \"" + value + "\" is not supported by the device engine");
	}

	void read_game_config(const Value &game, AgbConfig *c)
	{ // GameConfig(const Json&), configs.cpp:44-50
		c->rules = one_of(string_of(required(game, "rules", "game_config"), "rules"), "rules", { "FREESTYLE", "STANDARD", "RENJU", "CARO5", "CARO6" });
		c->rows = static_cast<int>(number_of(required(game, "rows", "game_config"), "rows"));
		c->cols = static_cast<int>(number_of(required(game, "cols", "game_config"), "cols"));
		c->draw_after = static_cast<int>(number_or(game, "draw_after", static_cast<double>(c->rows) * c->cols));
	}
	int read_selector(const Value &sel, const char *where, bool tree_selector, AgbConfig *c)
	{ // EdgeSelectorConfig(const Json&), configs.cpp:69-76
		const std::string policy = string_or(sel, "policy", "puct");
		if (tree_selector)
		{
			if (policy != "puct")
				throw std::runtime_error(std::string(where) + ": the device engine searches with the \"puct\" selector only (got \"" + policy + "\")");
			c->init_to = one_of(string_or(sel, "init_to", "parent"), "init_to", { "loss", "parent", "draw", "q_head" });
			c->noise_type = one_of(string_or(sel, "noise_type", "none"), "noise_type", { "none", "custom", "dirichlet", "gumbel" });
			c->noise_weight = static_cast<float>(number_or(sel, "noise_weight", 0.0));
			c->exploration_constant = static_cast<float>(number_or(sel, "exploration_constant", 1.25));
			if (number_or(sel, "exploration_scaling", 0.0) != 0.0)
				throw std::runtime_error(std::string(where) + ": exploration_scaling other than 0 is not supported");
			return 0;
		}
		c->final_selector = one_of(policy, "final_selector.policy", { "max_visit", "best", "max_value", "max_policy", "min_visit", "lcb" });
		c->final_exploration_constant = static_cast<float>(number_or(sel, "exploration_constant", 1.25));
		return 0;
	}
	void read_search_config(const Value &search, AgbConfig *c)
	{ // SearchConfig / TreeConfig / MCTSConfig / TSSConfig (const Json&), configs.cpp:52-129
		c->max_batch_size = static_cast<int>(number_or(search, "max_batch_size", 1));
		const Value &tree = required(search, "tree_config", "search_config"), &mcts = required(search, "mcts_config", "search_config"),
				&tss = required(search, "tss_config", "search_config");
		c->information_leak_threshold = static_cast<float>(number_or(tree, "information_leak_threshold", 0.01));
		if (const Value *sel = mcts.find("edge_selector_config"))
			read_selector(*sel, "mcts_config.edge_selector_config", true, c);
		else
		{ // get_value(cfg, "edge_selector_config", EdgeSelectorConfig()): the STRUCT's defaults, where init_to is "q_head" (configs.hpp:78)
			c->init_to = 3;
			c->noise_type = AGB_NOISE_NONE;
			c->noise_weight = 0.0f;
			c->exploration_constant = 1.25f;
		}
		const double max_children = number_or(mcts, "max_children", static_cast<double>(INT_MAX));
		c->max_children = max_children >= static_cast<double>(INT_MAX) ? 0 : static_cast<int>(max_children);
		c->policy_expansion_threshold = static_cast<float>(number_or(mcts, "policy_expansion_threshold", 1.0e-4));
		const double temperature = number_or(mcts, "policy_temperature", 1.0);
		c->policy_temperature = (temperature == 0.0) ? -1.0f : static_cast<float>(temperature); // AgbConfig: 0 means "default", negative the reference's 0
		c->solver_max_positions = static_cast<int>(number_or(tss, "max_positions", 100));
		// tss_config.mode and hash_table_size are parsed by the reference but not used on the self-play path (Search::solve always runs the
		// alpha-beta search with max_positions, Search.cpp:159-183; AlphaBetaSearch sizes its table itself, AlphaBetaSearch.cpp:55)
	}
	void read_selfplay_config(const Value &sp, AgbConfig *c)
	{ // SelfplayConfig(const Json&), configs.cpp:253-268
		bool_of(required(sp, "use_opening", "generation_config"), "use_opening"); // the host decides how games start (agb_generate_openings)
		c->use_symmetries = bool_of(required(sp, "use_symmetries", "generation_config"), "use_symmetries") ? 1 : 0;
		number_of(required(sp, "games_per_iteration", "generation_config"), "games_per_iteration");
		const int games_per_thread = static_cast<int>(number_of(required(sp, "games_per_thread", "generation_config"), "games_per_thread"));
		read_selector(required(sp, "final_selector", "generation_config"), "final_selector", false, c);
		read_search_config(required(sp, "search_config", "generation_config"), c);
		if (const Value *constraints = sp.find("constraints"))
		{ // Constraints(const Json&), configs.cpp:196-212
			if (string_of(required(*constraints, "type", "constraints"), "type") != "simulations")
				throw std::runtime_error("constraints.type \"time\" is not supported: self-play on the device is bounded by simulations");
			c->max_simulations = static_cast<int>(number_or(*constraints, "max_simulations", static_cast<double>(INT_MAX)));
		}
		else
			c->max_simulations = static_cast<int>(number_of(required(sp, "simulations", "generation_config"), "simulations"));
		// concurrency of the reference: one GeneratorThread per device_config entry, games_per_thread games each. The device engine plays them all
		// on one GPU unless the caller overrides `games` afterwards.
		const Value &devices = required(sp, "device_config", "generation_config");
		const int threads = devices.type == Value::Array ? static_cast<int>(devices.array.size()) : 1;
		c->games = games_per_thread * (threads > 0 ? threads : 1);
		c->max_boards = c->games * c->max_batch_size;
	}
}

extern "C" int agb_config_from_json(const char *json_text, AgbConfig *config, char *error, size_t error_size)
{
	if (error != nullptr and error_size > 0)
		error[0] = 0;
	if (json_text == nullptr or config == nullptr)
		return AGB_EINVAL;
	try
	{
		Parser parser(json_text);
		const Value root = parser.parse();
		if (root.type != Value::Object)
			throw std::runtime_error("the configuration must be a JSON object");
		AgbConfig c { };
		read_game_config(required(root, "game_config", "the configuration"), &c);
		// MasterLearningConfig (configs.cpp:290-298) names the self-play part "generation_config"; a bare SelfplayConfig may sit under "selfplay_config"
		const Value *sp = root.find("generation_config");
		if (sp == nullptr)
			sp = root.find("selfplay_config");
		if (sp == nullptr)
			throw std::runtime_error("missing key \"generation_config\" in the configuration");
		read_selfplay_config(*sp, &c);
		if (const Value *training = root.find("training_config"))
		{ // the network the trainer produces: TrainingConfig::network_arch / blocks / filters (configs.cpp:141-156)
			c.blocks = static_cast<int>(number_of(required(*training, "blocks", "training_config"), "blocks"));
			c.filters = static_cast<int>(number_of(required(*training, "filters", "training_config"), "filters"));
			c.q_head = one_of(string_of(required(*training, "network_arch", "training_config"), "network_arch"), "network_arch", { "ResnetPV", "ResnetPVQ" });
		}
		*config = c;
		return AGB_OK;
	}
	catch (std::exception &e)
	{
		if (error != nullptr and error_size > 0)
			std::snprintf(error, error_size, "%s", e.what());
		return AGB_EINVAL;
	}
}
