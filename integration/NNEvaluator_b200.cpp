// Link-time drop-in for the reference's src/search/monte_carlo/NNEvaluator.cpp.
//
// The reference has no virtual evaluator interface: Search::scheduleToNN, GameGenerator and GeneratorThread hold the concrete class
// ag::NNEvaluator (include/alphagomoku/search/monte_carlo/NNEvaluator.hpp:42-83). A drop-in therefore has to BE that class: this file
// defines every member function the header declares (and NNEvaluatorStats), with the same queue semantics, random draws, statistics and
// exceptions, but the evaluation itself -- pack_to_network -> AGNetwork::forward -> unpack_from_network -- is one call of the B200 engine's C ABI
// (agb_evaluate_features: K4 + inverse symmetry for tasks that carry their feature words, agb_evaluate: K1 + K3 + augment + K4 + inverse symmetry for the others). A maintainer compiles this file INSTEAD of NNEvaluator.cpp
// and links libagb200.so; nothing else in the reference changes. tests/host builds exactly that and runs the
// reference's own GeneratorThread::run loop on it.
//
// Where the weights come from: the reference's NetworkLoader hands out MinML graphs, whose file format is not in its tree. Until a converter
// exists the network is registered once per process as the engine's fp32 blob (agb200::register_network, the role NetworkLoader's path plays);
// loadGraph(loader) then ignores the loader's file.
#include <alphagomoku/search/monte_carlo/NNEvaluator.hpp>
#include <alphagomoku/search/monte_carlo/SearchTask.hpp>
#include <alphagomoku/selfplay/NetworkLoader.hpp>
#include <alphagomoku/utils/augmentations.hpp>
#include <alphagomoku/utils/misc.hpp>
#include <alphagomoku/utils/random.hpp>

#include <agb200.h>

#include <algorithm>
#include <cstring>
#include <future>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace agb200
{
	struct NetworkDescription
	{
			ag::GameConfig game;
			int blocks = 0, filters = 0;
			bool q_head = false;
			std::vector<float> blob;
			int device = 0;
	};
	namespace
	{
		std::mutex g_mutex;
		NetworkDescription g_network;
		// state the reference's class has no member for, keyed by the evaluator
		struct Engine
		{
				AgbEngine *handle = nullptr;
				int cells = 0;
				bool q_head = false;
				// a batch in two parts, like NNEvaluator::pack_to_network's two branches: tasks the solver has processed carry their feature words
				// (results [0, n_features)), the others are sent as boards and encoded on the device (results [n_features, n))
				std::vector<uint32_t> features;
				std::vector<int8_t> boards, sign_to_move, symmetry_features, symmetry_boards;
				std::vector<int> result_of; // batch index -> row of policy / value / q
				int n_features = 0, n_boards = 0;
				std::vector<float> policy, value, q;
				int evaluate()
				{ // pack_to_network -> forward -> unpack_from_network's device half, for both parts
					int rc = AGB_OK;
					if (n_features > 0)
						rc = agb_evaluate_features(handle, features.data(), symmetry_features.data(), n_features, policy.data(), value.data(), q_head ? q.data() : nullptr);
					if (rc == AGB_OK and n_boards > 0)
						rc = agb_evaluate(handle, boards.data(), sign_to_move.data(), symmetry_boards.data(), n_boards, policy.data() + static_cast<size_t>(n_features) * cells,
								value.data() + static_cast<size_t>(n_features) * 3, q_head ? q.data() + static_cast<size_t>(n_features) * cells * 3 : nullptr);
					return rc;
				}
				std::future<int> pending;
				double time_per_sample = 1.0e-4; // PerfEstimator's role: seconds per position, refreshed after every batch
				double launch_time = 0.0;
				~Engine()
				{
					if (pending.valid())
						pending.wait();
					agb_destroy(handle);
				}
		};
		std::unordered_map<const ag::NNEvaluator*, std::unique_ptr<Engine>> g_engines;
		Engine* find(const ag::NNEvaluator *self)
		{
			std::lock_guard<std::mutex> lock(g_mutex);
			const auto it = g_engines.find(self);
			return it == g_engines.end() ? nullptr : it->second.get();
		}
	}
	// the network every evaluator of this process loads (fp32 blob of alphagomoku_b200/netblob.py, BN folded)
	void register_network(const ag::GameConfig &game, int blocks, int filters, bool q_head, const float *blob, size_t count, int device)
	{
		std::lock_guard<std::mutex> lock(g_mutex);
		g_network.game = game;
		g_network.blocks = blocks;
		g_network.filters = filters;
		g_network.q_head = q_head;
		g_network.blob.assign(blob, blob + count);
		g_network.device = device;
	}
}

namespace ag
{
	NNEvaluatorStats::NNEvaluatorStats() :
			pack("pack   "),
			compute("compute"),
			unpack("unpack ")
	{
	}
	std::string NNEvaluatorStats::toString() const
	{
		std::string result = "----NNEvaluator (B200)----\n";
		result += "total samples = " + std::to_string(batch_sizes) + '\n';
		result += "avg batch size = " + std::to_string(static_cast<double>(batch_sizes) / compute.getTotalCount()) + '\n';
		result += pack.toString() + '\n';
		result += compute.toString() + '\n';
		result += unpack.toString() + '\n';
		return result;
	}
	NNEvaluatorStats& NNEvaluatorStats::operator+=(const NNEvaluatorStats &other) noexcept
	{
		this->batch_sizes += other.batch_sizes;
		this->pack += other.pack;
		this->compute += other.compute;
		this->unpack += other.unpack;
		return *this;
	}
	NNEvaluatorStats& NNEvaluatorStats::operator/=(int i) noexcept
	{
		this->batch_sizes /= i;
		this->pack /= i;
		this->compute /= i;
		this->unpack /= i;
		return *this;
	}

	NNEvaluator::NNEvaluator(const DeviceConfig &cfg) :
			config(cfg)
	{
	}
	bool NNEvaluator::isOnGPU() const noexcept
	{
		return true;
	}
	void NNEvaluator::clearStats() noexcept
	{
		stats = NNEvaluatorStats();
	}
	NNEvaluatorStats NNEvaluator::getStats() const noexcept
	{
		return stats;
	}
	bool NNEvaluator::isQueueFull() const noexcept
	{
		return getQueueSize() >= config.batch_size;
	}
	int NNEvaluator::getQueueSize() const noexcept
	{
		return waiting_queue.size();
	}
	void NNEvaluator::clearQueue() noexcept
	{
		waiting_queue.clear();
	}
	void NNEvaluator::useSymmetries(bool b) noexcept
	{
		use_symmetries = b;
	}
	void NNEvaluator::loadGraph(const NetworkLoader &loader)
	{
		(void) loader; // see the note on weights at the top of this file
		std::unique_ptr<agb200::Engine> engine = std::make_unique<agb200::Engine>();
		std::lock_guard<std::mutex> lock(agb200::g_mutex);
		const agb200::NetworkDescription &net = agb200::g_network;
		if (net.blob.empty())
			throw std::logic_error("NNEvaluator::loadGraph() : no network registered (agb200::register_network)");
		AgbConfig c { };
		c.rules = static_cast<int>(net.game.rules);
		c.rows = net.game.rows;
		c.cols = net.game.cols;
		c.draw_after = net.game.draw_after;
		c.device = net.device;
		c.max_boards = config.batch_size;
		c.blocks = net.blocks;
		c.filters = net.filters;
		c.q_head = net.q_head;
		if (agb_create(&c, &engine->handle) != AGB_OK)
			throw std::runtime_error(std::string("NNEvaluator::loadGraph() : ") + agb_last_error(nullptr));
		if (agb_load_weights(engine->handle, net.blob.data(), net.blob.size() * sizeof(float)) != AGB_OK)
			throw std::runtime_error(std::string("NNEvaluator::loadGraph() : ") + agb_last_error(engine->handle));
		engine->cells = net.game.rows * net.game.cols;
		engine->q_head = net.q_head;
		const size_t n = config.batch_size;
		engine->features.resize(n * engine->cells);
		engine->boards.resize(n * engine->cells);
		engine->sign_to_move.resize(n);
		engine->symmetry_features.resize(n);
		engine->symmetry_boards.resize(n);
		engine->result_of.resize(n);
		engine->policy.resize(n * engine->cells);
		engine->value.resize(n * 3);
		engine->q.resize(n * engine->cells * 3);
		agb200::g_engines[this] = std::move(engine);
	}
	void NNEvaluator::unloadGraph()
	{
		std::lock_guard<std::mutex> lock(agb200::g_mutex);
		agb200::g_engines.erase(this);
	}
	void NNEvaluator::addToQueue(SearchTask &task)
	{ // the same draw, from the same generator, as the reference (NNEvaluator.cpp:134-139)
		const int r = number_of_available_symmetries(MatrixShape(task.getBoard().rows(), task.getBoard().cols()));
		if (use_symmetries)
			waiting_queue.push_back( { &task, randInt(r) });
		else
			waiting_queue.push_back( { &task, 0 });
	}
	void NNEvaluator::addToQueue(SearchTask &task, int symmetry)
	{
		waiting_queue.push_back( { &task, symmetry });
	}
	double NNEvaluator::evaluateGraph()
	{
		agb200::Engine *engine = agb200::find(this);
		if (engine == nullptr)
			throw std::logic_error("graph is empty - the network has not been loaded");
		while (waiting_queue.size() > 0)
		{
			const int batch_size = std::min(static_cast<int>(waiting_queue.size()), config.batch_size);
			stats.batch_sizes += batch_size;
			in_progress_queue.assign(waiting_queue.begin(), waiting_queue.begin() + batch_size);
			waiting_queue.erase(waiting_queue.begin(), waiting_queue.begin() + batch_size);
			pack_to_network();
			stats.compute.startTimer();
			const double t0 = getTime();
			const int rc = engine->evaluate();
			stats.compute.stopTimer();
			if (rc != AGB_OK)
				throw std::runtime_error(std::string("NNEvaluator::evaluateGraph() : ") + agb_last_error(engine->handle));
			engine->time_per_sample = (getTime() - t0) / batch_size;
			unpack_from_network();
			in_progress_queue.clear();
		}
		return engine->time_per_sample;
	}
	double NNEvaluator::asyncEvaluateGraphLaunch()
	{
		agb200::Engine *engine = agb200::find(this);
		if (engine == nullptr)
			throw std::logic_error("graph is empty - the network has not been loaded");
		if (not in_progress_queue.empty())
			throw std::logic_error("some tasks are already being processed");
		const int batch_size = std::min(static_cast<int>(waiting_queue.size()), config.batch_size);
		if (batch_size > 0)
		{
			in_progress_queue.assign(waiting_queue.begin(), waiting_queue.begin() + batch_size);
			waiting_queue.erase(waiting_queue.begin(), waiting_queue.begin() + batch_size);
			pack_to_network();
			stats.compute.startTimer();
			engine->launch_time = getTime();
			// the batch is on its way while the caller selects the next one (GeneratorManager.cpp:127-138); results are read in Join
			engine->pending = std::async(std::launch::async, [engine]()
			{
				return engine->evaluate();
			});
		}
		return getTime() + batch_size * engine->time_per_sample; // estimated end time, like PerfEstimator::getEstimatedEndTime
	}
	void NNEvaluator::asyncEvaluateGraphJoin()
	{
		agb200::Engine *engine = agb200::find(this);
		if (engine == nullptr)
			throw std::logic_error("graph is empty - the network has not been loaded");
		const int batch_size = in_progress_queue.size();
		if (batch_size > 0)
		{
			const int rc = engine->pending.get();
			stats.compute.stopTimer();
			if (rc != AGB_OK)
				throw std::runtime_error(std::string("NNEvaluator::asyncEvaluateGraphJoin() : ") + agb_last_error(engine->handle));
			engine->time_per_sample = (getTime() - engine->launch_time) / batch_size;
			stats.batch_sizes += batch_size;
			unpack_from_network();
			in_progress_queue.clear();
		}
	}
	/*
	 * private
	 */
	AGNetwork& NNEvaluator::get_network()
	{
		throw std::logic_error("NNEvaluator::get_network() : the B200 evaluator holds no host-side network");
	}
	const AGNetwork& NNEvaluator::get_network() const
	{
		throw std::logic_error("NNEvaluator::get_network() : the B200 evaluator holds no host-side network");
	}
	void NNEvaluator::pack_to_network()
	{ // NNEvaluator.cpp:244-262. A task the solver has processed carries its feature words: the reference augments them IN PLACE and packs them,
	  // so a task that sits in the queue twice goes to the network augmented twice -- taking the words from the task reproduces that. The other
	  // tasks are sent as boards with their symmetry; pattern calculation, encoding and augmentation then happen on the device.
		TimerGuard timer(stats.pack);
		agb200::Engine *engine = agb200::find(this);
		const int n = static_cast<int>(in_progress_queue.size());
		engine->n_features = 0;
		for (int i = 0; i < n; i++)
			engine->n_features += in_progress_queue[i].ptr->wasProcessedBySolver() ? 1 : 0;
		engine->n_boards = n - engine->n_features;
		int next_features = 0, next_boards = 0;
		for (int i = 0; i < n; i++)
		{
			const TaskData td = in_progress_queue.at(i);
			if (td.ptr->wasProcessedBySolver())
			{
				const int k = next_features++;
				td.ptr->getFeatures().augment(td.symmetry);
				std::memcpy(engine->features.data() + static_cast<size_t>(k) * engine->cells, td.ptr->getFeatures().data(), engine->cells * sizeof(uint32_t));
				engine->symmetry_features[k] = static_cast<int8_t>(td.symmetry);
				engine->result_of[i] = k;
			}
			else
			{
				const int k = next_boards++;
				const matrix<Sign> &board = td.ptr->getBoard();
				int8_t *dst = engine->boards.data() + static_cast<size_t>(k) * engine->cells;
				for (int j = 0; j < engine->cells; j++)
					dst[j] = static_cast<int8_t>(board[j]);
				engine->sign_to_move[k] = static_cast<int8_t>(td.ptr->getSignToMove());
				engine->symmetry_boards[k] = static_cast<int8_t>(td.symmetry);
				engine->result_of[i] = engine->n_features + k;
			}
		}
	}
	void NNEvaluator::unpack_from_network()
	{ // the engine has already applied the inverse symmetry to the policy and the action values (NNEvaluator.cpp:263-286)
		TimerGuard timer(stats.unpack);
		agb200::Engine *engine = agb200::find(this);
		for (size_t i = 0; i < in_progress_queue.size(); i++)
		{
			const TaskData td = in_progress_queue.at(i);
			const size_t row = engine->result_of[i];
			std::memcpy(td.ptr->getPolicy().data(), engine->policy.data() + row * engine->cells, engine->cells * sizeof(float));
			matrix<Value> &action_values = td.ptr->getActionValues();
			for (int j = 0; j < engine->cells; j++)
			{
				const float *src = engine->q.data() + (row * engine->cells + j) * 3;
				action_values[j] = engine->q_head ? Value(src[0], src[1]) : Value(0.0f, 0.0f); // a "pv" network has no 'q' output
			}
			td.ptr->setValue(Value(engine->value[3 * row], engine->value[3 * row + 1]));
			if (td.ptr->getScore().isUnproven())
				td.ptr->setMovesLeft(0.0f); // ResnetPV / PVQ have no moves-left head
			td.ptr->markAsProcessedByNetwork();
		}
	}
} /* namespace ag */
