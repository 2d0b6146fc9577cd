// K6 (PUCT select with virtual loss), K7 (edge generation, expand, backup) and the per-ply driver (final move, subtree
// reuse) of the lockstep self-play engine. All games of an engine advance together: one warp owns one game; the search
// graphs (nodes, edges, per-game transposition tables) live in HBM arenas and never visit the host.
//
// Reference semantics (file:line in the reference tree):
//   Search::select / Tree::select           src/search/monte_carlo/Search.cpp:117-158, Tree.cpp:226-251
//   PUCTSelector + PUCT / PUCT_q_head ops   src/search/monte_carlo/EdgeSelector.cpp:27-32, 335-361, 389-424, 562-586, 1123-1166
//   UnifiedGenerator::generate              src/search/monte_carlo/EdgeGenerator.cpp:23-40, 49-127, 128-179, 269-303
//   Tree::expand / backup / leak correction src/search/monte_carlo/Tree.cpp:75-104, 257-383
//   NodeCache (transpositions, cleanup)     src/search/monte_carlo/NodeCache.cpp:221-299
//   GameGenerator::generate / make_move     src/selfplay/GameGenerator.cpp:46-185, utils/misc.cpp:171-179
// Floating point follows the reference operation by operation (this file is compiled with -fmad=false; the reference is
// built without FMA contraction, CMakeLists.txt:52-54), so visit counts, values and move choices are bit-identical when
// the evaluator outputs are.
//
#include "engine.hpp"
#include "partition.hpp"
#include "patterns_logic.cuh"
#include "records.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

namespace agb
{
	using namespace plogic;

	constexpr int kMaxPath = 96;
	// exact position compare of the node cache: one bit per cell and colour; 7 words cover the 400 cells of a 20x20 board
	constexpr int kWordsPerColour = 7, kBitWords = 2 * kWordsPerColour;
	constexpr unsigned kFullMask = 0xFFFFFFFFu;

	// Score (search/Score.hpp:47-321): 3 bits proven value, 13 bits eval + 4000
	namespace score
	{
		enum : int { LOSS = 0, DRAW = 1, UNKNOWN = 2, WIN = 3 };
		AGB_HD inline uint16_t make(int pv, int eval) { return static_cast<uint16_t>((pv << 13) | (4000 + eval)); }
		AGB_HD inline int eval(uint16_t s) { return (s & 8191) - 4000; }
		AGB_HD inline int pv(uint16_t s) { return (s >> 13) & 3; }
		AGB_HD inline bool is_infinite(uint16_t s) { return s == 0x0000 or s == 0xFFFF; }
		AGB_HD inline bool is_proven(uint16_t s) { return pv(s) != UNKNOWN and not is_infinite(s); }
		AGB_HD inline bool is_win(uint16_t s) { return pv(s) == WIN and not is_infinite(s); }
		AGB_HD inline bool is_unproven(uint16_t s) { return pv(s) == UNKNOWN; }
		AGB_HD inline int distance(uint16_t s)
		{
			switch (pv(s))
			{
				case LOSS:
				case DRAW: return eval(s);
				case WIN: return -eval(s);
				default: return 0;
			}
		}
		AGB_HD inline uint16_t negate(uint16_t s)
		{
			switch (pv(s))
			{
				case LOSS: return is_infinite(s) ? 0xFFFF : make(WIN, -eval(s));
				case DRAW: return make(DRAW, eval(s));
				case WIN: return is_infinite(s) ? 0x0000 : make(LOSS, -eval(s));
				default: return make(UNKNOWN, -eval(s));
			}
		}
		AGB_HD inline uint16_t invert_up(uint16_t s)
		{
			switch (pv(s))
			{
				case LOSS: return is_infinite(s) ? negate(s) : make(WIN, -(distance(s) + 1));
				case DRAW: return make(DRAW, distance(s) + 1);
				case WIN: return is_infinite(s) ? negate(s) : make(LOSS, distance(s) + 1);
				default: return negate(s);
			}
		}
		constexpr uint16_t kDefault = (UNKNOWN << 13) | 4000; // Score()
	}

	struct NodeD
	{
			int32_t edge_begin;
			int16_t n_edges;
			int16_t depth;
			float win, draw;
			float moves_left;
			int32_t visits;
			uint16_t score;
			int16_t vloss;
			int8_t stm;
			uint8_t flags; // bit0 root, bit1 fully expanded
			uint16_t pad;
			uint64_t hash;
	};
	struct EdgeD
	{
			float prior;
			float win, draw;
			int32_t visits;
			uint16_t move; // Move::toShort
			uint16_t score;
			uint16_t vloss_flag; // bit 15: being expanded, low 15 bits: virtual loss (Edge.hpp:25-46)
			uint16_t pad;
	};
	static_assert(sizeof(NodeD) == 40 and sizeof(EdgeD) == 24, "layout");

	struct TaskD
	{
			int32_t path_len;
			int32_t final_node;
			int32_t nn_slot; // slot in the evaluation batch, -1 if not evaluated
			int8_t stm;
			uint8_t stored; // task is part of the batch
			uint8_t proven_edge; // reached a proven edge: no solver / NN / edge generation
			uint8_t sticky; // solver flags of this task slot; never cleared, like the reference's reused SearchTask (bit0 static, bit1 recursive)
			float win, draw, moves_left;
			uint16_t score;
			uint16_t pad2;
			uint64_t hash;
			uint64_t bits[kBitWords]; // cross words, then circle words
			int32_t path_node[kMaxPath];
			int32_t path_edge[kMaxPath];
	};

	constexpr int kMaxGroups = 8;
	struct SelfplayState
	{
			int games = 0, batch = 0, cells = 0, S = 0;
			int max_nodes = 0, max_edges = 0, table_size = 0;
			// per game
			int8_t *root_board = nullptr; // [games][cells]
			uint64_t *root_bits = nullptr; // [games][kBitWords]
			uint64_t *root_hash = nullptr;
			int8_t *root_stm = nullptr;
			int32_t *root_node = nullptr;
			int32_t *n_nodes = nullptr, *n_edges = nullptr;
			int32_t *n_stored = nullptr; // tasks stored in the current batch
			int32_t *n_moves = nullptr;
			uint16_t *moves = nullptr; // [games][cells] played moves (incl. opening)
			int8_t *outcome = nullptr; // last finished outcome
			NodeD *nodes = nullptr; // [games][max_nodes]
			uint64_t *node_bits = nullptr; // [games][max_nodes][kBitWords]
			EdgeD *edges = nullptr; // [games][max_edges]
			int32_t *table = nullptr; // [games][table_size] open addressing, -1 empty
			int32_t *remap = nullptr; // [games][max_nodes] scratch for cleanup
			TaskD *tasks = nullptr; // [games][batch]
			// evaluation batch
			int8_t *task_boards = nullptr; // [games*batch][cells] compact (K1 input)
			int8_t *task_stm = nullptr;
			int32_t *eval_count = nullptr; // device counter: slots of the current batch (all leaf positions)
			uint8_t *slot_is_root = nullptr; // [games*batch]
			int32_t *nn_list = nullptr, *nn_count = nullptr; // slots that go to the network when the solver is on
			SolverOutputs solver_out { };
			SolverState solver { }; // per-game search memory (solver.cu)
			// pipeline groups: the games are split into `groups` independent halves that advance on their own streams, so that the
			// solver / tree kernels of one half overlap the network kernel of the other (the network launches share one stream)
			int groups = 1;
			int solver_sms = 0, net_sms = 0; // SM partition between K5 and K4 (AgbConfig::solver_sms; 0 = none)
			// the partition as two green contexts (partition.hpp): the group streams live in the solver's, the network stream in the other. When
			// the driver cannot provide them the partition falls back to SM-filling solver blocks (solver.cu) and at most 2 groups pay
			SmPartition partition { };
			bool green = false;
			bool reset_done = false; // agb_selfplay_reset / agb_load_games has put every per-game array into a defined state
			cudaStream_t group_stream[kMaxGroups] = { };
			cudaStream_t solver_stream[kMaxGroups] = { }; // green partition: K5 of each group on the solver's SMs, everything else of the group on the tree SMs
			cudaEvent_t solver_go[kMaxGroups] = { }, solver_done[kMaxGroups] = { };
			cudaStream_t nn_stream = nullptr;
			cudaStream_t nn_group_stream[kMaxGroups] = { }; // green partition: one network stream per group, so that K4 launches run in the order their inputs become ready
			cudaEvent_t ready[kMaxGroups] = { }, evaluated[kMaxGroups] = { }, joined = nullptr;
			uint32_t *features = nullptr;
			float *policy = nullptr, *value = nullptr, *q = nullptr;
			// SelfplayConfig::use_symmetries: every evaluation goes through a random board symmetry (NNEvaluator.cpp:134-146, 244-286)
			// EdgeSelectorConfig::noise_type / noise_weight: the root's priors as the tree selector sees them (PUCTSelector::noisy_policy)
			// agb_think: games search one move and then wait for the host (evaluation games between two engines, Player::getMove)
			uint8_t *paused = nullptr; // [games] 1: this game takes no part in the lockstep iterations
			uint16_t *decision = nullptr; // [games] the move chosen by the last search (Move::toShort), 0 = none yet
			float *root_noise = nullptr; // [games][cells], indexed by the root's edge index
			uint8_t *noise_ready = nullptr; // [games] drawn for the current search
			uint32_t *noise_counter = nullptr; // [games] searches drawn so far
			int8_t *task_sym = nullptr; // [games*batch] symmetry of the slot's evaluation
			uint32_t *sym_counter = nullptr; // [games] evaluations drawn so far (the random stream is keyed by the global game id)
			int8_t *sym_table = nullptr; // agb_set_symmetry_table: replayed symmetries instead of the random stream
			int sym_table_size = 0;
			uint32_t *features_aug = nullptr; // [games*batch][cells] augmented feature words
			float *policy_raw = nullptr, *q_raw = nullptr; // network outputs before the inverse symmetry
			uint64_t *zobrist = nullptr; // [cells][2] + [2]
			unsigned long long *stats = nullptr; // device AgbStats mirror (16 words)
			// openings pool
			int8_t *openings = nullptr; // [n_openings][cells]
			int8_t *opening_stm = nullptr;
			int n_openings = 0;
			uint32_t *opening_cursor = nullptr; // [games] restarts of each game so far (selects its next opening)
			// last sample per game (dense), for parity checks and K8
			int32_t *sample_visits = nullptr; // [games][cells]
			float *sample_prior = nullptr, *sample_win = nullptr, *sample_draw = nullptr; // [games][cells]
			uint16_t *sample_score = nullptr; // [games][cells]
			// K8: format-201 records. Per game: [u32 sample count][samples...] grows ply by ply; finished games are appended to a
			// global queue as complete GameDataStorage blobs
			uint8_t *rec_buf = nullptr; // [games][rec_cap]
			int32_t *rec_len = nullptr, *rec_samples = nullptr;
			int rec_cap = 0;
			uint8_t *fin_buf = nullptr;
			unsigned long long *fin_used = nullptr; // bytes used in fin_buf
			int32_t *fin_games = nullptr;
			unsigned long long fin_cap = 0;
			float *sample_root = nullptr; // [games][4] win, draw, (visits), (move)
	};

	namespace
	{
		enum : int { ST_EVALS = 0, ST_NODES = 1, ST_DUP = 2, ST_LEAKS = 3, ST_PROVEN = 4, ST_WASTED = 5, ST_MOVES = 6, ST_GAMES = 7, ST_OVERFLOW = 8 };
		enum : uint32_t { OVF_NODES = 2, OVF_EDGES = 4, OVF_PATH = 8, OVF_TABLE = 16, OVF_RECORD = 32, OVF_FINISHED = 64 };

		struct Params
		{
				SelfplayState s;
				int rules, draw_after;
				int max_simulations, init_to;
				float exploration_constant, leak_threshold;
				int q_head;
				int solver_mode; // 0: terminal checks in K7, 1: K5 solver on every leaf
				// pipeline group served by this launch: games [game_begin, game_begin + game_count), evaluation slots from slot_base on
				int game_begin, game_count, slot_base;
				int32_t *eval_count; // this group's slot counter
				int use_symmetries, first_game_id;
				int max_children; // MCTSConfig::max_children (0: unlimited)
				int final_selector; // SelfplayConfig::final_selector.policy (AGB_FINAL_*)
				int single_move; // agb_think: a game stops at its decision instead of playing on
				float policy_temperature; // MCTSConfig::policy_temperature (initialize_edges, EdgeGenerator.cpp:88-127)
				int noise_type; // AGB_NOISE_*
				float noise_weight;
				float final_exploration; // its exploration_constant (lcb)
				float expansion_threshold; // MCTSConfig::policy_expansion_threshold
				unsigned long long sym_seed;
				SymmetryStream sym;
				Tables tables;
				BoardStore store;
				uint32_t *status;
		};

		__device__ inline float expectation(float win, float draw)
		{
			return win + 0.5f * draw;
		}

		// ---- transposition table ---------------------------------------------------------------------------------
		__device__ int table_seek(const Params &p, int g, uint64_t hash, const uint64_t *bits, int stm, int lane)
		{ // NodeCache::seek: hash, then full board and side-to-move comparison
			const int mask = p.s.table_size - 1;
			const int32_t *table = p.s.table + static_cast<size_t>(g) * p.s.table_size;
			const NodeD *nodes = p.s.nodes + static_cast<size_t>(g) * p.s.max_nodes;
			int slot = static_cast<int>(hash) & mask;
			for (int probe = 0; probe < p.s.table_size; probe++, slot = (slot + 1) & mask)
			{
				const int idx = table[slot];
				if (idx < 0)
					return -1;
				if (nodes[idx].hash == hash and nodes[idx].stm == stm)
				{
					const uint64_t *nb = p.s.node_bits + (static_cast<size_t>(g) * p.s.max_nodes + idx) * kBitWords;
					const bool same = (lane >= kBitWords) or (nb[lane] == bits[lane]);
					if (__all_sync(kFullMask, same))
						return idx;
				}
			}
			return -1;
		}
		__device__ void table_insert(const Params &p, int g, uint64_t hash, int idx)
		{ // single thread
			const int mask = p.s.table_size - 1;
			int32_t *table = p.s.table + static_cast<size_t>(g) * p.s.table_size;
			int slot = static_cast<int>(hash) & mask;
			for (int probe = 0; probe < p.s.table_size; probe++, slot = (slot + 1) & mask)
				if (table[slot] < 0)
				{
					table[slot] = idx;
					return;
				}
			atomicOr(p.status, OVF_TABLE);
		}

		// ---- score bookkeeping (Tree.cpp:75-104) ----------------------------------------------------------------------
		__device__ bool has_information_leak(const EdgeD &e, const NodeD *node, float threshold)
		{
			if (node == nullptr or threshold >= 1.0f)
				return false;
			if (e.score != score::invert_up(node->score))
				return true;
			const float inv_win = 1.0f - (node->win + node->draw); // Value::getInverted: (loss_rate, draw_rate)
			const float dw = e.win - inv_win, dd = e.draw - node->draw;
			return (fabsf(dw) + fabsf(dd)) > threshold;
		}
		__device__ void update_node_score(NodeD &node, const EdgeD *edges, int lane)
		{ // max over edge scores; only a fully expanded node (or a win / unproven result) takes it over
			uint16_t best = 0;
			for (int i = lane; i < node.n_edges; i += 32)
				best = max(best, edges[node.edge_begin + i].score);
			for (int o = 16; o > 0; o >>= 1)
				best = max(best, static_cast<uint16_t>(__shfl_xor_sync(kFullMask, static_cast<int>(best), o)));
			if ((node.flags & 2) or score::is_win(best) or score::is_unproven(best))
				node.score = best;
		}
		__device__ void clip01(float &x)
		{
			x = fmaxf(0.0f, fminf(1.0f, x));
		}

		// Tree::correctInformationLeak (Tree.cpp:352-376); all lanes run it, lane 0 writes
		__device__ void correct_information_leak(const Params &p, int g, TaskD &task, int lane)
		{
			NodeD *nodes = p.s.nodes + static_cast<size_t>(g) * p.s.max_nodes;
			EdgeD *edges = p.s.edges + static_cast<size_t>(g) * p.s.max_edges;
			for (int i = task.path_len - 1; i >= 0; i--)
			{
				NodeD &node = nodes[task.path_node[i]];
				EdgeD &edge = edges[task.path_edge[i]];
				const int next = (i == task.path_len - 1) ? task.final_node : task.path_node[i + 1];
				const NodeD &nn = nodes[next];
				const float cw = edge.win, cd = edge.draw;
				const float tw = 1.0f - (nn.win + nn.draw), td = nn.draw;
				const float scale = static_cast<float>(edge.visits) / static_cast<float>(node.visits);
				const float nw = node.win + (tw - cw) * scale, nd = node.draw + (td - cd) * scale;
				__syncwarp();
				if (lane == 0)
				{
					edge.win = tw;
					edge.draw = td;
					node.win = nw;
					node.draw = nd;
					edge.score = score::invert_up(nn.score);
				}
				__syncwarp();
				NodeD tmp = node;
				update_node_score(tmp, edges, lane);
				if (lane == 0)
					node.score = tmp.score;
				__syncwarp();
			}
		}

		// ---- root noise (EdgeSelector.cpp:602-623, random.cpp:89-123). The reference draws from a thread-local std::mt19937; here a counter-based
		// stream keyed by (seed, global game id, search number) gives the same distributions independently of how games are sharded. ----
		__device__ inline float noise_uniform(unsigned long long key, unsigned int index)
		{ // (0, 1)
			unsigned long long z = key + 0x9E3779B97F4A7C15ull * (index + 1ull);
			z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
			z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
			z ^= z >> 31;
			return (static_cast<float>(z >> 40) + 0.5f) * (1.0f / 16777216.0f);
		}
		__device__ void draw_root_noise(const Params &p, int g, const NodeD &root, const EdgeD *edges)
		{ // single thread; fills p.s.root_noise[g][0 .. n_edges) with PUCTSelector::noisy_policy
			float *noisy = p.s.root_noise + static_cast<size_t>(g) * p.s.cells;
			const int n = root.n_edges;
			const float w = p.noise_weight;
			const unsigned long long key = (p.sym_seed ^ 0x6E6F697365ull) + (static_cast<unsigned long long>(p.first_game_id + g) << 32) + p.s.noise_counter[g]++;
			unsigned int draw = 0;
			if (p.noise_type == AGB_NOISE_CUSTOM)
			{ // createCustomNoise: u^4 of what is left, then a uniform shuffle
				float sum = 0.0f;
				for (int i = 0; i < n; i++)
				{
					const float u = noise_uniform(key, draw++);
					noisy[i] = static_cast<float>(pow(static_cast<double>(u), 4.0) * (1.0f - sum));
					sum += noisy[i];
				}
				for (int i = n - 1; i > 0; i--)
				{ // Fisher-Yates
					const int j = min(i, static_cast<int>(noise_uniform(key, draw++) * (i + 1)));
					const float t = noisy[i];
					noisy[i] = noisy[j];
					noisy[j] = t;
				}
				for (int i = 0; i < n; i++)
					noisy[i] = (1.0f - w) * edges[root.edge_begin + i].prior + w * noisy[i];
			}
			else if (p.noise_type == AGB_NOISE_DIRICHLET)
			{ // createDirichletNoise(size, 0.05): normalised gamma(0.05) draws (Marsaglia-Tsang on shape + 1, boosted by u^(1/shape))
				const double shape = 0.05, d = shape + 1.0 - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
				float sum = 0.0f;
				for (int i = 0; i < n; i++)
				{
					double sample = 0.0;
					for (int attempt = 0; attempt < 64; attempt++)
					{
						const double u1 = noise_uniform(key, draw++), u2 = noise_uniform(key, draw++), u3 = noise_uniform(key, draw++);
						const double x = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2); // standard normal
						const double v = (1.0 + c * x) * (1.0 + c * x) * (1.0 + c * x);
						if (v > 0.0 and log(u3) < 0.5 * x * x + d - d * v + d * log(v))
						{
							sample = d * v;
							break;
						}
					}
					sample *= pow(static_cast<double>(noise_uniform(key, draw++)), 1.0 / shape);
					noisy[i] = static_cast<float>(sample);
					sum += noisy[i];
				}
				const float inv = (sum > 0.0f) ? 1.0f / sum : 0.0f;
				for (int i = 0; i < n; i++)
					noisy[i] = (1.0f - w) * edges[root.edge_begin + i].prior + w * (sum > 0.0f ? noisy[i] * inv : 1.0f / n);
			}
			else
			{ // gumbel: softmax(log_eps(prior) + w * g), g = -log_eps(-log_eps(u))
				const float eps = 1.1920929e-07f;
				float m = -FLT_MAX;
				for (int i = 0; i < n; i++)
				{
					const float gum = -logf(eps - logf(eps + noise_uniform(key, draw++)));
					noisy[i] = logf(eps + edges[root.edge_begin + i].prior) + w * gum;
					m = fmaxf(m, noisy[i]);
				}
				float sum = 0.0f;
				for (int i = 0; i < n; i++)
				{
					noisy[i] = expf(noisy[i] - m);
					sum += noisy[i];
				}
				for (int i = 0; i < n; i++)
					noisy[i] /= sum;
			}
		}

		// ---- K6: select ------------------------------------------------------------------------------------------------
		__global__ void __launch_bounds__(128) select_kernel(const __grid_constant__ Params p)
		{
			const int lane = threadIdx.x & 31;
			if (blockIdx.x * 4 + (threadIdx.x >> 5) >= p.game_count)
				return;
			const int g = p.game_begin + blockIdx.x * 4 + (threadIdx.x >> 5);
			const int cells = p.s.cells, S = p.s.S;
			NodeD *nodes = p.s.nodes + static_cast<size_t>(g) * p.s.max_nodes;
			EdgeD *edges = p.s.edges + static_cast<size_t>(g) * p.s.max_edges;
			TaskD *tasks = p.s.tasks + static_cast<size_t>(g) * p.s.batch;
			if (p.s.paused[g])
			{ // waiting for the host (agb_think): no tasks, nothing to solve or evaluate
				if (lane == 0)
				{
					p.s.n_stored[g] = 0;
					if (p.solver_mode != 0)
						p.s.solver.game_slot_count[g] = 0;
				}
				return;
			}
			const int root = p.s.root_node[g];
			const int root_visits = (root >= 0) ? nodes[root].visits : 0;
			int stored = 0;
			int trials = 2 * p.s.batch;
			while (stored < p.s.batch and root_visits <= p.max_simulations)
			{
				TaskD &task = tasks[stored];
				// SearchTask::set
				__syncwarp();
				if (lane < kBitWords)
					task.bits[lane] = p.s.root_bits[static_cast<size_t>(g) * kBitWords + lane];
				if (lane == 0)
				{
					task.path_len = 0;
					task.final_node = -1;
					task.nn_slot = -1;
					task.stored = 1;
					task.proven_edge = 0;
					task.win = task.draw = task.moves_left = 0.0f;
					task.score = score::kDefault;
				}
				uint64_t hash = p.s.root_hash[g];
				int stm = p.s.root_stm[g];
				int node = root, len = 0;
				int outcome = 0; // 0 leaf, 1 proven edge, 2 information leak
				__syncwarp();
				while (node >= 0)
				{
					const NodeD N = nodes[node];
					// PUCTSelector::select: c = exploration_constant (+ 0 * log(..))
					const float psv = static_cast<float>(static_cast<double>(p.exploration_constant) * sqrt(static_cast<double>(N.visits + N.vloss)));
					float initial_q = 0.0f;
					if (p.init_to == 1)
						initial_q = expectation(N.win, N.draw);
					else if (p.init_to == 2)
						initial_q = 0.5f;
					// root noise: drawn once per search, then used instead of the priors at the root (PUCTSelector::select)
					const bool use_noise = p.noise_type != 0 and (N.flags & 1) != 0;
					const float *noisy = p.s.root_noise + static_cast<size_t>(g) * cells;
					if (use_noise)
					{
						if (lane == 0 and p.s.noise_ready[g] == 0)
						{
							draw_root_noise(p, g, N, edges);
							p.s.noise_ready[g] = 1;
						}
						__syncwarp();
					}
					float best_v = -FLT_MAX;
					int best_i = 0x7FFFFFFF;
					for (int i = lane; i < N.n_edges; i += 32)
					{
						const EdgeD e = edges[N.edge_begin + i];
						const int vl = e.vloss_flag & 0x7FFF;
						const float prior = use_noise ? noisy[i] : e.prior;
						float v;
						switch (score::is_proven(e.score) ? score::pv(e.score) : static_cast<int>(score::UNKNOWN))
						{
							case score::LOSS:
								v = -1000.0f + score::distance(e.score);
								break;
							case score::DRAW:
								v = 0.5f;
								break;
							case score::WIN:
								v = +1000.0f - score::distance(e.score);
								break;
							default:
							{
								const float visits_f = 1.0e-8f + e.visits;
								const float vloss_scale = visits_f / (visits_f + static_cast<float>(vl));
								float Q;
								if (e.vloss_flag & 0x8000)
									Q = -1000.0f;
								else if (p.init_to == 3 or e.visits > 0)
									Q = expectation(e.win, e.draw) * vloss_scale;
								else
									Q = initial_q;
								const float U = prior * psv / (1.0f + e.visits + vl);
								v = Q + U;
								break;
							}
						}
						if (v > best_v)
						{
							best_v = v;
							best_i = i;
						}
					}
					for (int o = 16; o > 0; o >>= 1)
					{ // first maximal edge wins (find_best_edge_impl uses a strict '>')
						const float ov = __shfl_xor_sync(kFullMask, best_v, o);
						const int oi = __shfl_xor_sync(kFullMask, best_i, o);
						if (ov > best_v or (ov == best_v and oi < best_i))
						{
							best_v = ov;
							best_i = oi;
						}
					}
					const int eidx = N.edge_begin + best_i;
					const EdgeD chosen = edges[eidx];
					// SearchTask::append + virtual loss
					const int row = (chosen.move >> 2) & 127, col = (chosen.move >> 9) & 127;
					const int cell = row * S + col;
					const int colour = stm - 1;
					hash ^= p.s.zobrist[cell * 2 + colour] ^ p.s.zobrist[cells * 2] ^ p.s.zobrist[cells * 2 + 1];
					__syncwarp();
					if (lane == 0)
					{
						if (len < kMaxPath)
						{
							task.path_node[len] = node;
							task.path_edge[len] = eidx;
						}
						else
							atomicOr(p.status, OVF_PATH);
						task.bits[colour * kWordsPerColour + (cell >> 6)] |= 1ull << (cell & 63);
						nodes[node].vloss = N.vloss + 1;
						edges[eidx].vloss_flag = (chosen.vloss_flag & 0x8000) | ((chosen.vloss_flag & 0x7FFF) + 1);
					}
					len = min(len + 1, kMaxPath);
					stm = 3 - stm;
					__syncwarp();
					if (score::is_proven(chosen.score))
					{
						outcome = 1;
						break;
					}
					const int child = table_seek(p, g, hash, task.bits, stm, lane);
					if (lane == 0)
					{
						task.final_node = child;
						if (child < 0)
							edges[eidx].vloss_flag |= 0x8000;
					}
					__syncwarp();
					if (has_information_leak(edges[eidx], child >= 0 ? &nodes[child] : nullptr, p.leak_threshold))
					{
						outcome = 2;
						break;
					}
					node = child;
				}
				if (lane == 0)
				{
					task.path_len = len;
					task.stm = static_cast<int8_t>(stm);
					task.hash = hash;
				}
				__syncwarp();
				if (len == 0)
				{ // empty tree: the root itself is the only task of this batch (Search.cpp:128-129)
					stored++;
					break;
				}
				bool keep = true;
				if (lane == 0)
				{ // statistics only (Search.cpp:130-131)
					for (int i = 0; i < stored; i++)
						if (tasks[i].path_len > 0 and tasks[i].path_edge[tasks[i].path_len - 1] == task.path_edge[len - 1])
						{
							atomicAdd(p.s.stats + ST_DUP, 1ull);
							break;
						}
				}
				if (outcome == 2)
				{ // Search.cpp:132-138
					correct_information_leak(p, g, task, lane);
					if (lane == 0)
					{
						for (int i = 0; i < len; i++)
						{
							nodes[task.path_node[i]].vloss -= 1;
							EdgeD &e = edges[task.path_edge[i]];
							e.vloss_flag = (e.vloss_flag & 0x8000) | ((e.vloss_flag & 0x7FFF) - 1);
						}
						task.stored = 0;
						atomicAdd(p.s.stats + ST_LEAKS, 1ull);
					}
					keep = false;
				}
				else if (outcome == 1)
				{ // Search.cpp:139-150
					if (lane == 0)
					{
						const uint16_t sc = edges[task.path_edge[len - 1]].score;
						task.final_node = -1;
						task.stm = static_cast<int8_t>(3 - stm);
						task.score = sc;
						switch (score::pv(sc))
						{ // Score::convertToValue
							case score::LOSS: task.win = 0.0f; task.draw = 0.0f; break;
							case score::DRAW: task.win = 0.0f; task.draw = 1.0f; break;
							default: task.win = 1.0f; task.draw = 0.0f; break;
						}
						task.proven_edge = 1;
						atomicAdd(p.s.stats + ST_PROVEN, 1ull);
					}
				}
				__syncwarp();
				if (keep)
					stored++;
				if (--trials <= 0)
					break;
			}
			if (lane == 0)
				p.s.n_stored[g] = stored;
			// evaluation batch: every stored task that is a root or not proven goes to the network (Search::scheduleToNN)
			int n_slots = 0;
			for (int t = 0; t < stored; t++)
			{
				TaskD &task = tasks[t];
				if (task.proven_edge)
					continue;
				int slot = 0;
				if (lane == 0)
				{
					slot = p.slot_base + atomicAdd(p.eval_count, 1);
					task.nn_slot = slot;
					p.s.task_stm[slot] = task.stm;
					p.s.slot_is_root[slot] = (task.path_len == 0) ? 1 : 0;
					if (p.use_symmetries and p.solver_mode == 0) // with the solver on, K5 draws for the tasks it schedules (SolverState::sym)
						p.s.task_sym[slot] = static_cast<int8_t>(p.sym.draw(g));
					if (p.solver_mode != 0)
						p.s.solver.game_slots[static_cast<size_t>(g) * p.s.batch + n_slots] = slot; // the solver walks them in task order
				}
				n_slots++;
				slot = __shfl_sync(kFullMask, slot, 0);
				for (int i = lane; i < cells; i += 32)
				{
					const uint64_t cross = task.bits[i >> 6], circle = task.bits[kWordsPerColour + (i >> 6)];
					p.s.task_boards[static_cast<size_t>(slot) * cells + i] = static_cast<int8_t>(((cross >> (i & 63)) & 1) | (((circle >> (i & 63)) & 1) << 1));
				}
			}
			if (lane == 0 and p.solver_mode != 0)
				p.s.solver.game_slot_count[g] = n_slots;
		}

		// initialize_edges (EdgeGenerator.cpp:88-127): the prior an edge gets from the policy value of its cell
		__device__ inline float tempered_prior(float policy_value, float temperature, float max_policy)
		{
			if (temperature == 1.0f)
				return policy_value;
			if (temperature == 0.0f)
				return (policy_value == max_policy) ? 1.0f : 0.0f;
			return powf(policy_value, 1.0f / temperature);
		}

		// ---- std::partial_sort as libstdc++ runs it (bits/stl_algo.h __partial_sort -> __heap_select + __sort_heap, bits/stl_heap.h),
		// with EdgeComparator<MaxPolicyPrior> (Edge.hpp:156-171). The order of equal elements is decided by these heap moves, so
		// prune_weak_moves (EdgeGenerator.cpp:69-84) is reproduced move for move. Single thread.
		__device__ inline bool edge_before(const EdgeD &a, const EdgeD &b)
		{
			if ((score::is_proven(a.score) or score::is_proven(b.score)) and a.score != b.score)
				return a.score > b.score;
			return a.prior > b.prior;
		}
		__device__ void heap_push(EdgeD *first, int hole, int top, const EdgeD &value)
		{
			int parent = (hole - 1) / 2;
			while (hole > top and edge_before(first[parent], value))
			{
				first[hole] = first[parent];
				hole = parent;
				parent = (hole - 1) / 2;
			}
			first[hole] = value;
		}
		__device__ void heap_adjust(EdgeD *first, int hole, int len, const EdgeD &value)
		{
			const int top = hole;
			int child = hole;
			while (child < (len - 1) / 2)
			{
				child = 2 * (child + 1);
				if (edge_before(first[child], first[child - 1]))
					child--;
				first[hole] = first[child];
				hole = child;
			}
			if ((len & 1) == 0 and child == (len - 2) / 2)
			{
				child = 2 * (child + 1);
				first[hole] = first[child - 1];
				hole = child - 1;
			}
			heap_push(first, hole, top, value);
		}
		__device__ void partial_sort_edges(EdgeD *first, int middle, int last)
		{
			if (middle >= 2)
				for (int parent = (middle - 2) / 2; parent >= 0; parent--)
				{ // __make_heap
					const EdgeD value = first[parent];
					heap_adjust(first, parent, middle, value);
				}
			for (int i = middle; i < last; i++)
				if (edge_before(first[i], first[0]))
				{ // __pop_heap(first, middle, i)
					const EdgeD value = first[i];
					first[i] = first[0];
					heap_adjust(first, 0, middle, value);
				}
			for (int end = middle; end > 1;)
			{ // __sort_heap
				end--;
				const EdgeD value = first[end];
				first[end] = first[0];
				heap_adjust(first, 0, end, value);
			}
		}

		// ---- K7: edge generation + expand + backup ------------------------------------------------------------------------
		__device__ void node_update_value(NodeD &n, float win, float draw)
		{ // Node::updateValue (Node.hpp:268-274): 1.0 / visits in double, narrowed to float
			n.visits++;
			const float tmp = static_cast<float>(1.0 / static_cast<double>(n.visits));
			n.win += (win - n.win) * tmp;
			n.draw += (draw - n.draw) * tmp;
			clip01(n.win);
			clip01(n.draw);
		}
		__device__ void edge_update_value(EdgeD &e, float win, float draw)
		{ // Edge::updateValue (Edge.hpp:111-117): 1.0f / visits in float
			e.visits++;
			const float tmp = 1.0f / static_cast<float>(e.visits);
			e.win += (win - e.win) * tmp;
			e.draw += (draw - e.draw) * tmp;
			clip01(e.win);
			clip01(e.draw);
		}

		__global__ void __launch_bounds__(128) expand_backup_kernel(const __grid_constant__ Params p)
		{
			const int lane = threadIdx.x & 31;
			if (blockIdx.x * 4 + (threadIdx.x >> 5) >= p.game_count)
				return;
			const int g = p.game_begin + blockIdx.x * 4 + (threadIdx.x >> 5);
			const int cells = p.s.cells, S = p.s.S;
			NodeD *nodes = p.s.nodes + static_cast<size_t>(g) * p.s.max_nodes;
			EdgeD *edges = p.s.edges + static_cast<size_t>(g) * p.s.max_edges;
			TaskD *tasks = p.s.tasks + static_cast<size_t>(g) * p.s.batch;
			const int stored = p.s.n_stored[g];

			for (int t = 0; t < stored; t++)
			{
				TaskD &task = tasks[t];
				if (not task.stored or task.proven_edge)
					continue; // proven-edge tasks skip edge generation and carry no edges (Tree::expand returns SKIPPED)
				const int slot = task.nn_slot;
				const int stm = task.stm;
				const size_t cbase = static_cast<size_t>(slot) * kCellPitch;
				int n_nodes = p.s.n_nodes[g], n_edges = p.s.n_edges[g];
				if (lane == 0 and p.solver_mode == 0)
				{ // NNEvaluator::unpack_from_network: value always, moves left only if the score is unproven
					task.win = p.s.value[slot * 3 + 0];
					task.draw = p.s.value[slot * 3 + 1];
					task.moves_left = 0.0f;
					atomicAdd(p.s.stats + ST_EVALS, 1ull);
				}
				__syncwarp();
				// UnifiedGenerator::generate, "not processed by solver" path: one edge per empty cell in row-major order,
				// written straight into the arena at the next free position (committed only if the node is new)
				int stones = 0;
				for (int i = lane; i < cells; i += 32)
					stones += (p.store.board[cbase + i] != NONE);
				for (int o = 16; o > 0; o >>= 1)
					stones += __shfl_xor_sync(kFullMask, stones, o);
				float max_policy = 0.0f; // maxValue(task.getPolicy()) over the whole board, only needed at temperature 0
				if (p.policy_temperature == 0.0f)
				{
					max_policy = -FLT_MAX;
					for (int i = lane; i < cells; i += 32)
						max_policy = fmaxf(max_policy, p.s.policy[static_cast<size_t>(slot) * cells + i]);
					for (int o = 16; o > 0; o >>= 1)
						max_policy = fmaxf(max_policy, __shfl_xor_sync(kFullMask, max_policy, o));
				}
				int count = 0;
				uint16_t tscore = score::kDefault;
				bool must_defend = false;
				if (n_edges + (cells - stones) > p.s.max_edges)
				{
					atomicOr(p.status, OVF_EDGES);
					continue;
				}
				if (p.solver_mode == 0)
				{
				const bool draw_now = (stones + 1) >= p.draw_after;
				int wins = 0, draws = 0, losses = 0;
				for (int i0 = 0; i0 < cells; i0 += 32)
				{
					const int i = i0 + lane;
					const bool empty = (i < cells) and (p.store.board[cbase + i] == NONE);
					const unsigned em = __ballot_sync(kFullMask, empty);
					int kind = score::UNKNOWN;
					if (empty)
					{ // check_terminal_conditions via getOutcome: five for the mover, renju foul, draw by move count
						const uint32_t pt = p.store.ptypes[cbase + i] >> (stm == CROSS ? 0 : 4);
						const bool five = ((pt & 7u) == PT_FIVE) or (((pt >> 8) & 7u) == PT_FIVE) or (((pt >> 16) & 7u) == PT_FIVE) or (((pt >> 24) & 7u) == PT_FIVE);
						if (five)
							kind = score::WIN;
						else if (p.rules == RULE_RENJU and stm == CROSS and p.store.forbidden[cbase + i])
							kind = score::LOSS;
						else if (draw_now)
							kind = score::DRAW;
						EdgeD e;
						e.prior = tempered_prior(p.s.policy[static_cast<size_t>(slot) * cells + i], p.policy_temperature, max_policy);
						e.win = p.q_head ? p.s.q[(static_cast<size_t>(slot) * cells + i) * 3 + 0] : 0.0f;
						e.draw = p.q_head ? p.s.q[(static_cast<size_t>(slot) * cells + i) * 3 + 1] : 0.0f;
						e.visits = 0;
						e.move = static_cast<uint16_t>(stm | ((i / S) << 2) | ((i % S) << 9));
						e.score = score::kDefault;
						e.vloss_flag = 0;
						e.pad = 0;
						if (kind == score::WIN)
						{
							e.score = score::make(score::WIN, -1);
							e.win = 1.0f;
							e.draw = 0.0f;
						}
						else if (kind == score::LOSS)
						{
							e.score = score::make(score::LOSS, 1);
							e.win = 0.0f;
							e.draw = 0.0f;
						}
						else if (kind == score::DRAW)
						{
							e.score = score::make(score::DRAW, 1);
							e.win = 0.0f;
							e.draw = 1.0f;
						}
						edges[n_edges + count + __popc(em & ((1u << lane) - 1u))] = e;
					}
					count += __popc(em);
					wins += __popc(__ballot_sync(kFullMask, kind == score::WIN));
					draws += __popc(__ballot_sync(kFullMask, kind == score::DRAW));
					losses += __popc(__ballot_sync(kFullMask, kind == score::LOSS));
				}
				__syncwarp();
				// position score from the terminal checks (EdgeGenerator.cpp:163-178)
				if (wins > 0)
					tscore = score::make(score::WIN, -1);
				else if (draws > 0)
					tscore = score::make(score::DRAW, 1);
				else if (losses == stones)
					tscore = score::make(score::LOSS, 1);
				if (lane == 0 and tscore != score::kDefault)
				{
					task.score = tscore;
					task.win = (score::pv(tscore) == score::WIN) ? 1.0f : 0.0f;
					task.draw = (score::pv(tscore) == score::DRAW) ? 1.0f : 0.0f;
				}
				}
				else
				{ // processed by the solver (K5): AlphaBetaSearch::solve outputs (AlphaBetaSearch.cpp:114-135), then what the network
				  // adds (NNEvaluator.cpp:263-286), then UnifiedGenerator::generate without terminal checks (EdgeGenerator.cpp:269-303)
					tscore = p.s.solver_out.score[slot];
					must_defend = p.s.solver_out.must_defend[slot] != 0;
					count = p.s.solver_out.n_actions[slot];
					const bool proven = score::is_proven(tscore);
					const bool went_to_nn = (task.path_len == 0) or not proven;
					if (lane == 0)
					{
						task.score = tscore;
						if (proven)
						{ // Score::convertToValue, distance as moves left
							task.win = (score::pv(tscore) == score::WIN) ? 1.0f : 0.0f;
							task.draw = (score::pv(tscore) == score::DRAW) ? 1.0f : 0.0f;
							task.moves_left = static_cast<float>(score::distance(tscore));
						}
						if (went_to_nn)
						{
							task.win = p.s.value[slot * 3 + 0];
							task.draw = p.s.value[slot * 3 + 1];
							if (not proven)
								task.moves_left = 0.0f;
							atomicAdd(p.s.stats + ST_EVALS, 1ull);
						}
						if (p.s.solver_out.nodes[slot] <= 1)
							task.sticky |= 1; // node_counter <= 1: statically solved
						if (proven)
							task.sticky |= 2; // (sic) a proven result is flagged as recursively solved
					}
					const uint16_t *am = p.s.solver_out.moves + static_cast<size_t>(slot) * p.s.solver_out.pitch;
					const uint16_t *as = p.s.solver_out.scores + static_cast<size_t>(slot) * p.s.solver_out.pitch;
					for (int i = lane; i < count; i += 32)
					{
						EdgeD e;
						const uint16_t mv = am[i];
						const int cell = ((mv >> 2) & 127) * S + ((mv >> 9) & 127);
						e.move = mv;
						e.score = as[i];
						e.visits = 0;
						e.vloss_flag = 0;
						e.pad = 0;
						// a task that skipped the network has an all-zero policy (SearchTask::set clears it)
						e.prior = went_to_nn ? tempered_prior(p.s.policy[static_cast<size_t>(slot) * cells + cell], p.policy_temperature, max_policy)
								: tempered_prior(0.0f, p.policy_temperature, 0.0f);
						e.win = 0.0f;
						e.draw = 0.0f;
						if (went_to_nn)
						{ // the network's action values replace the solver's for the whole board (NNEvaluator.cpp:279)
							if (p.q_head)
							{
								e.win = p.s.q[(static_cast<size_t>(slot) * cells + cell) * 3 + 0];
								e.draw = p.s.q[(static_cast<size_t>(slot) * cells + cell) * 3 + 1];
							}
						}
						else if (score::is_proven(e.score))
						{
							e.win = (score::pv(e.score) == score::WIN) ? 1.0f : 0.0f;
							e.draw = (score::pv(e.score) == score::DRAW) ? 1.0f : 0.0f;
						}
						edges[n_edges + i] = e;
					}
					__syncwarp();
				}
				// prune_weak_moves (proven position, not the root): keep the best-scoring edges in their order
				const bool is_root_task = (task.path_len == 0);
				if (score::is_proven(tscore) and not is_root_task)
				{
					uint16_t best = 0; // Score::loss() is (LOSS, 0) = 4000
					best = score::make(score::LOSS, 0);
					for (int i = lane; i < count; i += 32)
						best = max(best, edges[n_edges + i].score);
					for (int o = 16; o > 0; o >>= 1)
						best = max(best, static_cast<uint16_t>(__shfl_xor_sync(kFullMask, static_cast<int>(best), o)));
					int kept = 0;
					for (int i0 = 0; i0 < count; i0 += 32)
					{
						const int i = i0 + lane;
						EdgeD e;
						bool keep = false;
						if (i < count)
						{
							e = edges[n_edges + i];
							keep = (e.score == best);
						}
						const unsigned km = __ballot_sync(kFullMask, keep);
						__syncwarp();
						if (keep)
							edges[n_edges + kept + __popc(km & ((1u << lane) - 1u))] = e;
						kept += __popc(km);
						__syncwarp();
					}
					count = kept;
				}
				else if (not is_root_task and p.max_children > 0 and count > p.max_children and not must_defend)
				{ // prune_weak_moves, unproven position: the max_children best edges, then those above the expansion threshold
					if (lane == 0)
					{
						EdgeD *first = edges + n_edges;
						partial_sort_edges(first, p.max_children, count);
						float sum_policy = 0.0f;
						for (int i = 0; i < p.max_children; i++)
							sum_policy += first[i].prior;
						const float threshold = p.expansion_threshold * sum_policy;
						int kept = 0;
						for (int i = 0; i < p.max_children; i++)
							if (first[i].prior >= threshold)
								kept++;
						count = kept;
					}
					count = __shfl_sync(kFullMask, count, 0);
					__syncwarp();
				}
				// renormalize_policy: sequential float sum in edge order (EdgeGenerator.cpp:23-40)
				if (lane == 0 and count > 0)
				{
					float sum = 0.0f;
					for (int i = 0; i < count; i++)
						sum += edges[n_edges + i].prior;
					if (sum == 0.0f)
					{
						const float u = 1.0f / count;
						for (int i = 0; i < count; i++)
							edges[n_edges + i].prior = u;
					}
					else
					{
						const float inv = 1.0f / sum;
						for (int i = 0; i < count; i++)
							edges[n_edges + i].prior *= inv;
					}
				}
				__syncwarp();
				// Tree::expand
				if (count == 0)
					continue; // ExpandOutcome::SKIPPED_EXPANSION
				const int existing = table_seek(p, g, task.hash, task.bits, stm, lane);
				if (existing < 0)
				{
					if (n_nodes >= p.s.max_nodes)
					{
						atomicOr(p.status, OVF_NODES);
						continue;
					}
					NodeD node;
					node.edge_begin = n_edges;
					node.n_edges = static_cast<int16_t>(count);
					node.depth = static_cast<int16_t>(stones);
					node.win = node.draw = node.moves_left = 0.0f;
					node.visits = 0;
					node.score = score::kDefault;
					node.vloss = 0;
					node.stm = static_cast<int8_t>(stm);
					node.flags = 0;
					node.pad = 0;
					node.hash = task.hash;
					node_update_value(node, task.win, task.draw);
					node.moves_left += (task.moves_left - node.moves_left) / node.visits;
					if (must_defend or count + stones == cells)
						node.flags |= 2; // fully expanded
					if (p.solver_mode != 0)
						node.flags |= static_cast<uint8_t>(((task.sticky & 3) << 2) | (must_defend ? 16 : 0)); // Node::setAdditionaFlags
					if (is_root_task)
						node.flags |= 1;
					update_node_score(node, edges, lane);
					__syncwarp();
					if (lane == 0)
					{
						nodes[n_nodes] = node;
						table_insert(p, g, task.hash, n_nodes);
						task.final_node = n_nodes;
						if (is_root_task)
							p.s.root_node[g] = n_nodes;
						p.s.n_nodes[g] = n_nodes + 1;
						p.s.n_edges[g] = n_edges + count;
					}
					if (lane < kBitWords)
						p.s.node_bits[(static_cast<size_t>(g) * p.s.max_nodes + n_nodes) * kBitWords + lane] = task.bits[lane];
					__syncwarp();
				}
				else
				{ // the same position was reached along another path (or by an earlier task of this batch)
					if (lane == 0)
					{
						task.final_node = existing;
						atomicAdd(p.s.stats + ST_WASTED, 1ull);
					}
					__syncwarp();
					if (task.path_len > 0 and has_information_leak(edges[task.path_edge[task.path_len - 1]], &nodes[existing], p.leak_threshold))
						correct_information_leak(p, g, task, lane);
				}
				__syncwarp();
			}

			// Tree::backup for every stored task, in order (Search::backup)
			for (int t = 0; t < stored; t++)
			{
				const TaskD &task = tasks[t];
				if (not task.stored)
					continue;
				if (lane == 0)
					atomicAdd(p.s.stats + ST_NODES, 1ull);
				float moves_left = task.moves_left;
				for (int i = task.path_len - 1; i >= 0; i--)
				{
					NodeD node = nodes[task.path_node[i]];
					EdgeD edge = edges[task.path_edge[i]];
					const int next = (i == task.path_len - 1) ? task.final_node : task.path_node[i + 1];
					float w = task.win, d = task.draw;
					if (node.stm != task.stm)
						w = 1.0f - (task.win + task.draw); // Value::getInverted
					node_update_value(node, w, d);
					edge_update_value(edge, w, d);
					node.moves_left += (moves_left - node.moves_left) / node.visits;
					moves_left += 1.0f;
					if (next >= 0)
						edge.score = score::invert_up(nodes[next].score);
					node.vloss -= 1;
					edge.vloss_flag = (edge.vloss_flag & 0x7FFF) - 1; // decreaseVirtualLoss + clearFlags
					__syncwarp();
					if (lane == 0)
						edges[task.path_edge[i]] = edge;
					__syncwarp();
					update_node_score(node, edges, lane);
					if (lane == 0)
						nodes[task.path_node[i]] = node;
					__syncwarp();
				}
			}
			if (lane == 0)
				p.s.n_stored[g] = 0;
		}

		// A game that starts from a position has no move order for the stones already on the board; like Game::loadOpening
		// (src/game/Game.cpp:58-63) the record lists them as moves: cross and circle stones in row-major order, interleaved.
		__device__ int opening_moves(const int8_t *board, int cells, int S, uint16_t *moves)
		{
			int n = 0, ci = 0, oi = 0;
			while (true)
			{
				while (ci < cells and board[ci] != CROSS)
					ci++;
				while (oi < cells and board[oi] != CIRCLE)
					oi++;
				if (ci >= cells and oi >= cells)
					break;
				if (ci < cells)
				{
					moves[n++] = static_cast<uint16_t>(CROSS | ((ci / S) << 2) | ((ci % S) << 9));
					ci++;
				}
				if (oi < cells)
				{
					moves[n++] = static_cast<uint16_t>(CIRCLE | ((oi / S) << 2) | ((oi % S) << 9));
					oi++;
				}
			}
			return n;
		}

		// ---- final move selectors (EdgeSelector.cpp:426-536, 1340-1433): the value find_best_edge maximises, first maximum wins -----------------
		__device__ float final_selector_value(const Params &p, const EdgeD &e, const NodeD &parent, float parent_log_visit)
		{
			const int proven = score::pv(e.score);
			const float distance = static_cast<float>(score::distance(e.score));
			const float expectation = e.win + 0.5f * e.draw;
			const int vloss = e.vloss_flag & 0x7FFF;
			switch (p.final_selector)
			{
				default:
				case AGB_FINAL_MAX_VISIT:
					return static_cast<float>(e.visits);
				case AGB_FINAL_MIN_VISIT:
					return static_cast<float>(-e.visits);
				case AGB_FINAL_MAX_POLICY:
					return e.prior;
				case AGB_FINAL_MAX_VALUE:
					if (proven == score::LOSS)
						return -1000.0f + distance;
					if (proven == score::DRAW)
						return 0.5f;
					if (proven == score::WIN)
						return +1000.0f - distance;
					return expectation;
				case AGB_FINAL_BEST:
					if (proven == score::LOSS)
						return -1.0e8f + distance;
					if (proven == score::WIN)
						return +1.0e8f - distance;
					return static_cast<float>(e.visits) + expectation * static_cast<float>(parent.visits) + 0.001f * e.prior;
				case AGB_FINAL_LCB:
				{
					if (proven == score::LOSS)
						return -1.0e6f + distance + e.prior;
					if (proven == score::WIN)
						return +1.0e6f - distance + e.prior;
					const float q = (e.visits > 0) ? expectation : (parent.win + 0.5f * parent.draw);
					const float u = p.final_exploration * sqrtf(parent_log_visit / (1.0f + static_cast<float>(e.visits) + static_cast<float>(vloss)));
					const float visits = 1.0e-8f + static_cast<float>(e.visits); // getVirtualLoss(edge), EdgeSelector.cpp:26-31
					return q * (visits / (visits + static_cast<float>(vloss))) - u;
				}
			}
		}

		// Tree::setBoard -> NodeCache::cleanup (src/search/monte_carlo/Tree.cpp:128-151, NodeCache.cpp): keep every node whose position can still
		// occur from the new root position `bits` (all its stones are on the node's board), compact the kept nodes and their edge blocks in place,
		// rebuild the hash table and look the new root up. The warp of game g runs it: after every move of a self-play game (prepare_search) and when
		// a player is handed a new position (Player::setBoard, agb_think).
		__device__ void keep_possible_nodes(const Params &p, int g, int lane, const uint64_t *bits)
		{
			NodeD *nodes = p.s.nodes + static_cast<size_t>(g) * p.s.max_nodes;
			EdgeD *edges = p.s.edges + static_cast<size_t>(g) * p.s.max_edges;
			// prepare_search -> Tree::setBoard -> NodeCache::cleanup: keep every node whose position can still occur
			const int n_nodes = p.s.n_nodes[g];
			int32_t *remap = p.s.remap + static_cast<size_t>(g) * p.s.max_nodes;
			uint64_t *node_bits = p.s.node_bits + static_cast<size_t>(g) * p.s.max_nodes * kBitWords;
			uint64_t rb[kBitWords];
			for (int k = 0; k < kBitWords; k++)
				rb[k] = bits[k];
			int kept = 0;
			for (int i0 = 0; i0 < n_nodes; i0 += 32)
			{
				const int i = i0 + lane;
				bool keep = false;
				if (i < n_nodes)
				{
					keep = true;
					for (int k = 0; k < kBitWords; k++)
						keep = keep and ((node_bits[static_cast<size_t>(i) * kBitWords + k] & rb[k]) == rb[k]);
				}
				const unsigned km = __ballot_sync(kFullMask, keep);
				if (i < n_nodes)
					remap[i] = keep ? kept + __popc(km & ((1u << lane) - 1u)) : -1;
				kept += __popc(km);
			}
			__syncwarp();
			// compact nodes (ascending, destinations never overtake sources) and their edge blocks
			int edge_cursor = 0;
			for (int i = 0; i < n_nodes; i++)
			{
				const int dst = remap[i];
				if (dst < 0)
					continue;
				NodeD node = nodes[i];
				uint64_t nbits = (lane < kBitWords) ? node_bits[static_cast<size_t>(i) * kBitWords + lane] : 0;
				const int old_begin = node.edge_begin;
				for (int e0 = 0; e0 < node.n_edges; e0 += 32)
				{
					EdgeD e;
					if (e0 + lane < node.n_edges)
						e = edges[old_begin + e0 + lane];
					__syncwarp();
					if (e0 + lane < node.n_edges)
						edges[edge_cursor + e0 + lane] = e;
					__syncwarp();
				}
				node.edge_begin = edge_cursor;
				node.flags &= ~1; // root mark is re-applied below
				edge_cursor += node.n_edges;
				__syncwarp();
				if (lane == 0)
					nodes[dst] = node;
				if (lane < kBitWords)
					node_bits[static_cast<size_t>(dst) * kBitWords + lane] = nbits;
				__syncwarp();
			}
			int32_t *table = p.s.table + static_cast<size_t>(g) * p.s.table_size;
			for (int i = lane; i < p.s.table_size; i += 32)
				table[i] = -1;
			__syncwarp();
			if (lane == 0)
			{
				for (int i = 0; i < kept; i++)
					table_insert(p, g, nodes[i].hash, i);
				p.s.n_nodes[g] = kept;
				p.s.n_edges[g] = edge_cursor;
			}
			__syncwarp();
			const int new_root = table_seek(p, g, p.s.root_hash[g], bits, p.s.root_stm[g], lane);
			if (lane == 0)
			{
				p.s.root_node[g] = new_root;
				if (new_root >= 0)
					nodes[new_root].flags |= 1;
			}
		}

		// ---- per-ply driver: final move, record, game end, subtree reuse ------------------------------------------------------
		__global__ void __launch_bounds__(128) make_move_kernel(const __grid_constant__ Params p)
		{
			__shared__ int8_t sboards[4][kCellPitch];
			const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
			if (blockIdx.x * 4 + warp >= p.game_count)
				return;
			const int g = p.game_begin + blockIdx.x * 4 + warp;
			const int cells = p.s.cells, S = p.s.S;
			NodeD *nodes = p.s.nodes + static_cast<size_t>(g) * p.s.max_nodes;
			EdgeD *edges = p.s.edges + static_cast<size_t>(g) * p.s.max_edges;
			const int root = p.s.root_node[g];
			// paused 2: the game is over but its record did not fit the finished-game queue; it waits, with its record, until the host has
			// popped (agb_pop_finished) and is published by the first make-move pass that finds room
			const bool republish = p.s.paused[g] == 2;
			if (not republish and (root < 0 or p.s.paused[g]))
				return;
			int8_t *board = p.s.root_board + static_cast<size_t>(g) * cells;
			uint64_t *bits = p.s.root_bits + static_cast<size_t>(g) * kBitWords;
			int outcome = republish ? p.s.outcome[g] : 0;
			if (not republish)
			{
			const NodeD R = nodes[root];
			// GameGenerator.cpp:95-101: fewer simulations when the root is drawish
			const float reduction = fminf(1.0f, fmaxf(0.0f, (R.draw - 0.75f) / (1.0f - 0.75f)));
			const int simulations = static_cast<int>(p.max_simulations - reduction * (p.max_simulations - 50));
			if (not (R.visits > simulations or score::is_proven(R.score)))
				return;
			// final selector: the first edge with the largest value (find_best_edge_impl, EdgeSelector.cpp:561-586)
			const float parent_log_visit = static_cast<float>(log(static_cast<double>(R.visits + R.vloss)));
			float best_v = -FLT_MAX;
			int best_i = 0x7FFFFFFF;
			for (int i = lane; i < R.n_edges; i += 32)
			{
				const float v = final_selector_value(p, edges[R.edge_begin + i], R, parent_log_visit);
				if (v > best_v)
				{
					best_v = v;
					best_i = i;
				}
			}
			for (int o = 16; o > 0; o >>= 1)
			{
				const float ov = __shfl_xor_sync(kFullMask, best_v, o);
				const int oi = __shfl_xor_sync(kFullMask, best_i, o);
				if (ov > best_v or (ov == best_v and oi < best_i))
				{
					best_v = ov;
					best_i = oi;
				}
			}
			const EdgeD chosen = edges[R.edge_begin + best_i];
			if (p.single_move)
			{ // Player::getMove: report the move and the root's value, leave the position to the host
				if (lane == 0)
				{
					p.s.decision[g] = chosen.move;
					p.s.sample_root[g * 4 + 0] = R.win;
					p.s.sample_root[g * 4 + 1] = R.draw;
					p.s.sample_root[g * 4 + 2] = static_cast<float>(R.visits);
					p.s.sample_root[g * 4 + 3] = static_cast<float>(chosen.move);
					p.s.paused[g] = 1;
				}
				return;
			}
			// SearchDataPack(rootNode, board) (dataset/data_packs.cpp:24-43), kept dense for the record writer
			for (int i = lane; i < cells; i += 32)
			{
				p.s.sample_visits[static_cast<size_t>(g) * cells + i] = 0;
				p.s.sample_prior[static_cast<size_t>(g) * cells + i] = 0.0f;
				p.s.sample_win[static_cast<size_t>(g) * cells + i] = 0.0f;
				p.s.sample_draw[static_cast<size_t>(g) * cells + i] = 0.0f;
				p.s.sample_score[static_cast<size_t>(g) * cells + i] = score::kDefault;
			}
			__syncwarp();
			for (int i = lane; i < R.n_edges; i += 32)
			{
				const EdgeD e = edges[R.edge_begin + i];
				const int cell = ((e.move >> 2) & 127) * S + ((e.move >> 9) & 127);
				p.s.sample_visits[static_cast<size_t>(g) * cells + cell] = e.visits;
				p.s.sample_prior[static_cast<size_t>(g) * cells + cell] = e.prior;
				p.s.sample_win[static_cast<size_t>(g) * cells + cell] = e.win;
				p.s.sample_draw[static_cast<size_t>(g) * cells + cell] = e.draw;
				p.s.sample_score[static_cast<size_t>(g) * cells + cell] = e.score;
			}
			if (lane == 0)
			{
				p.s.sample_root[g * 4 + 0] = R.win;
				p.s.sample_root[g * 4 + 1] = R.draw;
				p.s.sample_root[g * 4 + 2] = static_cast<float>(R.visits);
				p.s.sample_root[g * 4 + 3] = static_cast<float>(chosen.move);
				atomicAdd(p.s.stats + ST_MOVES, 1ull);
			}
			// game_data_storage.addSample(sample): quantise and append this ply (K8)
			__syncwarp();
			if (lane == 0)
			{
				uint8_t *rec = p.s.rec_buf + static_cast<size_t>(g) * p.s.rec_cap;
				int len = p.s.rec_len[g];
				if (len + static_cast<int>(records::max_sample_bytes(cells)) > p.s.rec_cap)
					atomicOr(p.status, OVF_RECORD);
				else
				{
					const size_t base = static_cast<size_t>(g) * cells;
					len += static_cast<int>(records::serialize_sample_v201(rec + len, cells, p.s.root_board + base, p.s.sample_visits + base, p.s.sample_prior + base,
							p.s.sample_win + base, p.s.sample_draw + base, p.s.sample_score + base, R.score, static_cast<uint16_t>((R.flags >> 2) & 7)));
					p.s.rec_len[g] = len;
					p.s.rec_samples[g] += 1;
				}
			}
			__syncwarp();
			// game.makeMove
			const int row = (chosen.move >> 2) & 127, col = (chosen.move >> 9) & 127, sign = chosen.move & 3;
			const int cell = row * S + col;
			__syncwarp();
			if (lane == 0)
			{
				board[cell] = static_cast<int8_t>(sign);
				bits[(sign - 1) * kWordsPerColour + (cell >> 6)] |= 1ull << (cell & 63);
				p.s.root_hash[g] ^= p.s.zobrist[cell * 2 + sign - 1] ^ p.s.zobrist[cells * 2] ^ p.s.zobrist[cells * 2 + 1];
				p.s.root_stm[g] = static_cast<int8_t>(3 - sign);
				p.s.moves[static_cast<size_t>(g) * cells + p.s.n_moves[g]] = chosen.move;
				p.s.n_moves[g] += 1;
			}
			__syncwarp();
			for (int i = lane; i < cells; i += 32)
				sboards[warp][i] = board[i];
			__syncwarp();
			if (lane == 0)
			{
				bool overflow = false;
				outcome = outcome_of(sboards[warp], S, p.rules, p.draw_after, row, col, sign, p.tables, overflow);
				if (overflow)
					atomicOr(p.status, 1u);
			}
			outcome = __shfl_sync(kFullMask, outcome, 0);
			if (outcome != 0 and lane == 0)
			{
				p.s.outcome[g] = static_cast<int8_t>(outcome);
				atomicAdd(p.s.stats + ST_GAMES, 1ull);
			}
			} // not republish
			if (outcome != 0)
			{ // game over: publish the outcome and start the next game from the openings pool (or an empty board)
				// GameDataStorage::serialize (GameDataStorage.cpp:217-251): u32 samples, samples, u32 moves + u16[], outcome, rows = cols = 0
				// (GameGenerator default-constructs its storage, GameGenerator.hpp:39); games without samples are not recorded
				const int n_samples = p.s.rec_samples[g];
				if (n_samples > 0)
				{
					const int n_moves = p.s.n_moves[g];
					const int body = p.s.rec_len[g];
					const unsigned long long total = static_cast<unsigned long long>(body) + 4 + 2ull * n_moves + 12;
					unsigned long long dst_off = ~0ull;
					if (lane == 0)
					{ // reserve room, or none at all (the counter never runs past the capacity)
						unsigned long long seen = *p.s.fin_used;
						while (seen + total <= p.s.fin_cap)
						{
							const unsigned long long prev = atomicCAS(p.s.fin_used, seen, seen + total);
							if (prev == seen)
							{
								dst_off = seen;
								break;
							}
							seen = prev;
						}
					}
					dst_off = __shfl_sync(kFullMask, dst_off, 0);
					if (dst_off == ~0ull)
					{ // queue full: keep the finished game and its record as they are and wait for the host to pop (reported once, nothing is lost)
						if (lane == 0)
						{
							p.s.paused[g] = 2;
							atomicOr(p.status, OVF_FINISHED);
						}
						return;
					}
					else
					{
						uint8_t *dst = p.s.fin_buf + dst_off;
						const uint8_t *rec = p.s.rec_buf + static_cast<size_t>(g) * p.s.rec_cap;
						for (int i = lane; i < body; i += 32)
							dst[i] = (i < 4) ? static_cast<uint8_t>((static_cast<uint32_t>(n_samples) >> (8 * i)) & 0xFF) : rec[i];
						for (int i = lane; i < n_moves; i += 32)
						{
							const uint16_t mv = p.s.moves[static_cast<size_t>(g) * cells + i];
							dst[body + 4 + 2 * i] = static_cast<uint8_t>(mv & 0xFF);
							dst[body + 4 + 2 * i + 1] = static_cast<uint8_t>(mv >> 8);
						}
						if (lane == 0)
						{
							size_t off = body;
							records::put32(dst, off, static_cast<uint32_t>(n_moves));
							off += 2 * static_cast<size_t>(n_moves);
							records::put32(dst, off, static_cast<uint32_t>(outcome));
							records::put32(dst, off, 0u);
							records::put32(dst, off, 0u);
							atomicAdd(p.s.fin_games, 1);
						}
					}
				}
				__syncwarp();
				int8_t next_stm = CROSS;
				const int8_t *src = nullptr;
				if (p.s.n_openings > 0)
				{
					// which opening comes next depends only on the game's global id and on how often it has restarted: reproducible whatever the
					// scheduling, the stream layout or the sharding
					unsigned k = 0;
					if (lane == 0)
					{
						unsigned long long z = p.sym_seed ^ 0x6F70656E696E67ull ^ (static_cast<unsigned long long>(p.first_game_id + g) << 32) ^ p.s.opening_cursor[g]++;
						z += 0x9E3779B97F4A7C15ull;
						z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
						z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
						k = static_cast<unsigned>((z ^ (z >> 31)) % static_cast<unsigned long long>(p.s.n_openings));
					}
					k = __shfl_sync(kFullMask, k, 0);
					src = p.s.openings + static_cast<size_t>(k) * cells;
					next_stm = p.s.opening_stm[k];
				}
				uint64_t h = 0;
				__syncwarp();
				for (int i = lane; i < cells; i += 32)
					board[i] = src ? src[i] : 0;
				__syncwarp();
				if (lane < kBitWords)
				{
					uint64_t w = 0;
					const int colour = lane / kWordsPerColour, word = lane % kWordsPerColour;
					for (int i = word * 64; i < min(cells, word * 64 + 64); i++)
						if (board[i] == colour + 1)
						{
							w |= 1ull << (i & 63);
							h ^= p.s.zobrist[i * 2 + colour];
						}
					bits[lane] = w;
				}
				for (int o = 8; o > 0; o >>= 1) // lanes kBitWords..15 contribute zero
					h ^= __shfl_xor_sync(kFullMask, h, o);
				if (lane == 0)
				{
					p.s.root_hash[g] = h ^ p.s.zobrist[cells * 2 + next_stm - 1];
					p.s.root_stm[g] = next_stm;
					p.s.n_moves[g] = opening_moves(board, cells, S, p.s.moves + static_cast<size_t>(g) * cells);
					p.s.root_node[g] = -1;
					p.s.n_nodes[g] = 0;
					p.s.n_edges[g] = 0;
					p.s.rec_len[g] = 4;
					p.s.rec_samples[g] = 0;
					p.s.paused[g] = 0;
					if (p.s.noise_ready != nullptr)
						p.s.noise_ready[g] = 0;
					if (p.solver_mode != 0)
						p.s.solver.generation[g] = (p.s.solver.generation[g] + 1) % 64;
				}
				int32_t *table = p.s.table + static_cast<size_t>(g) * p.s.table_size;
				for (int i = lane; i < p.s.table_size; i += 32)
					table[i] = -1;
				return;
			}
			if (lane == 0 and p.s.noise_ready != nullptr)
				p.s.noise_ready[g] = 0; // prepare_search makes a new EdgeSelector: the next search draws new noise
			// prepare_search -> Search::setBoard: the solver's table enters a new generation
			if (lane == 0 and p.solver_mode != 0)
				p.s.solver.generation[g] = (p.s.solver.generation[g] + 1) % 64;
			keep_possible_nodes(p, g, lane, bits);
		}

		// Player::setBoard (src/evaluation/Player.cpp:98-108) for every game: a new position arrives from the host, the tree keeps what can still
		// occur (the subtree under the moves played since the last search) and the solver's table enters a new generation. active[g] == 0: the
		// game sits this round out.
		__global__ void __launch_bounds__(128) rebase_games_kernel(const __grid_constant__ Params p, const int8_t *__restrict__ boards, const int8_t *__restrict__ stm,
				const int8_t *__restrict__ active)
		{
			const int lane = threadIdx.x & 31;
			const int g = blockIdx.x * 4 + (threadIdx.x >> 5);
			if (g >= p.s.games)
				return;
			const int cells = p.s.cells;
			if (lane == 0)
			{
				p.s.paused[g] = active[g] ? 0 : 1;
				p.s.decision[g] = 0;
				p.s.n_stored[g] = 0;
			}
			if (not active[g])
				return;
			int8_t *board = p.s.root_board + static_cast<size_t>(g) * cells;
			uint64_t *bits = p.s.root_bits + static_cast<size_t>(g) * kBitWords;
			for (int i = lane; i < cells; i += 32)
				board[i] = boards[static_cast<size_t>(g) * cells + i];
			__syncwarp();
			uint64_t h = 0;
			if (lane < kBitWords)
			{
				uint64_t w = 0;
				const int colour = lane / kWordsPerColour, word = lane % kWordsPerColour;
				for (int i = word * 64; i < min(cells, word * 64 + 64); i++)
					if (board[i] == colour + 1)
					{
						w |= 1ull << (i & 63);
						h ^= p.s.zobrist[i * 2 + colour];
					}
				bits[lane] = w;
			}
			for (int o = 8; o > 0; o >>= 1) // lanes kBitWords..15 contribute zero
				h ^= __shfl_xor_sync(kFullMask, h, o);
			h = __shfl_sync(kFullMask, h, 0);
			if (lane == 0)
			{
				p.s.root_stm[g] = stm[g];
				p.s.root_hash[g] = h ^ p.s.zobrist[cells * 2 + stm[g] - 1];
				p.s.outcome[g] = 0;
				if (p.s.noise_ready != nullptr)
					p.s.noise_ready[g] = 0; // a new EdgeSelector per setBoard: the next search draws new noise
				if (p.solver_mode != 0)
					p.s.solver.generation[g] = (p.s.solver.generation[g] + 1) % 64; // Search::setBoard -> increaseGeneration
			}
			__syncwarp();
			keep_possible_nodes(p, g, lane, bits);
		}

		__global__ void reset_games_kernel(const __grid_constant__ Params p, int keep_history)
		{ // Zobrist hash + bitboards of every game's root position; keep_history: resumed games keep their move lists and samples
			const int g = blockIdx.x * blockDim.x + threadIdx.x;
			if (g >= p.s.games)
				return;
			const int cells = p.s.cells;
			uint64_t bits[kBitWords] = { };
			uint64_t h = 0;
			for (int i = 0; i < cells; i++)
			{
				const int v = p.s.root_board[static_cast<size_t>(g) * cells + i];
				if (v == CROSS or v == CIRCLE)
				{
					bits[(v - 1) * kWordsPerColour + (i >> 6)] |= 1ull << (i & 63);
					h ^= p.s.zobrist[i * 2 + v - 1];
				}
			}
			for (int k = 0; k < kBitWords; k++)
				p.s.root_bits[static_cast<size_t>(g) * kBitWords + k] = bits[k];
			p.s.root_hash[g] = h ^ p.s.zobrist[cells * 2 + p.s.root_stm[g] - 1];
			p.s.root_node[g] = -1;
			p.s.n_nodes[g] = 0;
			p.s.n_edges[g] = 0;
			p.s.n_stored[g] = 0;
			p.s.outcome[g] = 0;
			if (p.s.noise_ready != nullptr)
				p.s.noise_ready[g] = 0;
			if (not keep_history)
			{
				p.s.n_moves[g] = opening_moves(p.s.root_board + static_cast<size_t>(g) * cells, cells, p.s.S, p.s.moves + static_cast<size_t>(g) * cells);
				p.s.rec_len[g] = 4;
				p.s.rec_samples[g] = 0;
			}
			if (p.solver_mode != 0)
				p.s.solver.generation[g] = (p.s.solver.generation[g] + 1) % 64; // prepare_search -> Search::setBoard -> increaseGeneration
		}

		uint64_t splitmix64(uint64_t &x)
		{
			uint64_t z = (x += 0x9E3779B97F4A7C15ull);
			z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
			z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
			return z ^ (z >> 31);
		}
		Params make_params(AgbEngine *e)
		{
			Params p;
			p.s = *e->selfplay;
			p.rules = e->cfg.rules;
			p.draw_after = e->cfg.draw_after > 0 ? e->cfg.draw_after : e->cells;
			p.max_simulations = e->cfg.max_simulations;
			p.init_to = e->cfg.init_to;
			p.exploration_constant = e->cfg.exploration_constant;
			p.leak_threshold = e->cfg.information_leak_threshold;
			p.q_head = e->cfg.q_head;
			p.solver_mode = e->cfg.solver_max_positions > 0 ? 1 : 0;
			p.game_begin = 0;
			p.game_count = e->selfplay->games;
			p.slot_base = 0;
			p.eval_count = e->selfplay->eval_count;
			p.use_symmetries = e->cfg.use_symmetries != 0;
			p.max_children = e->cfg.max_children > 0 ? e->cfg.max_children : 0;
			p.single_move = e->think_mode ? 1 : 0;
			// 0 in the (possibly zero-initialised) config means the default 1; a negative value asks for the reference's temperature 0
			p.policy_temperature = (e->cfg.policy_temperature == 0.0f) ? 1.0f : ((e->cfg.policy_temperature < 0.0f) ? 0.0f : e->cfg.policy_temperature);
			p.final_selector = e->cfg.final_selector;
			p.noise_type = (e->cfg.noise_weight > 0.0f) ? e->cfg.noise_type : 0;
			p.noise_weight = e->cfg.noise_weight;
			p.final_exploration = e->cfg.final_exploration_constant;
			p.expansion_threshold = e->cfg.policy_expansion_threshold;
			p.first_game_id = e->cfg.first_game_id;
			p.sym_seed = e->cfg.seed * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull;
			p.sym = SymmetryStream { p.s.task_sym, p.s.sym_counter, p.sym_seed, e->cfg.first_game_id, p.s.sym_table, p.s.sym_table_size };
			p.tables = e->tables;
			p.store = e->store;
			p.status = e->d_status;
			return p;
		}
	}

	int net_forward_dev_gather(AgbEngine *e, const uint32_t *features_dev, const int *count_dev, const int *gather_dev, int max_boards, float *policy_dev,
			float *value_dev, float *q_dev, int slot_base, cudaStream_t stream, int max_sms, cudaStream_t tail_stream, cudaEvent_t trunk_done);

	int selfplay_create(AgbEngine *e)
	{
		const AgbConfig &c = e->cfg;
		if (c.blocks <= 0)
			return e->fail(AGB_EINVAL, "self-play needs a network (blocks > 0)");
		if (c.max_batch_size <= 0 or c.games * c.max_batch_size > c.max_boards)
			return e->fail(AGB_EINVAL, "games * max_batch_size must fit in max_boards");
		if (c.max_simulations < 50) // the reference asserts it (get_simulations_for_move, src/utils/misc.cpp:171-179): a drawish root is searched for
			return e->fail(AGB_EINVAL, "max_simulations must be at least 50 (the simulation budget of a drawish position)"); // 50 simulations, which select never reaches with fewer
		SelfplayState *s = new SelfplayState();
		e->selfplay = s;
		s->games = c.games;
		s->batch = c.max_batch_size;
		s->cells = e->cells;
		s->S = c.rows;
		s->max_nodes = c.max_nodes_per_game > 0 ? c.max_nodes_per_game : 2048;
		s->max_edges = c.max_edges_per_game > 0 ? c.max_edges_per_game : s->max_nodes * 128;
		int ts = 1;
		while (ts < 2 * s->max_nodes)
			ts *= 2;
		s->table_size = ts;
		const size_t G = c.games, cells = e->cells, T = G * c.max_batch_size;
		bool ok = true;
		const auto alloc = [&](auto **ptr, size_t count)
		{
			ok = ok and cudaMalloc(reinterpret_cast<void**>(ptr), count * sizeof(**ptr)) == cudaSuccess;
		};
		alloc(&s->root_board, G * cells);
		alloc(&s->root_bits, G * kBitWords);
		alloc(&s->root_hash, G);
		alloc(&s->root_stm, G);
		alloc(&s->root_node, G);
		alloc(&s->n_nodes, G);
		alloc(&s->n_edges, G);
		alloc(&s->n_stored, G);
		alloc(&s->n_moves, G);
		alloc(&s->moves, G * cells);
		alloc(&s->outcome, G);
		alloc(&s->nodes, G * s->max_nodes);
		alloc(&s->node_bits, G * s->max_nodes * kBitWords);
		alloc(&s->edges, G * s->max_edges);
		alloc(&s->table, G * s->table_size);
		alloc(&s->remap, G * s->max_nodes);
		alloc(&s->tasks, T);
		alloc(&s->task_boards, T * cells);
		alloc(&s->task_stm, T);
		alloc(&s->eval_count, kMaxGroups);
		alloc(&s->slot_is_root, T);
		alloc(&s->nn_list, T);
		alloc(&s->nn_count, kMaxGroups);
		s->solver_out.pitch = static_cast<int>(cells);
		alloc(&s->solver_out.moves, T * cells);
		alloc(&s->solver_out.scores, T * cells);
		alloc(&s->solver_out.n_actions, T);
		alloc(&s->solver_out.score, T);
		alloc(&s->solver_out.must_defend, T);
		alloc(&s->solver_out.nodes, T);
		alloc(&s->features, T * cells);
		alloc(&s->policy, T * cells);
		alloc(&s->value, T * 3);
		alloc(&s->q, T * cells * 3);
		alloc(&s->paused, G);
		alloc(&s->decision, G);
		if (c.noise_type != 0 and c.noise_weight > 0.0f)
		{
			alloc(&s->root_noise, G * cells);
			alloc(&s->noise_ready, G);
			alloc(&s->noise_counter, G);
		}
		if (c.use_symmetries)
		{
			alloc(&s->task_sym, T);
			alloc(&s->sym_counter, G);
			alloc(&s->features_aug, T * cells);
			alloc(&s->policy_raw, T * cells);
			if (c.q_head)
				alloc(&s->q_raw, T * cells * 3);
		}
		alloc(&s->zobrist, cells * 2 + 2);
		alloc(&s->stats, 16);
		alloc(&s->opening_cursor, G);
		alloc(&s->sample_visits, G * cells);
		alloc(&s->sample_prior, G * cells);
		alloc(&s->sample_win, G * cells);
		alloc(&s->sample_draw, G * cells);
		alloc(&s->sample_score, G * cells);
		s->rec_cap = 4 + static_cast<int>(records::max_sample_bytes(static_cast<int>(cells))) * (static_cast<int>(cells) + 1);
		alloc(&s->rec_buf, G * s->rec_cap);
		alloc(&s->rec_len, G);
		alloc(&s->rec_samples, G);
		s->fin_cap = static_cast<unsigned long long>(G) * 65536ull + (1ull << 22);
		if (const char *cap = getenv("AGB_FINISHED_QUEUE_BYTES")) // tests of the queue-full path: a queue that a few games fill
			s->fin_cap = std::max(1024ull, std::strtoull(cap, nullptr, 10));
		alloc(&s->fin_buf, s->fin_cap);
		alloc(&s->fin_used, 1);
		alloc(&s->fin_games, 1);
		alloc(&s->sample_root, G * 4);
		if (ok)
			ok = cudaMemset(s->tasks, 0, T * sizeof(TaskD)) == cudaSuccess; // sticky per-slot flags start cleared
		if (not ok)
			return e->fail(AGB_ENOMEM, std::string("self-play arenas: ") + cudaGetErrorString(cudaGetLastError()));
		const bool pipelined = (c.pipeline_groups > 1) or (c.pipeline_groups == 0 and c.solver_max_positions > 1 and c.games >= 1024);
		int sms = 148;
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device);
		// Opt-in (AGB_GREEN_CONTEXTS=1): measured at steady state the three-way split loses to the SM-filling solver blocks (178 k against 188 k
		// evaluations per second): the tree kernels' 8 SMs are mostly idle and K4 does not speed up in proportion to its SMs
		if (pipelined and c.solver_max_positions > 1 and c.solver_sms >= 0 and getenv("AGB_GREEN_CONTEXTS") != nullptr)
		{ // real SM partitions (green contexts); the provisioned sizes are multiples of 8 SMs. Automatic: 64 of 148 for the solver, 8 for the tree kernels
			std::string why;
			const int want = c.solver_sms > 0 ? std::min(c.solver_sms, sms - 24) : sms * 64 / 148;
			s->green = partition_create(c.device, std::max(8, (want + 4) / 8 * 8), 8, &s->partition, &why);
			if (getenv("AGB_VERBOSE") != nullptr)
			{
				if (s->green)
					fprintf(stderr, "agb200: green-context SM partition: solver %d, network %d, tree kernels %d SMs %s\n", s->partition.solver_sms, s->partition.net_sms,
							s->partition.tree_sms, why.c_str());
				else
					fprintf(stderr, "agb200: no green-context SM partition (%s); using SM-filling solver blocks\n", why.c_str());
			}
		}
		// groups: with green contexts the launches of several groups share the solver's SMs, so more groups hide each launch's tail (6: measured
		// best of 2..8 at steady state); without them the solver runs in SM-filling blocks and more than 2 groups only get in each other's way
		// three groups once each still has a thousand games: a group's cycle is solver launch + network launch + tree kernels, and with two groups
		// that chain, not the SMs, bounds the step as soon as both kernels are balanced (steady state: 2 groups 262-270 k, 3 groups 280 k evaluations/s)
		s->groups = c.pipeline_groups > 0 ? c.pipeline_groups : (pipelined ? (s->green ? 6 : (c.games >= 3072 ? 3 : 2)) : 1);
		s->solver_sms = 0;
		if (s->green)
		{
			s->solver_sms = s->partition.solver_sms;
			s->net_sms = s->partition.net_sms;
		}
		else
		{ // AgbConfig::solver_sms
			if (s->groups > 1 and c.solver_max_positions > 1 and c.solver_sms >= 0)
				s->solver_sms = (c.solver_sms > 0 ? std::min(c.solver_sms, sms - 2) : sms * 28 / 148) & ~1;
			s->net_sms = s->solver_sms > 0 ? sms - s->solver_sms : 0;
		}
		if (s->groups > kMaxGroups or s->groups > c.games)
			return e->fail(AGB_EINVAL, "pipeline_groups must be 1..8 and not exceed the number of games");
		if (s->groups > 1)
		{
			for (int k = 0; k < s->groups; k++)
			{
				if (s->green)
				{
					if (not partition_stream(s->partition.tree_ctx, &s->group_stream[k]) or not partition_stream(s->partition.solver_ctx, &s->solver_stream[k]))
						return e->fail(AGB_ECUDA, "cuGreenCtxStreamCreate failed");
					AGB_CUDA_CHECK(e, cudaEventCreateWithFlags(&s->solver_go[k], cudaEventDisableTiming));
					AGB_CUDA_CHECK(e, cudaEventCreateWithFlags(&s->solver_done[k], cudaEventDisableTiming));
				}
				else
					AGB_CUDA_CHECK(e, cudaStreamCreateWithFlags(&s->group_stream[k], cudaStreamNonBlocking));
				AGB_CUDA_CHECK(e, cudaEventCreateWithFlags(&s->ready[k], cudaEventDisableTiming));
				AGB_CUDA_CHECK(e, cudaEventCreateWithFlags(&s->evaluated[k], cudaEventDisableTiming));
			}
			if (s->green)
			{
				if (not partition_stream(s->partition.net_ctx, &s->nn_stream))
					return e->fail(AGB_ECUDA, "cuGreenCtxStreamCreate failed");
				if (getenv("AGB_NET_PER_GROUP") != nullptr) // experiment: K4 launches in the order their inputs become ready instead of round-robin
					for (int k = 0; k < s->groups; k++)
						if (not partition_stream(s->partition.net_ctx, &s->nn_group_stream[k]))
							return e->fail(AGB_ECUDA, "cuGreenCtxStreamCreate failed");
			}
			else
				AGB_CUDA_CHECK(e, cudaStreamCreateWithFlags(&s->nn_stream, cudaStreamNonBlocking));
			AGB_CUDA_CHECK(e, cudaEventCreateWithFlags(&s->joined, cudaEventDisableTiming));
		}
		if (c.solver_max_positions > 0)
		{
			const int rc = solver_state_create(e, c.games, c.max_batch_size, &s->solver);
			if (rc != AGB_OK)
				return rc;
		}
		std::vector<uint64_t> keys(cells * 2 + 2);
		uint64_t seed = c.seed ^ 0xA5A5A5A55A5A5A5Aull;
		for (auto &k : keys)
			k = splitmix64(seed);
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(s->zobrist, keys.data(), keys.size() * 8, cudaMemcpyHostToDevice, e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->stats, 0, 16 * 8, e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->opening_cursor, 0, G * sizeof(uint32_t), e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->paused, 0, G, e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->decision, 0, G * sizeof(uint16_t), e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->fin_used, 0, 8, e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->fin_games, 0, 4, e->stream));
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		return AGB_OK;
	}
	void selfplay_destroy(AgbEngine *e)
	{
		SelfplayState *s = e->selfplay;
		if (s == nullptr)
			return;
		void *ptrs[] = { s->root_board, s->root_bits, s->root_hash, s->root_stm, s->root_node, s->n_nodes, s->n_edges, s->n_stored, s->n_moves, s->moves,
				s->outcome, s->nodes, s->node_bits, s->edges, s->table, s->remap, s->tasks, s->task_boards, s->task_stm, s->eval_count, s->features, s->policy,
				s->value, s->q, s->paused, s->decision, s->root_noise, s->noise_ready, s->noise_counter, s->task_sym, s->sym_counter, s->sym_table, s->features_aug, s->policy_raw, s->q_raw, s->zobrist, s->stats, s->openings, s->opening_stm, s->opening_cursor, s->sample_visits, s->sample_prior, s->sample_win,
				s->sample_root, s->sample_draw, s->sample_score, s->slot_is_root, s->nn_list, s->nn_count, s->solver_out.moves, s->solver_out.scores,
				s->solver_out.n_actions, s->solver_out.score, s->solver_out.must_defend, s->solver_out.nodes, s->rec_buf, s->rec_len, s->rec_samples, s->fin_buf, s->fin_used, s->fin_games };
		for (void *ptr : ptrs)
			if (ptr)
				cudaFree(ptr);
		solver_state_destroy(&s->solver);
		for (int k = 0; k < kMaxGroups; k++)
		{
			if (s->group_stream[k])
				cudaStreamDestroy(s->group_stream[k]);
			if (s->ready[k])
				cudaEventDestroy(s->ready[k]);
			if (s->evaluated[k])
				cudaEventDestroy(s->evaluated[k]);
		}
		if (s->nn_stream)
			cudaStreamDestroy(s->nn_stream);
		for (int k = 0; k < kMaxGroups; k++)
		{
			if (s->nn_group_stream[k])
				cudaStreamDestroy(s->nn_group_stream[k]);
			if (s->solver_stream[k])
				cudaStreamDestroy(s->solver_stream[k]);
			if (s->solver_go[k])
				cudaEventDestroy(s->solver_go[k]);
			if (s->solver_done[k])
				cudaEventDestroy(s->solver_done[k]);
		}
		partition_destroy(&s->partition);
		if (s->joined)
			cudaEventDestroy(s->joined);
		delete s;
		e->selfplay = nullptr;
	}
}

extern "C"
{
	using namespace agb;

	int agb_selfplay_reset(AgbEngine *e, const int8_t *boards_host, const int8_t *sign_to_move_host)
	{
		const agb::DeviceGuard on_device(e);
		SelfplayState *s = e->selfplay;
		if (s == nullptr)
			return e->fail(AGB_ESTATE, "engine was created without games");
		const size_t G = s->games, cells = s->cells;
		if (boards_host != nullptr)
		{
			if (sign_to_move_host == nullptr)
				return e->fail(AGB_EINVAL, "sign_to_move is required with boards");
			const int rc_valid = validate_boards(e, boards_host, sign_to_move_host, G);
			if (rc_valid != AGB_OK)
				return rc_valid;
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(s->root_board, boards_host, G * cells, cudaMemcpyHostToDevice, e->stream));
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(s->root_stm, sign_to_move_host, G, cudaMemcpyHostToDevice, e->stream));
			// the starting positions double as the pool finished games restart from
			if (s->openings == nullptr)
			{
				AGB_CUDA_CHECK(e, cudaMalloc(&s->openings, G * cells));
				AGB_CUDA_CHECK(e, cudaMalloc(&s->opening_stm, G));
			}
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(s->openings, boards_host, G * cells, cudaMemcpyHostToDevice, e->stream));
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(s->opening_stm, sign_to_move_host, G, cudaMemcpyHostToDevice, e->stream));
			s->n_openings = static_cast<int>(G);
		}
		else
		{
			AGB_CUDA_CHECK(e, cudaMemsetAsync(s->root_board, 0, G * cells, e->stream));
			AGB_CUDA_CHECK(e, cudaMemsetAsync(s->root_stm, CROSS, G, e->stream));
			s->n_openings = 0;
		}
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->table, 0xFF, G * s->table_size * sizeof(int32_t), e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->paused, 0, G, e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->decision, 0, G * sizeof(uint16_t), e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->opening_cursor, 0, G * sizeof(uint32_t), e->stream));
		if (s->noise_counter != nullptr)
			AGB_CUDA_CHECK(e, cudaMemsetAsync(s->noise_counter, 0, G * sizeof(uint32_t), e->stream));
		if (s->sym_counter != nullptr)
		{
			AGB_CUDA_CHECK(e, cudaMemsetAsync(s->sym_counter, 0, G * sizeof(uint32_t), e->stream));
			AGB_CUDA_CHECK(e, cudaMemsetAsync(s->task_sym, 0, G * s->batch, e->stream));
		}
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->stats, 0, 16 * 8, e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(e->d_status, 0, 4, e->stream)); // a fresh start: earlier overflows were reported when they happened
		e->overflow_seen = 0;
		if (e->cfg.solver_max_positions > 0)
		{
			const int rc = solver_state_reset(e, &s->solver);
			if (rc != AGB_OK)
				return rc;
		}
		const Params p = make_params(e);
		reset_games_kernel<<<static_cast<unsigned>((G + 127) / 128), 128, 0, e->stream>>>(p, 0);
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		s->reset_done = true;
		return AGB_OK;
	}

	// ---- games in flight: GeneratorManager::saveState / loadState (GeneratorManager.cpp:240-290), GameGenerator::save / load
	// (GameGenerator.cpp:122-141). Like the reference, what survives is each game's position, move list and the samples recorded so
	// far; the search trees are rebuilt (prepare_search). Blob: AgbSavedHeader, then per game: board[cells] int8, sign to move int8,
	// n_moves int32, moves[n_moves] uint16, samples int32, record bytes int32, record[bytes].
	struct AgbSavedHeader
	{
			uint32_t magic, version;
			int32_t games, rows, cols, rules;
	};
	int agb_save_games(AgbEngine *e, void *blob_host, size_t capacity, size_t *used)
	{
		const agb::DeviceGuard on_device(e);
		SelfplayState *s = e->selfplay;
		if (s == nullptr)
			return e->fail(AGB_ESTATE, "engine was created without games");
		if (used == nullptr)
			return e->fail(AGB_EINVAL, "null pointer");
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		const size_t G = s->games, cells = s->cells;
		{ // finished games still on the device would be lost by a save / load round trip: the host pops them first (like the reference, whose
		  // finished games are already in the manager's buffer when saveState runs, GeneratorManager.cpp:240-263)
			unsigned long long pending = 0;
			AGB_CUDA_CHECK(e, cudaMemcpy(&pending, s->fin_used, 8, cudaMemcpyDeviceToHost));
			if (pending != 0)
				return e->fail(AGB_ESTATE, "finished games are waiting on the device: call agb_pop_finished before agb_save_games");
		}
		std::vector<int8_t> boards(G * cells), stm(G);
		std::vector<int32_t> n_moves(G), rec_len(G), rec_samples(G);
		std::vector<uint16_t> moves(G * cells);
		std::vector<uint8_t> rec(G * s->rec_cap);
		// version 2: the per-game random streams (evaluation symmetries, root noise, opening choice) and the openings pool, so that a resumed run
		// continues exactly like an uninterrupted one
		std::vector<uint32_t> sym_counter(G, 0), noise_counter(G, 0), opening_cursor(G, 0);
		std::vector<int8_t> openings(static_cast<size_t>(s->n_openings) * cells), opening_stm(s->n_openings);
		if (s->sym_counter != nullptr)
			AGB_CUDA_CHECK(e, cudaMemcpy(sym_counter.data(), s->sym_counter, G * 4, cudaMemcpyDeviceToHost));
		if (s->noise_counter != nullptr)
			AGB_CUDA_CHECK(e, cudaMemcpy(noise_counter.data(), s->noise_counter, G * 4, cudaMemcpyDeviceToHost));
		AGB_CUDA_CHECK(e, cudaMemcpy(opening_cursor.data(), s->opening_cursor, G * 4, cudaMemcpyDeviceToHost));
		if (s->n_openings > 0)
		{
			AGB_CUDA_CHECK(e, cudaMemcpy(openings.data(), s->openings, openings.size(), cudaMemcpyDeviceToHost));
			AGB_CUDA_CHECK(e, cudaMemcpy(opening_stm.data(), s->opening_stm, opening_stm.size(), cudaMemcpyDeviceToHost));
		}
		AGB_CUDA_CHECK(e, cudaMemcpy(boards.data(), s->root_board, boards.size(), cudaMemcpyDeviceToHost));
		AGB_CUDA_CHECK(e, cudaMemcpy(stm.data(), s->root_stm, G, cudaMemcpyDeviceToHost));
		AGB_CUDA_CHECK(e, cudaMemcpy(n_moves.data(), s->n_moves, G * 4, cudaMemcpyDeviceToHost));
		AGB_CUDA_CHECK(e, cudaMemcpy(moves.data(), s->moves, moves.size() * 2, cudaMemcpyDeviceToHost));
		AGB_CUDA_CHECK(e, cudaMemcpy(rec_len.data(), s->rec_len, G * 4, cudaMemcpyDeviceToHost));
		AGB_CUDA_CHECK(e, cudaMemcpy(rec_samples.data(), s->rec_samples, G * 4, cudaMemcpyDeviceToHost));
		AGB_CUDA_CHECK(e, cudaMemcpy(rec.data(), s->rec_buf, rec.size(), cudaMemcpyDeviceToHost));
		std::vector<uint8_t> out;
		const auto put = [&](const void *ptr, size_t bytes)
		{
			const uint8_t *b = static_cast<const uint8_t*>(ptr);
			out.insert(out.end(), b, b + bytes);
		};
		const AgbSavedHeader header { 0x53424741u /* "AGBS" */, 2u, s->games, e->cfg.rows, e->cfg.cols, e->cfg.rules };
		put(&header, sizeof(header));
		for (size_t g = 0; g < G; g++)
		{
			put(boards.data() + g * cells, cells);
			put(&stm[g], 1);
			put(&n_moves[g], 4);
			put(moves.data() + g * cells, static_cast<size_t>(n_moves[g]) * 2);
			put(&rec_samples[g], 4);
			put(&rec_len[g], 4);
			put(rec.data() + g * s->rec_cap, rec_len[g]);
		}
		put(sym_counter.data(), G * 4);
		put(noise_counter.data(), G * 4);
		put(opening_cursor.data(), G * 4);
		const int32_t n_openings = s->n_openings;
		put(&n_openings, 4);
		put(openings.data(), openings.size());
		put(opening_stm.data(), opening_stm.size());
		*used = out.size();
		if (blob_host == nullptr or capacity < out.size())
			return e->fail(AGB_ENOMEM, "state buffer too small: need " + std::to_string(out.size()) + " bytes");
		std::memcpy(blob_host, out.data(), out.size());
		return AGB_OK;
	}
	int agb_load_games(AgbEngine *e, const void *blob_host, size_t bytes)
	{
		const agb::DeviceGuard on_device(e);
		SelfplayState *s = e->selfplay;
		if (s == nullptr)
			return e->fail(AGB_ESTATE, "engine was created without games");
		const uint8_t *cur = static_cast<const uint8_t*>(blob_host), *end = cur + bytes;
		AgbSavedHeader header { };
		if (blob_host == nullptr or bytes < sizeof(header))
			return e->fail(AGB_EINVAL, "saved state is truncated");
		std::memcpy(&header, cur, sizeof(header));
		cur += sizeof(header);
		if (header.magic != 0x53424741u or (header.version != 1u and header.version != 2u))
			return e->fail(AGB_EINVAL, "not a saved-games blob of this library");
		if (header.games != s->games or header.rows != e->cfg.rows or header.cols != e->cfg.cols or header.rules != e->cfg.rules)
			return e->fail(AGB_EINVAL, "saved state was written for another configuration (games, board or rules differ)");
		const size_t G = s->games, cells = s->cells;
		std::vector<int8_t> boards(G * cells), stm(G);
		std::vector<int32_t> n_moves(G), rec_len(G), rec_samples(G);
		std::vector<uint16_t> moves(G * cells, 0);
		std::vector<uint8_t> rec(G * s->rec_cap, 0);
		const auto get = [&](void *dst, size_t n) -> bool
		{
			if (static_cast<size_t>(end - cur) < n)
				return false;
			std::memcpy(dst, cur, n);
			cur += n;
			return true;
		};
		for (size_t g = 0; g < G; g++)
		{
			bool ok = get(boards.data() + g * cells, cells) and get(&stm[g], 1) and get(&n_moves[g], 4);
			ok = ok and n_moves[g] >= 0 and n_moves[g] <= static_cast<int32_t>(cells) and get(moves.data() + g * cells, static_cast<size_t>(n_moves[g]) * 2);
			ok = ok and get(&rec_samples[g], 4) and get(&rec_len[g], 4) and rec_len[g] >= 4 and rec_len[g] <= s->rec_cap and get(rec.data() + g * s->rec_cap, rec_len[g]);
			if (not ok)
				return e->fail(AGB_EINVAL, "saved state is truncated or corrupt at game " + std::to_string(g));
		}
		std::vector<uint32_t> sym_counter(G, 0), noise_counter(G, 0), opening_cursor(G, 0);
		std::vector<int8_t> openings, opening_stm;
		int32_t n_openings = -1;
		if (header.version >= 2u)
		{
			bool ok = get(sym_counter.data(), G * 4) and get(noise_counter.data(), G * 4) and get(opening_cursor.data(), G * 4) and get(&n_openings, 4);
			ok = ok and n_openings >= 0 and n_openings <= static_cast<int32_t>(G);
			if (ok)
			{
				openings.resize(static_cast<size_t>(n_openings) * cells);
				opening_stm.resize(n_openings);
				ok = get(openings.data(), openings.size()) and get(opening_stm.data(), opening_stm.size());
			}
			if (not ok)
				return e->fail(AGB_EINVAL, "saved state is truncated or corrupt in its random-stream section");
		}
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		if (header.version >= 2u)
		{
			if (s->sym_counter != nullptr)
				AGB_CUDA_CHECK(e, cudaMemcpy(s->sym_counter, sym_counter.data(), G * 4, cudaMemcpyHostToDevice));
			if (s->noise_counter != nullptr)
				AGB_CUDA_CHECK(e, cudaMemcpy(s->noise_counter, noise_counter.data(), G * 4, cudaMemcpyHostToDevice));
			AGB_CUDA_CHECK(e, cudaMemcpy(s->opening_cursor, opening_cursor.data(), G * 4, cudaMemcpyHostToDevice));
			if (n_openings > 0)
			{
				if (s->openings == nullptr)
				{
					AGB_CUDA_CHECK(e, cudaMalloc(&s->openings, G * cells));
					AGB_CUDA_CHECK(e, cudaMalloc(&s->opening_stm, G));
				}
				AGB_CUDA_CHECK(e, cudaMemcpy(s->openings, openings.data(), openings.size(), cudaMemcpyHostToDevice));
				AGB_CUDA_CHECK(e, cudaMemcpy(s->opening_stm, opening_stm.data(), opening_stm.size(), cudaMemcpyHostToDevice));
			}
			s->n_openings = n_openings;
		}
		AGB_CUDA_CHECK(e, cudaMemcpy(s->root_board, boards.data(), boards.size(), cudaMemcpyHostToDevice));
		AGB_CUDA_CHECK(e, cudaMemcpy(s->root_stm, stm.data(), G, cudaMemcpyHostToDevice));
		AGB_CUDA_CHECK(e, cudaMemcpy(s->n_moves, n_moves.data(), G * 4, cudaMemcpyHostToDevice));
		AGB_CUDA_CHECK(e, cudaMemcpy(s->moves, moves.data(), moves.size() * 2, cudaMemcpyHostToDevice));
		AGB_CUDA_CHECK(e, cudaMemcpy(s->rec_len, rec_len.data(), G * 4, cudaMemcpyHostToDevice));
		AGB_CUDA_CHECK(e, cudaMemcpy(s->rec_samples, rec_samples.data(), G * 4, cudaMemcpyHostToDevice));
		AGB_CUDA_CHECK(e, cudaMemcpy(s->rec_buf, rec.data(), rec.size(), cudaMemcpyHostToDevice));
		// prepare_search for every game: empty trees, fresh hashes (the solver tables stay, like the reference's)
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->table, 0xFF, G * s->table_size * sizeof(int32_t), e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(e->d_status, 0, 4, e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->paused, 0, G, e->stream));
		e->overflow_seen = 0;
		const Params p = make_params(e);
		reset_games_kernel<<<static_cast<unsigned>((G + 127) / 128), 128, 0, e->stream>>>(p, 1);
		e->launches++;
		AGB_CUDA_CHECK(e, cudaGetLastError());
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		return AGB_OK;
	}

	// Player::setBoard + search until isSearchOver + getMove (src/evaluation/Player.cpp) for many games at once: every active game
	// searches the given position with a fresh tree and stops at its decision; the host plays the moves (evaluation games between two
	// engines, alphagomoku_b200/arena.py). moves[games]: Move::toShort, 0 for inactive games; root_values[games][2]: win and draw rate
	// of the root after the search.
	int agb_think(AgbEngine *e, const int8_t *boards_host, const int8_t *sign_to_move_host, const int8_t *active_host, uint16_t *moves_host,
			float *root_values_host, int max_steps)
	{
		const agb::DeviceGuard on_device(e);
		SelfplayState *s = e->selfplay;
		if (s == nullptr)
			return e->fail(AGB_ESTATE, "engine was created without games");
		if (boards_host == nullptr or sign_to_move_host == nullptr or active_host == nullptr or moves_host == nullptr)
			return e->fail(AGB_EINVAL, "null pointer");
		const int G = s->games;
		int rc = validate_boards(e, boards_host, sign_to_move_host, static_cast<size_t>(G));
		if (rc == AGB_OK and not s->reset_done)
			rc = agb_selfplay_reset(e, nullptr, nullptr); // brand-new players: empty trees
		if (rc != AGB_OK)
			return rc;
		int remaining = 0;
		for (int g = 0; g < G; g++)
			remaining += active_host[g] ? 1 : 0;
		std::vector<uint8_t> paused(G);
		{ // Player::setBoard for every active game: the trees keep what the new positions can still reach (fresh after agb_selfplay_reset)
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(s->task_boards, boards_host, static_cast<size_t>(G) * s->cells, cudaMemcpyHostToDevice, e->stream));
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(s->task_stm, sign_to_move_host, G, cudaMemcpyHostToDevice, e->stream));
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(s->slot_is_root, active_host, G, cudaMemcpyHostToDevice, e->stream));
			const Params p = make_params(e);
			rebase_games_kernel<<<static_cast<unsigned>((G + 3) / 4), 128, 0, e->stream>>>(p, s->task_boards, s->task_stm, reinterpret_cast<const int8_t*>(s->slot_is_root));
			e->launches++;
			AGB_CUDA_CHECK(e, cudaGetLastError());
		}
		e->think_mode = true;
		int steps = 0;
		while (remaining > 0)
		{
			rc = agb_step(e, 4);
			if (rc != AGB_OK)
				break;
			steps += 4;
			AGB_CUDA_CHECK(e, cudaMemcpy(paused.data(), s->paused, G, cudaMemcpyDeviceToHost));
			remaining = 0;
			for (int g = 0; g < G; g++)
				remaining += paused[g] ? 0 : 1;
			if (remaining > 0 and max_steps > 0 and steps >= max_steps)
			{
				rc = e->fail(AGB_ESTATE, "agb_think: " + std::to_string(remaining) + " games have not decided after " + std::to_string(steps) + " steps");
				break;
			}
		}
		e->think_mode = false;
		if (rc != AGB_OK)
			return rc;
		AGB_CUDA_CHECK(e, cudaMemcpy(moves_host, s->decision, G * sizeof(uint16_t), cudaMemcpyDeviceToHost));
		if (root_values_host != nullptr)
		{
			std::vector<float> roots(static_cast<size_t>(G) * 4);
			AGB_CUDA_CHECK(e, cudaMemcpy(roots.data(), s->sample_root, roots.size() * sizeof(float), cudaMemcpyDeviceToHost));
			for (int g = 0; g < G; g++)
			{
				root_values_host[2 * g + 0] = active_host[g] ? roots[4 * g + 0] : 0.0f;
				root_values_host[2 * g + 1] = active_host[g] ? roots[4 * g + 1] : 0.0f;
			}
		}
		return AGB_OK;
	}

	int agb_set_symmetry_table(AgbEngine *e, const int8_t *table_host, int n)
	{
		const agb::DeviceGuard on_device(e);
		SelfplayState *s = e->selfplay;
		if (s == nullptr or s->sym_counter == nullptr)
			return e->fail(AGB_ESTATE, "engine was created without games or without use_symmetries");
		if (n < 0 or (n > 0 and table_host == nullptr))
			return e->fail(AGB_EINVAL, "bad symmetry table");
		for (int i = 0; i < n; i++)
			if (table_host[i] < 0 or table_host[i] > 7)
				return e->fail(AGB_EINVAL, "symmetry must be in 0..7");
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		if (s->sym_table != nullptr)
			cudaFree(s->sym_table);
		s->sym_table = nullptr;
		s->sym_table_size = 0;
		if (n > 0)
		{
			AGB_CUDA_CHECK(e, cudaMalloc(&s->sym_table, n));
			AGB_CUDA_CHECK(e, cudaMemcpy(s->sym_table, table_host, n, cudaMemcpyHostToDevice));
			s->sym_table_size = n;
		}
		return AGB_OK;
	}

	int agb_set_solver_keys(AgbEngine *e, const uint64_t *keys_host, size_t n_words)
	{
		const agb::DeviceGuard on_device(e);
		SelfplayState *s = e->selfplay;
		const size_t per_set = static_cast<size_t>(e->cells) * 4;
		const size_t games = (s != nullptr) ? static_cast<size_t>(s->games) : 0;
		if (keys_host == nullptr or (n_words != per_set and (games == 0 or n_words != per_set * games)))
			return e->fail(AGB_EINVAL, "expected 2 x 64-bit words for each of 2 * rows * cols (cell, colour) pairs, once or once per game");
		// agb_solve uses the first set (kept on the host for a scratch that is created later)
		e->solver_keys_host.assign(keys_host, keys_host + per_set);
		if (s != nullptr and s->solver.keys != nullptr)
		{
			if (n_words != per_set)
			{ // one key set per game, like one AlphaBetaSearch per GameGenerator in the reference
				uint64_t *fresh = nullptr;
				AGB_CUDA_CHECK(e, cudaMalloc(&fresh, n_words * sizeof(uint64_t)));
				cudaFree(s->solver.keys);
				s->solver.keys = fresh;
				s->solver.keys_stride = per_set;
			}
			else
				s->solver.keys_stride = 0;
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(s->solver.keys, keys_host, n_words * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
		}
		solve_scratch_destroy(e); // rebuilt with the new words on the next agb_solve
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		return AGB_OK;
	}

	// diagnostics (not part of the public header): per game, SM clocks and positions visited by the last solver launch
	int agb_debug_solver_load(AgbEngine *e, unsigned long long *out_host)
	{
		const agb::DeviceGuard on_device(e);
		SelfplayState *s = e->selfplay;
		if (s == nullptr or s->solver.game_cycles == nullptr)
			return e->fail(AGB_ESTATE, "solver is off");
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		AGB_CUDA_CHECK(e, cudaMemcpy(out_host, s->solver.game_cycles, static_cast<size_t>(s->games) * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
		return AGB_OK;
	}

	int agb_step(AgbEngine *e, int n_steps)
	{
		const agb::DeviceGuard on_device(e);
		SelfplayState *s = e->selfplay;
		if (s == nullptr)
			return e->fail(AGB_ESTATE, "engine was created without games");
		if (e->net == nullptr)
			return e->fail(AGB_ESTATE, "no weights loaded");
		const int groups = s->groups;
		const int per_group = (s->games + groups - 1) / groups;
		while (static_cast<int>(e->events.size()) < 4 * n_steps * groups)
		{
			cudaEvent_t ev;
			AGB_CUDA_CHECK(e, cudaEventCreate(&ev));
			e->events.push_back(ev);
		}
		unsigned long long evals_before = 0;
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(&evals_before, s->stats + ST_EVALS, 8, cudaMemcpyDeviceToHost, e->stream));
		// fork: the group streams and the network stream start after whatever is already queued on the engine's stream
		if (groups > 1)
		{
			AGB_CUDA_CHECK(e, cudaEventRecord(s->joined, e->stream));
			for (int k = 0; k < groups; k++)
				AGB_CUDA_CHECK(e, cudaStreamWaitEvent(s->group_stream[k], s->joined, 0));
			AGB_CUDA_CHECK(e, cudaStreamWaitEvent(s->nn_stream, s->joined, 0));
			for (int k = 0; k < groups; k++)
				if (s->nn_group_stream[k] != nullptr)
					AGB_CUDA_CHECK(e, cudaStreamWaitEvent(s->nn_group_stream[k], s->joined, 0));
		}
		for (int step = 0; step < n_steps; step++)
			for (int k = 0; k < groups; k++)
			{ // one lockstep iteration of group k; with two groups the only cross-stream edges are group -> network -> group
				Params p = make_params(e);
				p.game_begin = k * per_group;
				p.game_count = std::min(per_group, s->games - p.game_begin);
				p.slot_base = p.game_begin * s->batch;
				p.eval_count = s->eval_count + k;
				int32_t *nn_count = s->nn_count + k;
				const int max_tasks = p.game_count * s->batch;
				const unsigned grid = static_cast<unsigned>((p.game_count + 3) / 4);
				cudaStream_t gs = (groups > 1) ? s->group_stream[k] : e->stream;
				cudaStream_t ns = (groups > 1) ? (s->nn_group_stream[k] != nullptr ? s->nn_group_stream[k] : s->nn_stream) : e->stream;
				cudaEvent_t *ev = e->events.data() + 4 * (step * groups + k); // before / after K5 (group stream), before / after K4 (network stream)

				AGB_CUDA_CHECK(e, cudaMemsetAsync(p.eval_count, 0, sizeof(int32_t), gs));
				select_kernel<<<grid, 128, 0, gs>>>(p);
				e->launches++;
				// K1 + K3 on the leaf positions, then K5 and K4; all read their batch size from device memory
				int rc = launch_set_boards_counted(e, s->task_boards, s->task_stm, p.eval_count, max_tasks, s->features, p.slot_base, gs);
				if (rc != AGB_OK)
					return rc;
				// K5 runs on the solver's SMs: with green contexts that is another stream than the group's (whose kernels run on the tree SMs)
				cudaStream_t ks = (s->green and p.solver_mode != 0) ? s->solver_stream[k] : gs;
				if (ks != gs)
				{
					AGB_CUDA_CHECK(e, cudaMemsetAsync(nn_count, 0, sizeof(int32_t), gs));
					AGB_CUDA_CHECK(e, cudaEventRecord(s->solver_go[k], gs));
					AGB_CUDA_CHECK(e, cudaStreamWaitEvent(ks, s->solver_go[k], 0));
				}
				AGB_CUDA_CHECK(e, cudaEventRecord(ev[0], ks));
				if (p.solver_mode != 0)
				{ // K5 on every leaf; only unproven positions (and roots) go on to the network
					if (ks == gs)
						AGB_CUDA_CHECK(e, cudaMemsetAsync(nn_count, 0, sizeof(int32_t), gs));
					SolverState solver = s->solver;
					if (p.use_symmetries)
						solver.sym = p.sym; // the evaluator's symmetry draw happens where tasks are scheduled to the network, i.e. in K5
					rc = launch_solve_games(e, solver, p.game_begin, p.game_count, s->solver_out, s->slot_is_root, s->nn_list + p.slot_base, nn_count, ks, s->solver_sms, s->green);
					if (rc != AGB_OK)
						return rc;
				}
				AGB_CUDA_CHECK(e, cudaEventRecord(ev[1], ks));
				if (ks != gs)
				{
					AGB_CUDA_CHECK(e, cudaEventRecord(s->solver_done[k], ks));
					AGB_CUDA_CHECK(e, cudaStreamWaitEvent(gs, s->solver_done[k], 0));
				}
				const bool sym = p.use_symmetries != 0;
				const size_t off = static_cast<size_t>(p.slot_base);
				// Only K4 itself goes on the network stream; what surrounds it (augment before, dense value layers and inverse symmetries after)
				// stays on the group's stream, so that with several groups one K4 launch follows the other without a gap.
				if (sym)
				{ // NNEvaluator::pack_to_network: features.augment(symmetry), over all slots of the group
					rc = launch_augment(e, s->features + off * s->cells, s->features_aug + off * s->cells, s->task_sym + off, max_tasks, gs);
					if (rc != AGB_OK)
						return rc;
				}
				if (groups > 1)
				{
					AGB_CUDA_CHECK(e, cudaEventRecord(s->ready[k], gs));
					AGB_CUDA_CHECK(e, cudaStreamWaitEvent(ns, s->ready[k], 0));
				}
				AGB_CUDA_CHECK(e, cudaEventRecord(ev[2], ns));
				rc = net_forward_dev_gather(e, sym ? s->features_aug : s->features, p.solver_mode != 0 ? nn_count : p.eval_count,
						p.solver_mode != 0 ? s->nn_list + p.slot_base : nullptr, max_tasks, sym ? s->policy_raw : s->policy, s->value, sym ? s->q_raw : s->q, p.slot_base, ns, s->net_sms,
						gs, s->evaluated[k]);
				if (rc != AGB_OK)
					return rc;
				AGB_CUDA_CHECK(e, cudaEventRecord(ev[3], ns));
				if (sym)
				{ // unpack_from_network: the inverse symmetry on the policy and the action values
					rc = launch_symmetry_f32(e, s->policy_raw + off * s->cells, s->policy + off * s->cells, s->task_sym + off, max_tasks, 1, true, gs);
					if (rc == AGB_OK and e->cfg.q_head)
						rc = launch_symmetry_f32(e, s->q_raw + off * s->cells * 3, s->q + off * s->cells * 3, s->task_sym + off, max_tasks, 3, true, gs);
					if (rc != AGB_OK)
						return rc;
				}
				expand_backup_kernel<<<grid, 128, 0, gs>>>(p);
				make_move_kernel<<<grid, 128, 0, gs>>>(p);
				e->launches += 2;
				AGB_CUDA_CHECK(e, cudaGetLastError());
			}
		// join: everything the groups queued is ordered before what follows on the engine's stream
		if (groups > 1)
			for (int k = 0; k < groups; k++)
			{
				AGB_CUDA_CHECK(e, cudaEventRecord(s->ready[k], s->group_stream[k]));
				AGB_CUDA_CHECK(e, cudaStreamWaitEvent(e->stream, s->ready[k], 0));
			}
		uint32_t status = 0;
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(&status, e->d_status, 4, cudaMemcpyDeviceToHost, e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(e->d_status, 0, 4, e->stream)); // reported once below; the engine stays usable (AgbStats keeps the flags)
		unsigned long long evals_after = 0;
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(&evals_after, s->stats + ST_EVALS, 8, cudaMemcpyDeviceToHost, e->stream));
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		double call_nn_ms = 0.0, call_solver_ms = 0.0;
		for (int i = 0; i < n_steps * groups; i++)
		{ // device time of the network kernels (K4 + value head) and of the solver kernel (K5) of this call
			float ms = 0.0f;
			if (cudaEventElapsedTime(&ms, e->events[4 * i + 2], e->events[4 * i + 3]) == cudaSuccess)
			{
				e->nn_kernel_ns += static_cast<uint64_t>(ms * 1.0e6);
				call_nn_ms += ms;
			}
			if (e->cfg.solver_max_positions > 0 and cudaEventElapsedTime(&ms, e->events[4 * i], e->events[4 * i + 1]) == cudaSuccess)
				call_solver_ms += ms;
		}
		// K5: with three or more groups the solver launches of different groups queue behind each other on the solver's SMs, so their event
		// intervals overlap; what is reported (and balanced) is the time during which at least one of them was queued or running
		float call_first = 0.0f, call_last = 0.0f, call_solver_busy = 0.0f;
		bool spans_ok = false;
		if (e->cfg.solver_max_positions > 0)
		{
			std::vector<std::pair<float, float>> solver_spans;
			spans_ok = true;
			for (int i = 0; i < n_steps * groups and spans_ok; i++)
			{
				float t[4] = { 0, 0, 0, 0 };
				for (int k = 0; k < 4 and spans_ok; k++)
					spans_ok = cudaEventElapsedTime(&t[k], e->events[0], e->events[4 * i + k]) == cudaSuccess;
				solver_spans.emplace_back(t[0], t[1]);
				call_first = (i == 0) ? t[0] : std::min(call_first, t[0]);
				call_last = std::max(call_last, t[3]);
			}
			if (spans_ok and not solver_spans.empty())
			{
				std::sort(solver_spans.begin(), solver_spans.end());
				float open_from = solver_spans[0].first, open_to = solver_spans[0].second;
				for (const auto &span : solver_spans)
				{
					if (span.first > open_to)
					{
						call_solver_busy += open_to - open_from;
						open_from = span.first;
						open_to = span.second;
					}
					else
						open_to = std::max(open_to, span.second);
				}
				call_solver_busy += open_to - open_from;
				e->solver_kernel_ns += static_cast<uint64_t>(call_solver_busy * 1.0e6);
			}
			else
				e->solver_kernel_ns += static_cast<uint64_t>(call_solver_ms * 1.0e6);
		}
		if (s->solver_sms > 0 and not s->green and e->cfg.solver_sms == 0 and n_steps >= 2 and call_nn_ms > 0.0 and call_solver_ms > 0.0 and groups >= 3)
		{ // automatic partition with three or more groups: the solver launches of different groups queue behind each other on the solver's SMs (a
		  // launch's tail runs next to the next group's games) and the network launches follow each other on the network stream, so either side
		  // can be kept busy all the time. Measure how long each side had NOTHING queued or running during this call and move SMs from the idler
		  // side to the busier one, half of what the difference suggests, in whole TPCs. Results do not depend on it.
			const int sms = s->solver_sms + s->net_sms;
			const float wall = call_last - call_first, solver_busy = call_solver_busy, net_busy = static_cast<float>(call_nn_ms);
			if (spans_ok and wall > 0.0f)
			{
				const double solver_idle = std::max(0.0, 1.0 - solver_busy / wall), net_idle = std::max(0.0, 1.0 - net_busy / wall);
				const double delta = 0.5 * (net_idle * s->net_sms - solver_idle * s->solver_sms);
				int next = s->solver_sms + 2 * static_cast<int>(std::lround(delta / 2.0)); // whole TPCs, rounded to the nearest (not down)
				// more SMs than hold the games of all the other groups at once (28 warps per SM) cannot be filled
				const int useful = std::max(8, static_cast<int>(0.77 * (groups - 1) * per_group / 28.0) & ~1);
				next = std::max(8, std::min(next, std::min(sms / 2, useful)));
				s->solver_sms = next;
				s->net_sms = sms - next;
			}
		}
		else if (s->solver_sms > 0 and not s->green and e->cfg.solver_sms == 0 and n_steps >= 2 and call_nn_ms > 0.0 and call_solver_ms > 0.0)
		{ // automatic partition, two groups: both kernels scale with their SMs, so split the SMs in proportion to the SM-time each needed in this
		  // call (K5 and K4 launches then take equally long); move three quarters of the way, in whole TPCs. Results do not depend on it.
			const int sms = s->solver_sms + s->net_sms;
			// Lean 15 % towards the network: measured on the steady-state workload the step is shortest when the solver's launches take about that
			// much longer than the network's (bias 0.85: 237 k evaluations/s on 46 solver SMs, 1.0: 232 k on 50, 1.15: 227 k on 56) -- a network
			// launch that waits a little for its solver loses less than network CTAs that lack SMs all the time. Early game: 345 / 339 / 335 k.
			static const double bias = getenv("AGB_BALANCE_BIAS") != nullptr ? atof(getenv("AGB_BALANCE_BIAS")) : 0.85;
			const double solver_work = bias * call_solver_ms * s->solver_sms, net_work = call_nn_ms * s->net_sms;
			const double target = sms * solver_work / (solver_work + net_work);
			int next = static_cast<int>(s->solver_sms + 0.75 * (target - s->solver_sms) + 0.5) & ~1;
			// more SMs than about three quarters of what holds every game of a group at once (28 warps per SM) do not shorten a solver launch any
			// further -- it then ends with its slowest games, not with the queue -- and only starve the network (steady state, 2048 games per group:
			// 56 SMs 138 ms of K5 per step, 74 SMs 138 ms; K4 102 against 126 ms)
			const int useful = std::max(8, static_cast<int>(0.77 * per_group / 28.0) & ~1);
			next = std::max(8, std::min(next, std::min(sms / 2, useful)));
			s->solver_sms = next;
			s->net_sms = sms - next;
		}
		if (const char *trace = getenv("AGB_STEP_TRACE"))
		{ // diagnostics: when each group's solver and network phases ran, in ms from the first event of the call
			if (FILE *f = fopen(trace, "a"))
			{
				for (int i = 0; i < n_steps * groups; i++)
				{
					float t[4] = { 0, 0, 0, 0 };
					for (int k = 0; k < 4; k++)
						cudaEventElapsedTime(&t[k], e->events[0], e->events[4 * i + k]);
					fprintf(f, "step %d group %d: solver %.2f-%.2f network %.2f-%.2f\n", i / groups, i % groups, t[0], t[1], t[2], t[3]);
				}
				fclose(f);
			}
		}
		e->nn_kernel_launches += static_cast<uint64_t>(n_steps) * groups;
		e->nn_positions += evals_after - evals_before;
		e->overflow_seen |= status;
		if (status == OVF_FINISHED)
			return e->fail(AGB_EOVERFLOW, "finished-game queue is full: call agb_pop_finished; the finished games wait with their records (nothing was lost) "
					"and are published by the next agb_step");
		if (status != 0)
			return e->fail(AGB_EOVERFLOW, "device-side overflow, flags=" + std::to_string(status)
					+ " (1 renju recursion, 2 nodes, 4 edges, 8 path, 16 table, 32 record, 64 finished queue, 256/512 solver forbidden-move recursion/cache, 1024 solver action stack, 2048 solver frames)");
		return AGB_OK;
	}

	int agb_pop_finished(AgbEngine *e, void *records_host, size_t capacity, size_t *used, int *n_games)
	{
		const agb::DeviceGuard on_device(e);
		SelfplayState *s = e->selfplay;
		if (s == nullptr)
			return e->fail(AGB_ESTATE, "engine was created without games");
		if (used == nullptr or n_games == nullptr)
			return e->fail(AGB_EINVAL, "null pointer");
		*used = 0;
		*n_games = 0;
		unsigned long long bytes = 0;
		int32_t games = 0;
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(&bytes, s->fin_used, 8, cudaMemcpyDeviceToHost, e->stream));
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(&games, s->fin_games, 4, cudaMemcpyDeviceToHost, e->stream));
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		if (bytes == 0)
			return AGB_OK;
		if (bytes > capacity or records_host == nullptr)
			return e->fail(AGB_ENOMEM, "record buffer too small: need " + std::to_string(bytes) + " bytes");
		AGB_CUDA_CHECK(e, cudaMemcpyAsync(records_host, s->fin_buf, bytes, cudaMemcpyDeviceToHost, e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->fin_used, 0, 8, e->stream));
		AGB_CUDA_CHECK(e, cudaMemsetAsync(s->fin_games, 0, 4, e->stream));
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		*used = bytes;
		*n_games = games;
		return AGB_OK;
	}
	int agb_get_stats(AgbEngine *e, AgbStats *stats)
	{
		const agb::DeviceGuard on_device(e);
		if (stats == nullptr)
			return AGB_EINVAL;
		*stats = AgbStats { };
		stats->nb_kernel_launches = e->launches;
		stats->nn_kernel_ns = e->nn_kernel_ns;
		stats->nn_kernel_launches = e->nn_kernel_launches;
		stats->nn_positions = e->nn_positions;
		stats->solver_kernel_ns = e->solver_kernel_ns;
		stats->solver_sms = (e->selfplay != nullptr) ? static_cast<uint64_t>(e->selfplay->solver_sms) : 0;
		stats->pipeline_groups = (e->selfplay != nullptr) ? static_cast<uint64_t>(e->selfplay->groups) : 0;
		if (e->selfplay != nullptr)
		{
			unsigned long long h[16];
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(h, e->selfplay->stats, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
			uint32_t status = 0;
			AGB_CUDA_CHECK(e, cudaMemcpyAsync(&status, e->d_status, 4, cudaMemcpyDeviceToHost, e->stream));
			AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
			status |= e->overflow_seen; // flags already reported (and cleared on the device) stay visible here
			stats->nb_network_evaluations = h[ST_EVALS];
			stats->nb_node_count = h[ST_NODES];
			stats->nb_duplicate_nodes = h[ST_DUP];
			stats->nb_information_leaks = h[ST_LEAKS];
			stats->nb_proven_states = h[ST_PROVEN];
			stats->nb_wasted_expansions = h[ST_WASTED];
			stats->nb_moves_played = h[ST_MOVES];
			stats->nb_games_finished = h[ST_GAMES];
			stats->overflow_flags = status;
		}
		return AGB_OK;
	}
	int agb_get_root(AgbEngine *e, int game, int32_t *visits_host, float *priors_host, float *q_host, float *root_value3_host, int32_t *root_visits)
	{
		const agb::DeviceGuard on_device(e);
		SelfplayState *s = e->selfplay;
		if (s == nullptr or game < 0 or game >= s->games)
			return e->fail(AGB_EINVAL, "bad game index");
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		int32_t root = -1;
		AGB_CUDA_CHECK(e, cudaMemcpy(&root, s->root_node + game, 4, cudaMemcpyDeviceToHost));
		const int cells = s->cells;
		if (visits_host)
			std::memset(visits_host, 0, cells * 4);
		if (priors_host)
			std::memset(priors_host, 0, cells * 4);
		if (q_host)
			std::memset(q_host, 0, cells * 4);
		if (root_visits)
			*root_visits = 0;
		if (root < 0)
		{ // no root yet: a default-constructed Node (Node.hpp), i.e. Value() whose loss rate is 1
			if (root_value3_host)
			{
				root_value3_host[0] = 0.0f;
				root_value3_host[1] = 0.0f;
				root_value3_host[2] = 1.0f;
			}
			return AGB_OK;
		}
		NodeD node;
		AGB_CUDA_CHECK(e, cudaMemcpy(&node, s->nodes + static_cast<size_t>(game) * s->max_nodes + root, sizeof(NodeD), cudaMemcpyDeviceToHost));
		std::vector<EdgeD> edges(node.n_edges);
		AGB_CUDA_CHECK(e, cudaMemcpy(edges.data(), s->edges + static_cast<size_t>(game) * s->max_edges + node.edge_begin, sizeof(EdgeD) * node.n_edges, cudaMemcpyDeviceToHost));
		for (const EdgeD &ed : edges)
		{
			const int cell = ((ed.move >> 2) & 127) * s->S + ((ed.move >> 9) & 127);
			if (visits_host)
				visits_host[cell] = ed.visits;
			if (priors_host)
				priors_host[cell] = ed.prior;
			if (q_host)
				q_host[cell] = ed.win + 0.5f * ed.draw;
		}
		if (root_value3_host)
		{
			root_value3_host[0] = node.win;
			root_value3_host[1] = node.draw;
			root_value3_host[2] = 1.0f - (node.win + node.draw);
		}
		if (root_visits)
			*root_visits = node.visits;
		return AGB_OK;
	}
	int agb_get_root_scores(AgbEngine *e, int game, uint16_t *edge_scores_host, uint16_t *root_score)
	{
		const agb::DeviceGuard on_device(e);
		SelfplayState *s = e->selfplay;
		if (s == nullptr or game < 0 or game >= s->games or edge_scores_host == nullptr or root_score == nullptr)
			return e->fail(AGB_EINVAL, "bad game index or null pointer");
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		for (int i = 0; i < s->cells; i++)
			edge_scores_host[i] = score::kDefault;
		*root_score = score::kDefault;
		int32_t root = -1;
		AGB_CUDA_CHECK(e, cudaMemcpy(&root, s->root_node + game, 4, cudaMemcpyDeviceToHost));
		if (root < 0)
			return AGB_OK;
		NodeD node;
		AGB_CUDA_CHECK(e, cudaMemcpy(&node, s->nodes + static_cast<size_t>(game) * s->max_nodes + root, sizeof(NodeD), cudaMemcpyDeviceToHost));
		std::vector<EdgeD> edges(node.n_edges);
		AGB_CUDA_CHECK(e, cudaMemcpy(edges.data(), s->edges + static_cast<size_t>(game) * s->max_edges + node.edge_begin, sizeof(EdgeD) * node.n_edges, cudaMemcpyDeviceToHost));
		for (const EdgeD &ed : edges)
			edge_scores_host[((ed.move >> 2) & 127) * s->S + ((ed.move >> 9) & 127)] = ed.score;
		*root_score = node.score;
		return AGB_OK;
	}
	int agb_get_root_noise(AgbEngine *e, int game, float *noisy_policy_host)
	{
		const agb::DeviceGuard on_device(e);
		SelfplayState *s = e->selfplay;
		if (s == nullptr or game < 0 or game >= s->games or noisy_policy_host == nullptr)
			return e->fail(AGB_EINVAL, "bad game index or null pointer");
		if (s->root_noise == nullptr)
			return e->fail(AGB_ESTATE, "engine was created without root noise");
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		const int cells = s->cells;
		std::memset(noisy_policy_host, 0, cells * sizeof(float));
		int32_t root = -1;
		uint8_t ready = 0;
		AGB_CUDA_CHECK(e, cudaMemcpy(&root, s->root_node + game, 4, cudaMemcpyDeviceToHost));
		AGB_CUDA_CHECK(e, cudaMemcpy(&ready, s->noise_ready + game, 1, cudaMemcpyDeviceToHost));
		if (root < 0 or not ready)
			return AGB_OK;
		NodeD node;
		AGB_CUDA_CHECK(e, cudaMemcpy(&node, s->nodes + static_cast<size_t>(game) * s->max_nodes + root, sizeof(NodeD), cudaMemcpyDeviceToHost));
		std::vector<EdgeD> edges(node.n_edges);
		std::vector<float> noisy(node.n_edges);
		AGB_CUDA_CHECK(e, cudaMemcpy(edges.data(), s->edges + static_cast<size_t>(game) * s->max_edges + node.edge_begin, sizeof(EdgeD) * node.n_edges, cudaMemcpyDeviceToHost));
		AGB_CUDA_CHECK(e, cudaMemcpy(noisy.data(), s->root_noise + static_cast<size_t>(game) * cells, sizeof(float) * node.n_edges, cudaMemcpyDeviceToHost));
		for (int i = 0; i < node.n_edges; i++)
			noisy_policy_host[((edges[i].move >> 2) & 127) * s->S + ((edges[i].move >> 9) & 127)] = noisy[i];
		return AGB_OK;
	}
	int agb_get_board(AgbEngine *e, int game, int8_t *board_host, int8_t *sign_to_move, int32_t *move_number)
	{
		const agb::DeviceGuard on_device(e);
		SelfplayState *s = e->selfplay;
		if (s == nullptr or game < 0 or game >= s->games)
			return e->fail(AGB_EINVAL, "bad game index");
		AGB_CUDA_CHECK(e, cudaStreamSynchronize(e->stream));
		if (board_host)
			AGB_CUDA_CHECK(e, cudaMemcpy(board_host, s->root_board + static_cast<size_t>(game) * s->cells, s->cells, cudaMemcpyDeviceToHost));
		if (sign_to_move)
			AGB_CUDA_CHECK(e, cudaMemcpy(sign_to_move, s->root_stm + game, 1, cudaMemcpyDeviceToHost));
		if (move_number)
		{
			std::vector<int8_t> b(s->cells);
			AGB_CUDA_CHECK(e, cudaMemcpy(b.data(), s->root_board + static_cast<size_t>(game) * s->cells, s->cells, cudaMemcpyDeviceToHost));
			int n = 0;
			for (int8_t v : b)
				n += (v != 0);
			*move_number = n;
		}
		return AGB_OK;
	}
}
