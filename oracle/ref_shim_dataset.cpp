// TEST INFRASTRUCTURE -- C-ABI shim over the reference's trainer-side readers: GameDataBuffer::load -> GameDataStorage::getSample
// (src/dataset/GameDataStorage.cpp:109-148, SearchDataStorage_v201::storeTo, src/dataset/SearchDataStorage.cpp:237-274) and
// SamplerVisits::prepare_training_data (src/dataset/Sampler.cpp:96-131), so that the records the device engine writes can be read back
// through the code the reference's trainer uses.
#include <memory>
#include <string>
#include <vector>
// prepare_training_data is a private virtual of the samplers
#define private public
#include <alphagomoku/dataset/Sampler.hpp>
#undef private
#include <alphagomoku/dataset/GameDataBuffer.hpp>
#include <alphagomoku/dataset/GameDataStorage.hpp>
#include <alphagomoku/dataset/data_packs.hpp>
#include <alphagomoku/game/Game.hpp>
#include <alphagomoku/utils/file_util.hpp>

#include <minml/utils/json.hpp>
#include <minml/utils/serialization.hpp>

#include <cstdio>
#include <cstring>

using namespace ag;

extern "C"
{
	// FileLoader (src/utils/file_util.cpp:63-118) on a file: the JSON part re-dumped without indentation and the binary part, as the reference
	// splits them. Returns the binary size, or -1.
	long agref_file_load(const char *path, int uncompress, char *json_out, size_t json_capacity, uint8_t *binary_out, size_t binary_capacity)
	{
		try
		{
			FileLoader fl(path, uncompress != 0);
			const std::string text = fl.getJson().dump(-1);
			if (text.size() + 1 > json_capacity or fl.getBinaryData().size() > binary_capacity)
				return -1;
			std::memcpy(json_out, text.c_str(), text.size() + 1);
			std::memcpy(binary_out, fl.getBinaryData().data(), fl.getBinaryData().size());
			return static_cast<long>(fl.getBinaryData().size());
		}
		catch (const std::exception &ex)
		{
			std::fprintf(stderr, "agref_file_load: %s\n", ex.what());
			return -1;
		}
	}
	// A saved_state/thread_<i>.bin file through the reference's own code: loaded like GameGenerator::load (Game(json, binary),
	// GameDataStorage(binary, offset, 201); src/selfplay/GameGenerator.cpp:131-141, GeneratorManager.cpp:112-119) and written back like
	// GameGenerator::save + GeneratorThread::saveGames (:122-129, :98-111). info[3 * i ...] = moves played, samples stored, sign to move of game i.
	int agref_saved_state_roundtrip(const char *in_path, const char *out_path, int32_t *info, int capacity)
	{
		try
		{
			FileLoader fl(in_path, true);
			Json out_json(JsonType::Array);
			SerializedObject out_so;
			const int n = fl.getJson().size();
			for (int i = 0; i < n and i < capacity; i++)
			{
				const Json &entry = fl.getJson()[i];
				Game game(entry, fl.getBinaryData());
				size_t offset = entry["offset"].getLong();
				GameDataStorage storage(fl.getBinaryData(), offset, 201);
				info[3 * i + 0] = game.numberOfMoves();
				info[3 * i + 1] = storage.numberOfSamples();
				info[3 * i + 2] = static_cast<int>(game.getSignToMove());
				Json result = game.serialize(out_so);
				result["state"] = entry["state"].getInt();
				result["offset"] = out_so.size();
				storage.serialize(out_so);
				out_json[i] = result;
			}
			FileSaver fs(out_path);
			fs.save(out_json, out_so, 2, true);
			return n;
		}
		catch (const std::exception &ex)
		{
			std::fprintf(stderr, "agref_saved_state_roundtrip: %s\n", ex.what());
			return -1;
		}
	}
	// FileSaver::save(json, binary, indent, compress) (file_util.cpp:42-52) of a JSON text and a blob
	int agref_file_save(const char *path, const char *json_text, const uint8_t *binary, size_t binary_size, int indent, int compress)
	{
		try
		{
			SerializedObject so;
			so.save(binary, binary_size);
			FileSaver fs(path);
			fs.save(Json::load(json_text), so, indent, compress != 0);
			return 0;
		}
		catch (const std::exception &ex)
		{
			std::fprintf(stderr, "agref_file_save: %s\n", ex.what());
			return -1;
		}
	}

	// One sample of one game of a GameDataBuffer file, as the trainer sees it. Per cell: board int8, visits int32, prior f32, action values
	// (win, draw) f32, action scores u16. scalars[8]: minimax win, minimax draw, minimax score, moves left, outcome, played move (toShort), flags,
	// 0. With targets != nullptr also SamplerVisits' training targets per cell: policy target f32, action value targets (win, draw) f32,
	// visit count int32 (as float); target_scalars[6]: value target win / draw, minimax target win / draw, moves left, sign to move.
	int agref_buffer_sample_with(const char *path, int game, int sample, int8_t *board, int32_t *visits, float *prior, float *values, uint16_t *scores,
			float *scalars, float *policy_target, float *value_targets, float *target_visits, float *target_scalars, int sampler_kind);
	int agref_buffer_sample(const char *path, int game, int sample, int8_t *board, int32_t *visits, float *prior, float *values, uint16_t *scores,
			float *scalars, float *policy_target, float *value_targets, float *target_visits, float *target_scalars)
	{
		return agref_buffer_sample_with(path, game, sample, board, visits, prior, values, scores, scalars, policy_target, value_targets, target_visits,
				target_scalars, 0);
	}
	// sampler_kind 0: SamplerVisits, 1: SamplerValues (createSampler, Sampler.cpp:218-225)
	int agref_buffer_sample_with(const char *path, int game, int sample, int8_t *board, int32_t *visits, float *prior, float *values, uint16_t *scores,
			float *scalars, float *policy_target, float *value_targets, float *target_visits, float *target_scalars, int sampler_kind)
	{
		static std::string loaded_path;
		static std::unique_ptr<GameDataBuffer> buffer;
		try
		{
			if (buffer == nullptr or loaded_path != path)
			{
				buffer = std::make_unique<GameDataBuffer>();
				buffer->load(path);
				loaded_path = path;
			}
			const GameConfig cfg = buffer->getConfig();
			SearchDataPack pack(cfg.rows, cfg.cols);
			buffer->getGameData(game).getSample(pack, sample);
			const int cells = cfg.rows * cfg.cols;
			for (int i = 0; i < cells; i++)
			{
				board[i] = static_cast<int8_t>(pack.board[i]);
				visits[i] = pack.visit_count[i];
				prior[i] = pack.policy_prior[i];
				values[2 * i] = pack.action_values[i].win_rate;
				values[2 * i + 1] = pack.action_values[i].draw_rate;
				scores[i] = Score::to_short(pack.action_scores[i]);
			}
			scalars[0] = pack.minimax_value.win_rate;
			scalars[1] = pack.minimax_value.draw_rate;
			scalars[2] = Score::to_short(pack.minimax_score);
			scalars[3] = pack.moves_left;
			scalars[4] = static_cast<int>(pack.game_outcome);
			scalars[5] = pack.played_move.toShort();
			scalars[6] = pack.flags.raw();
			scalars[7] = 0.0f;
			if (policy_target != nullptr)
			{
				SamplerVisits by_visits;
				SamplerValues by_values;
				TrainingDataPack target(cfg.rows, cfg.cols);
				if (sampler_kind == 0)
					by_visits.prepare_training_data(target, pack);
				else
					by_values.prepare_training_data(target, pack);
				for (int i = 0; i < cells; i++)
				{
					policy_target[i] = target.policy_target[i];
					value_targets[2 * i] = target.action_values_target[i].win_rate;
					value_targets[2 * i + 1] = target.action_values_target[i].draw_rate;
					target_visits[i] = target.visit_count[i];
				}
				target_scalars[0] = target.value_target.win_rate;
				target_scalars[1] = target.value_target.draw_rate;
				target_scalars[2] = target.minimax_target.win_rate;
				target_scalars[3] = target.minimax_target.draw_rate;
				target_scalars[4] = target.moves_left;
				target_scalars[5] = static_cast<int>(target.sign_to_move);
			}
			return 0;
		}
		catch (const std::exception &ex)
		{
			std::fprintf(stderr, "agref_buffer_sample: %s\n", ex.what());
			return -1;
		}
	}
}
