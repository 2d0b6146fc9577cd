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

#include <cstdio>
#include <cstring>

using namespace ag;

extern "C"
{
	// One sample of one game of a GameDataBuffer file, as the trainer sees it. Per cell: board int8, visits int32, prior f32, action values
	// (win, draw) f32, action scores u16. scalars[8]: minimax win, minimax draw, minimax score, moves left, outcome, played move (toShort), flags,
	// 0. With targets != nullptr also SamplerVisits' training targets per cell: policy target f32, action value targets (win, draw) f32,
	// visit count int32 (as float); target_scalars[6]: value target win / draw, minimax target win / draw, moves left, sign to move.
	int agref_buffer_sample(const char *path, int game, int sample, int8_t *board, int32_t *visits, float *prior, float *values, uint16_t *scores,
			float *scalars, float *policy_target, float *value_targets, float *target_visits, float *target_scalars)
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
				SamplerVisits sampler;
				TrainingDataPack target(cfg.rows, cfg.cols);
				sampler.prepare_training_data(target, pack);
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
