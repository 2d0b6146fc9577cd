#include "engine.hpp"
namespace agb
{
	int selfplay_create(AgbEngine *e) { return e->fail(AGB_ESTATE, "self-play not built yet"); }
	void selfplay_destroy(AgbEngine *) {}
}
extern "C"
{
	int agb_selfplay_reset(AgbEngine *e, const int8_t *, const int8_t *) { return e->fail(AGB_ESTATE, "self-play not built yet"); }
	int agb_step(AgbEngine *e, int) { return e->fail(AGB_ESTATE, "self-play not built yet"); }
	int agb_pop_finished(AgbEngine *e, void *, size_t, size_t *, int *) { return e->fail(AGB_ESTATE, "self-play not built yet"); }
	int agb_get_stats(AgbEngine *e, AgbStats *stats)
	{
		if (stats == nullptr) return AGB_EINVAL;
		*stats = AgbStats { };
		stats->nb_kernel_launches = e->launches;
		return AGB_OK;
	}
	int agb_get_root(AgbEngine *e, int, int32_t *, float *, float *, float *, int32_t *) { return e->fail(AGB_ESTATE, "self-play not built yet"); }
	int agb_get_board(AgbEngine *e, int, int8_t *, int8_t *, int32_t *) { return e->fail(AGB_ESTATE, "self-play not built yet"); }
}
