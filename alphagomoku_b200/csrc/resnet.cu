#include "engine.hpp"
namespace agb
{
	int net_create(AgbEngine *e) { return e->fail(AGB_ESTATE, "network not built yet"); }
	void net_destroy(AgbEngine *) {}
}
extern "C"
{
	int agb_load_weights(AgbEngine *e, const void *, size_t) { return e->fail(AGB_ESTATE, "network not built yet"); }
	size_t agb_weights_size(const AgbEngine *) { return 0; }
	int agb_forward(AgbEngine *e, const uint32_t *, int, float *, float *, float *) { return e->fail(AGB_ESTATE, "network not built yet"); }
	int agb_forward_dev(AgbEngine *e, const uint32_t *, int, float *, float *, float *) { return e->fail(AGB_ESTATE, "network not built yet"); }
}
