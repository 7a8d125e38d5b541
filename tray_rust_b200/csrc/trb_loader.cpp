// trb_loader.cpp — JSON / OBJ / MERL loading (Scene::load_file, src/scene.rs:101-146). Placeholder until the
// loader lands: the entry points exist so the ABI is complete.
#include "../../include/trb.h"
extern "C" {
trb_status trb_desc_load_json(const char*, uint32_t, uint32_t, uint32_t, trb_scene_desc** out) { if (out) *out = nullptr; return TRB_UNSUPPORTED; }
void trb_desc_free(trb_scene_desc*) {}
trb_status trb_scene_load_json(const char*, uint32_t, uint32_t, uint32_t, int, trb_scene** out) { if (out) *out = nullptr; return TRB_UNSUPPORTED; }
}
