// HP-2 placeholder during bring-up (replaced by the real engine).
#include "../../include/dfsfm_b200.h"
#include "engine_common.h"
struct dfsfm_refine { int dummy; };
extern "C" {
int dfsfm_refine_create(dfsfm_refine_t** out, int, int, int) { return dfsfm::guard([&] { (void)out; throw dfsfm::Error("refine engine not built yet"); }); }
void dfsfm_refine_destroy(dfsfm_refine_t*) {}
int dfsfm_refine_set_param(dfsfm_refine_t*, const char*, const float*, int64_t, int64_t, int) { return dfsfm::guard([&] { throw dfsfm::Error("refine engine not built yet"); }); }
int dfsfm_refine_chunk(dfsfm_refine_t*, int, const float* const*, const int32_t*, const int32_t*, const float*, int, int, const float*, const float*,
                       const uint8_t*, const int32_t*, const int32_t*, const uint8_t*, float*, float*, float*, void*) {
    return dfsfm::guard([&] { throw dfsfm::Error("refine engine not built yet"); });
}
}
