// Placeholder entry points for the tcgen05 similarity-tile kernels (InfoNCE, all-pairs top-k) until
// simtile.cu lands; they fail loudly.
#include "common.cuh"
#define CFM_NOT_BUILT(name) do { cfm::set_error(name ": tcgen05 kernel not built yet"); return CFM_ERR_UNSUPPORTED; } while (0)
extern "C" int cfm_allpairs_topk(const float*, const float*, const void*, const void*, int64_t, int64_t, int64_t, int64_t, int64_t, double, int64_t, float*, int64_t*, int32_t*, float*, int32_t*, void*) { CFM_NOT_BUILT("cfm_allpairs_topk"); }
extern "C" int cfm_topk_merge(const float*, const int64_t*, int64_t, int64_t, int64_t, float*, int64_t*, void*) { CFM_NOT_BUILT("cfm_topk_merge"); }
extern "C" int cfm_allpairs_rank(const float*, const float*, int64_t, int64_t, int64_t, const int64_t*, int64_t*, void*) { CFM_NOT_BUILT("cfm_allpairs_rank"); }
