// Tensor-core (tcgen05 / TMA) top-N candidate path -- see DESIGN.md "Top-N".  Filled in below.
#pragma once
#include "lrk_common.cuh"
static inline void topn_tc_release(lrk_handle_s*) {}
static inline void topn_tc_invalidate(lrk_handle_s*) {}
static inline bool topn_tc_profitable(lrk_handle_s*, int32_t, int) { return false; }
static inline int topn_tc_run(lrk_handle_s* h, const int32_t*, int32_t, int, int, int32_t*, double*, int32_t*) {
    return lrk_fail(h, LRK_ERR_INVALID, "lrk_topn", "tensor-core candidate path not built into this library", __FILE__, __LINE__);
}
