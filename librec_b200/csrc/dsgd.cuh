// DSGD strata over NCCL -- see DESIGN.md "Multi-GPU".  Filled in below.
#pragma once
#include "lrk_common.cuh"
static inline void dsgd_release(lrk_handle_s*) {}
static inline int dsgd_unique_id(uint8_t*) { return lrk_fail(nullptr, LRK_ERR_NCCL, "lrk_comm_unique_id", "DSGD not built", __FILE__, __LINE__); }
static inline int dsgd_comm_init(lrk_handle_s* h, int, int, const uint8_t*) { return lrk_fail(h, LRK_ERR_NCCL, "lrk_comm_init", "DSGD not built", __FILE__, __LINE__); }
static inline int dsgd_set_train_csr(lrk_handle_s* h, int32_t, int32_t, const int64_t*, const int32_t*, const double*) { return lrk_fail(h, LRK_ERR_INVALID, "dsgd", "not built", __FILE__, __LINE__); }
static inline int dsgd_set_factors(lrk_handle_s* h, const double*, const double*, const double*, const double*, double) { return lrk_fail(h, LRK_ERR_INVALID, "dsgd", "not built", __FILE__, __LINE__); }
static inline int dsgd_get_factors(lrk_handle_s* h, double*, double*, double*, double*) { return lrk_fail(h, LRK_ERR_INVALID, "dsgd", "not built", __FILE__, __LINE__); }
static inline int dsgd_epoch(lrk_handle_s* h, float, float, float, double, int32_t, double*) { return lrk_fail(h, LRK_ERR_INVALID, "dsgd", "not built", __FILE__, __LINE__); }
