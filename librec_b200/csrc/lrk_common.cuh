// Shared declarations of the B200-native LibRec MF path (handle layout, error plumbing).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include "../../include/librec_b200.h"

#define LRK_MAX_FACTORS 256
#define LRK_MAX_TOPN 512

struct lrk_handle_s {
    lrk_config_t cfg{};
    int k = 0;        // rec.factor.number
    int ld = 0;       // padded fp32 row length: power of two >= max(k,4); multiple of 128 above 128
    int G = 0;        // lanes cooperating on one rating (each lane owns V float4 of a row)
    int V = 1;
    int32_t U = 0, I = 0;
    int64_t nnz = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;          // second stream: the H2D copy of the rating values overlaps the sort of the staging
    cudaEvent_t ev_copy0 = nullptr, ev_copy1 = nullptr;

    // train CSR (device): membership for BPR sampling and the top-N train mask
    int64_t* d_rowptr = nullptr;
    int32_t* d_col = nullptr;
    // shuffled COO stream for the SGD epoch (SoA, 12 B / rating)
    int32_t* d_su = nullptr;
    int32_t* d_si = nullptr;
    float* d_sr = nullptr;
    bool has_train = false;
    void* group = nullptr;    // GroupUnits*: the stream is unit-ordered (staging_group.cuh) and the epoch runs sgd_group_epoch_kernel
    int64_t run_tiles = 0;    // 32-rating item-run tiles in the staged stream (lrk_stage_stats)
    uint32_t max_item_deg = 0;
    double hot_share = 0.0;   // largest share one item has of the train ratings (stability cap of the SGD grid)
    float* d_pnorm2 = nullptr;        // mean |p_u|^2 at the start of the epoch (curvature term of that step)
    float pnorm2_host = 0.f;          // its host copy, refreshed with every loss read-back (picks the kernel variant)
    float pnorm2_prev = 0.f;          // the value one epoch earlier (growth of the user factors, sgd_launch_gv)
    uint32_t* d_item_deg = nullptr;   // ratings per item in this handle's shard (staleness-aware step of run tiles, sgd.cuh)
    uint32_t* d_item_cum = nullptr;   // RankSGD: inclusive prefix sums of d_item_deg (negative sampling table)

    // factors: fp32 working copies (padded rows) + fp64 masters (dense rows, what Java sees)
    float *P32 = nullptr, *Q32 = nullptr, *bu32 = nullptr, *bi32 = nullptr;
    double *P64 = nullptr, *Q64 = nullptr, *bu64 = nullptr, *bi64 = nullptr;
    double mu = 0.0;
    bool has_factors = false;
    bool f64_valid = false;   // masters are in sync with the fp32 working copies

    // Safeguard of the fast SGD mode: factors are snapshotted before an epoch; a non-finite (or exploding) loss
    // rolls the epoch back and re-runs it with 4x fewer ratings in flight (sticky, relaxed again after 8 good epochs)
    float *bk_P = nullptr, *bk_Q = nullptr, *bk_bu = nullptr, *bk_bi = nullptr;
    int conc_div = 1;           // divisor of the SGD grid
    int good_epochs = 0;
    double prev_loss = -1.0;    // loss of the last accepted epoch (< 0: none yet)
    int64_t rollbacks = 0;
    int epochs_done = 0;        // accepted epochs since lrk_set_factors (picks the kernel variant of the first epochs)
    double* d_loss = nullptr;   // device accumulator
    double* h_loss = nullptr;   // pinned
    float* h_pnorm2 = nullptr;  // pinned
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_epoch_ms = 0.f;
    uint64_t launches = 0;

    // top-N statistics
    int64_t topn_fast_users = 0, topn_fallback_users = 0, topn_resweep_users = 0;
    float topn_ms = 0.f;
    float topn_phase_ms[4] = {0.f, 0.f, 0.f, 0.f};
    float topn_err_ratio = 0.f;   // largest observed |fp16 sweep score - exact score| / certificate bound (must be < 1)
    // result buffers of lrk_topn, reused across calls (cudaMalloc/cudaFree of 130 MB cost 20-400 ms per call)
    int32_t *tn_users = nullptr, *tn_items = nullptr, *tn_counts = nullptr;
    double* tn_scores = nullptr;
    // tensor-core top-N state (bf16 copies, norms); see topn_tc.cuh
    void* tc = nullptr;

    void* gbpr = nullptr;     // GbprState* (sgd_gbpr.cuh)
    void* svdpp = nullptr;    // SvdppState* (sgd_svdpp.cuh)
    void* aobpr = nullptr;    // AobprState* (sgd_aobpr.cuh)
    void* als = nullptr;      // AlsState* (als.cuh): WRMF / eALS

    // reference-order (wavefront) schedule, see sgd_exact.cuh
    void* exact = nullptr;

    // single-process multi-GPU parent (multi.cuh): no device state of its own, forwards to one child per device
    void* multi = nullptr;
    bool same_process = false; // DSGD child of a multi handle: its ring neighbours live in this process (CUDA IPC cannot map them;
                               // the in-kernel ring takes their buffers by plain peer pointers instead)
    lrk_handle_s** siblings = nullptr;   // same_process: the handles of all ranks, indexed by rank
    bool score_only = false;   // scoring child of a multi handle: the train CSR is kept for the top-N mask only (no COO stream)

    // DSGD
    void* comm = nullptr;   // ncclComm_t
    int rank = 0, world = 1;
    void* dsgd = nullptr;   // DsgdState*

    // allocation bookkeeping: device buffers are reused across calls when they are large enough
    std::unordered_map<void*, size_t> caps;
    void* scratch = nullptr;      // staging workspace (temporaries of lrk_set_train_csr), grown on demand
    size_t scratch_bytes = 0;

    std::string err;
};

extern thread_local std::string g_lrk_tls_error;

static inline int lrk_fail(lrk_handle_s* h, int code, const char* what, const char* detail, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s:%d)", what, detail ? detail : "", file, line);
    if (h) h->err = buf;
    g_lrk_tls_error = buf;
    return code;
}

#define LRK_CUDA(h, call)                                                                        \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return lrk_fail((h), e__ == cudaErrorMemoryAllocation ? LRK_ERR_NOMEM : LRK_ERR_CUDA, \
                            #call, cudaGetErrorString(e__), __FILE__, __LINE__);                 \
    } while (0)

#define LRK_REQUIRE(h, cond, msg)                                                         \
    do {                                                                                  \
        if (!(cond)) return lrk_fail((h), LRK_ERR_INVALID, "invalid argument", msg, __FILE__, __LINE__); \
    } while (0)

#define LRK_LAUNCH_CHECK(h)                  \
    do {                                     \
        (h)->launches++;                     \
        LRK_CUDA((h), cudaGetLastError());   \
    } while (0)

template <typename T>
static inline int lrk_dev_alloc(lrk_handle_s* h, T** p, size_t count) {
    if (count == 0) count = 1;
    const size_t bytes = count * sizeof(T);
    if (*p) {
        auto it = h->caps.find((void*)*p);
        if (it != h->caps.end() && it->second >= bytes) return LRK_OK;   // reuse
        if (it != h->caps.end()) h->caps.erase(it);
        cudaFree(*p);
        *p = nullptr;
    }
    LRK_CUDA(h, cudaMalloc((void**)p, bytes));
    h->caps[(void*)*p] = bytes;
    return LRK_OK;
}
template <typename T>
static inline void lrk_dev_free(T** p) {
    if (*p) { cudaFree(*p); *p = nullptr; }
}
// bump allocator over the handle's staging workspace
struct LrkScratch {
    char* base = nullptr;
    size_t used = 0, cap = 0;
    template <typename T>
    T* take(size_t count) {
        const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        if (used + bytes > cap) return nullptr;
        T* r = reinterpret_cast<T*>(base + used);
        used += bytes;
        return r;
    }
};
static inline int lrk_scratch_begin(lrk_handle_s* h, size_t bytes, LrkScratch* out) {
    if (h->scratch_bytes < bytes) {
        if (h->scratch) { cudaFree(h->scratch); h->scratch = nullptr; h->scratch_bytes = 0; }
        LRK_CUDA(h, cudaMalloc(&h->scratch, bytes));
        h->scratch_bytes = bytes;
    }
    out->base = (char*)h->scratch; out->used = 0; out->cap = h->scratch_bytes;
    return LRK_OK;
}

static inline int lrk_ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
// models whose prediction adds biases (GBPR: item biases only, its user biases stay zero): b_i + p_u.q_i (+ b_u + mu)
static inline bool lrk_has_bias(const lrk_handle_s* h) { return h->cfg.model == LRK_MODEL_BIASEDMF || h->cfg.model == LRK_MODEL_GBPR || h->cfg.model == LRK_MODEL_SVDPP; }
// WRMF / eALS: alternating least squares on the fp64 masters (als.cuh); lrk_sgd_epoch = one ALS iteration
static inline bool lrk_is_als(const lrk_handle_s* h) { return h->cfg.model == LRK_MODEL_WRMF || h->cfg.model == LRK_MODEL_EALS; }
// BiasedMF / PMF: one update per train rating with the rating as target (item-run tiles, staleness-aware step)
static inline bool lrk_is_rating_model(const lrk_handle_s* h) { return h->cfg.model == LRK_MODEL_BIASEDMF || h->cfg.model == LRK_MODEL_PMF; }
