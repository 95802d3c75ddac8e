// BiasedMF / PMF epoch over a unit-ordered stream (staging_group.cuh) -- the default fast kernel of the rating models.
//
// Reference loop restated per rating (core/src/main/java/net/librec/): recommender/cf/rating/BiasedMFRecommender.java:72-98,
// recommender/cf/rating/PMFSimilarityRecommender.java:64-82.  The reference walks users in CSR order, items ascending, one rating
// at a time.  Here a WORKER (G lanes, one float4 of a row per lane) owns a unit: up to 16 consecutive users whose factor rows (and
// biases) it keeps in shared memory for the whole unit -- read once, written once, exclusively, with plain loads / stores -- and
// walks the unit's ratings item by item (rotated ascending order).  For every (unit, item) pair it reads the item row once, applies
// the pair's ratings SEQUENTIALLY (every rating sees the previous one's update of the item row and of its own user row: exact
// Gauss-Seidel inside the unit, as in the reference) and adds the accumulated change of the item row to global memory with one
// vector reduction (red.global.add.v4.f32).  What is concurrent is only the item side ACROSS units.
//
// Why: the item-run-tile kernel (sgd.cuh) pays one row gather + one row RED per rating for the user side; lrk_probe_l2 shows the L2
// serves such REDs at ~6.3 TB/s against ~19.6 TB/s for gathers, and that kernel sits at ~0.87 of the 1:1 mix rate.  Units cut the
// gathers + REDs per rating to the number of distinct (unit, item) pairs per rating: 0.60-0.69 on the ML-20M shape, 0.53-0.64 on the
// Netflix shape, with the user side down to two row transfers per user and unit.
//
// Concurrency of the item side: W workers hold an item row for the length of a pair (+ one rating of prefetch), so about
// deg_i * (W * hold) / n ratings of item i are in flight at once.  As in sgd.cuh the RED of a pair is scaled by (1 - exp(-x)) / x,
// x = lr * max(1, mean |p_u|^2 of the unit) * (deg_i - ratings of the pair) * inflight_frac: plain SGD for x -> 0, the sequential
// limit for the items everybody rates.  The curvature term comes from the unit's own rows (computed while they are loaded).
//
// Heavy users (more ratings in the block than a unit holds) are single-user units per slice; their slices run concurrently, so such a
// unit adds (row - row as loaded) with a RED instead of storing the row.
#pragma once
#include "lrk_common.cuh"
#include "sgd.cuh"
#include "staging_group.cuh"

struct SgdGroupParams {
    const int32_t* __restrict__ su;
    const int32_t* __restrict__ si;
    const float* __restrict__ sr;
    const int4* __restrict__ units;     // first unit of this launch
    int32_t n_units;
    unsigned int* counter;              // dynamic unit fetch (zeroed before the launch)
    float* P; float* Q; float* bu; float* bi;
    float mu, lr, reg_u, reg_i, reg_b;
    double* loss;
    int ld;
    const uint32_t* __restrict__ item_deg;   // ratings per item in this launch's shard (block-local ids); NULL: no damping
    float inflight_frac;                // (resident workers * ratings a worker holds an item row for) / ratings of the launch
};

#define LRK_GROUP_HOLD 2.5f            // mean ratings a worker holds an item row for: pair length (~1.6) + one rating of prefetch

template <int G, int V>
__host__ __device__ constexpr int sgd_group_smem_floats_per_worker() { return LRK_GS * (4 * G * V) + LRK_GS; }

// One worker = G lanes.  All G-lane workers of a warp run in lock-step: loops are warp-uniform, everything that depends on a
// worker's own unit is predicated (the predicated blocks contain no warp-synchronous operation).  A lane only ever touches ITS four
// columns of the unit's rows in shared memory, and only sub-lane 0 touches the biases, so no __syncwarp is needed inside a unit.
template <int G, int V, bool BIASED>
__device__ __forceinline__ void sgd_group_segment(const SgdGroupParams& p, float* smem_cta, double& loss_d) {
    constexpr int NWW = 32 / G;                    // workers per warp
    constexpr int LDS = 4 * G * V;                 // floats per row
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int sub = lane % G;
    const int grp = lane / G;
    const int wk = (threadIdx.x >> 5) * NWW + grp;
    float* Ps = smem_cta + (size_t)wk * sgd_group_smem_floats_per_worker<G, V>();
    float* bus = Ps + LRK_GS * LDS;
    const float lr = p.lr, reg_u = p.reg_u, reg_i = p.reg_i, reg_b = p.reg_b, mu = p.mu;
    bool alive = true;

    for (;;) {
        int unit = -1;
        if (alive && sub == 0) unit = (int)atomicAdd(p.counter, 1u);
        unit = __shfl_sync(FULL, unit, 0, G);
        if (unit < 0 || unit >= p.n_units) alive = false;
        if (!__any_sync(FULL, alive)) break;
        int4 d = make_int4(0, 0, 0, 0);
        if (alive) d = __ldg(p.units + unit);
        const int64_t start = (int64_t)(uint32_t)d.x;
        const int count = alive ? d.y : 0;
        const int first = d.z;
        const int nus = d.w & 0xffff;
        const bool shared = (d.w >> 16) != 0;

        // ---- the unit's user rows -> shared memory (each lane its own four columns of every row)
        float4 p0[V];
        float pn2 = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) p0[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < nus; ++j) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float4 x = ldcg4(p.P + (int64_t)(first + j) * p.ld + (v * G + sub) * 4);
                *reinterpret_cast<float4*>(Ps + j * LDS + (v * G + sub) * 4) = x;
                pn2 += dot4(x, x);
                if (j == 0) p0[v] = x;
            }
            if (BIASED && sub == 0) bus[j] = __ldcg(p.bu + first + j);
        }
        pn2 = group_sum<G>(pn2);
        const float curv = fmaxf(1.f, nus > 0 ? pn2 / (float)nus : 0.f);
        const float bu0 = (BIASED && sub == 0 && nus > 0) ? bus[0] : 0.f;

        // ---- walk the unit's ratings: chunks of G entries (lane `sub` loads entry c + sub), one rating per step and worker
        int32_t cur = -1;                 // item whose row is in q (block-local id)
        float4 q[V], dq[V], qn[V];
        int32_t qn_item = -1;
        float bic = 0.f, dbi = 0.f, bin = 0.f;
        int pair_len = 0;
        float loss_f = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) { q[v] = make_float4(0.f, 0.f, 0.f, 0.f); dq[v] = q[v]; qn[v] = q[v]; }

        auto flush = [&]() {
            if (cur >= 0) {
                float damp = 1.f;
                if (p.item_deg) {
                    // ratings of this item in flight in OTHER workers (the pair itself was applied sequentially: nothing stale in it)
                    const float others = fmaxf((float)__ldg(p.item_deg + cur) - (float)pair_len, 0.f);
                    const float x = lr * curv * others * p.inflight_frac;
                    if (x > 1e-3f) damp = (1.f - __expf(-x)) / x;
                }
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float4 t = make_float4(dq[v].x * damp, dq[v].y * damp, dq[v].z * damp, dq[v].w * damp);
                    apply4<true>(p.Q + (int64_t)cur * p.ld + (v * G + sub) * 4, t, t);
                }
                if (BIASED && sub == 0) apply1<true>(p.bi + cur, 0.f, dbi * damp);
            }
        };

        int warp_count = count;                      // the chunk loop is warp-uniform: longest unit among the warp's workers
#pragma unroll
        for (int m = G; m < 32; m <<= 1) warp_count = max(warp_count, __shfl_xor_sync(FULL, warp_count, m));
        for (int c = 0; c < warp_count; c += G) {
            int32_t u_l = -1, i_l = -1;
            float r_l = 0.f;
            if (c + sub < count) {
                const int64_t e = start + c + sub;
                u_l = __ldcs(p.su + e) - first; i_l = __ldcs(p.si + e); r_l = __ldcs(p.sr + e);
            }
            int32_t un = __shfl_sync(FULL, u_l, 0, G), in_ = __shfl_sync(FULL, i_l, 0, G);
            float rn = __shfl_sync(FULL, r_l, 0, G);
            // first entry of the chunk: its item row unless it continues the current pair or was prefetched
            if (un >= 0 && in_ != cur && in_ != qn_item) {
#pragma unroll
                for (int v = 0; v < V; ++v) qn[v] = ldcg4(p.Q + (int64_t)in_ * p.ld + (v * G + sub) * 4);
                if (BIASED && sub == 0) bin = __ldcg(p.bi + in_);
                qn_item = in_;
            }
#pragma unroll
            for (int s = 0; s < G; ++s) {
                const int32_t uc = un, ic = in_;
                const float rc = rn;
                if (s + 1 < G) {
                    un = __shfl_sync(FULL, u_l, s + 1, G); in_ = __shfl_sync(FULL, i_l, s + 1, G); rn = __shfl_sync(FULL, r_l, s + 1, G);
                }
                const bool act = uc >= 0;
                if (act && ic != cur) {          // new pair: flush the old one, take the prefetched row
                    flush();
#pragma unroll
                    for (int v = 0; v < V; ++v) { q[v] = qn[v]; dq[v] = make_float4(0.f, 0.f, 0.f, 0.f); }
                    bic = bin; dbi = 0.f; cur = ic; pair_len = 0; qn_item = -1;
                }
                // prefetch the next pair's item row while this rating is processed
                if (s + 1 < G && un >= 0 && in_ != ic) {
                    {
#pragma unroll
                        for (int v = 0; v < V; ++v) qn[v] = ldcg4(p.Q + (int64_t)in_ * p.ld + (v * G + sub) * 4);
                        if (BIASED && sub == 0) bin = __ldcg(p.bi + in_);
                        qn_item = in_;
                    }
                }
                float4 pc[V];
                float part = 0.f, buc = 0.f;
                const int urow = act ? uc : 0;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    pc[v] = *reinterpret_cast<const float4*>(Ps + urow * LDS + (v * G + sub) * 4);
                    part += dot4(pc[v], q[v]);
                }
                if (BIASED && sub == 0) { buc = bus[urow]; part += buc + bic + mu; }
                const float pred = group_sum<G>(part);
                const float err = rc - pred;
                if (act) {
                    float reg_acc = 0.f;
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const float4 a = pc[v], b = q[v];
                        float4 np, dqs;
                        np.x = a.x + lr * (err * b.x - reg_u * a.x); dqs.x = lr * (err * a.x - reg_i * b.x);
                        np.y = a.y + lr * (err * b.y - reg_u * a.y); dqs.y = lr * (err * a.y - reg_i * b.y);
                        np.z = a.z + lr * (err * b.z - reg_u * a.z); dqs.z = lr * (err * a.z - reg_i * b.z);
                        np.w = a.w + lr * (err * b.w - reg_u * a.w); dqs.w = lr * (err * a.w - reg_i * b.w);
                        *reinterpret_cast<float4*>(Ps + urow * LDS + (v * G + sub) * 4) = np;
                        q[v].x += dqs.x; q[v].y += dqs.y; q[v].z += dqs.z; q[v].w += dqs.w;
                        dq[v].x += dqs.x; dq[v].y += dqs.y; dq[v].z += dqs.z; dq[v].w += dqs.w;
                        reg_acc += reg_u * dot4(a, a) + reg_i * dot4(b, b);
                    }
                    if (sub == 0) {
                        reg_acc += err * err;
                        if (BIASED) {
                            bus[urow] = buc + lr * (err - reg_b * buc);
                            const float dbs = lr * (err - reg_b * bic);
                            reg_acc += reg_b * (buc * buc + bic * bic);
                            bic += dbs; dbi += dbs;
                        }
                    }
                    loss_f += reg_acc;
                    ++pair_len;
                }
            }
        }
        flush();
        cur = -1;
        loss_d += (double)loss_f;

        // ---- write the unit's user rows back: exclusive owner -> plain stores; slice of a heavy user -> RED of the change
        if (shared) {
            if (nus > 0) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float4 x = *reinterpret_cast<const float4*>(Ps + (v * G + sub) * 4);
                    const float4 t = make_float4(x.x - p0[v].x, x.y - p0[v].y, x.z - p0[v].z, x.w - p0[v].w);
                    apply4<true>(p.P + (int64_t)first * p.ld + (v * G + sub) * 4, t, t);
                }
                if (BIASED && sub == 0) apply1<true>(p.bu + first, 0.f, bus[0] - bu0);
            }
        } else {
            for (int j = 0; j < nus; ++j) {
#pragma unroll
                for (int v = 0; v < V; ++v)
                    __stcg(reinterpret_cast<float4*>(p.P + (int64_t)(first + j) * p.ld + (v * G + sub) * 4),
                           *reinterpret_cast<const float4*>(Ps + j * LDS + (v * G + sub) * 4));
                if (BIASED && sub == 0) __stcg(p.bu + first + j, bus[j]);
            }
        }
    }
}

template <int G, int V, bool BIASED>
__global__ void __launch_bounds__(256, 3) sgd_group_epoch_kernel(SgdGroupParams p) {
    extern __shared__ float4 lrk_group_smem4[];
    double loss_d = 0.0;
    sgd_group_segment<G, V, BIASED>(p, reinterpret_cast<float*>(lrk_group_smem4), loss_d);
    block_loss_commit(loss_d, p.loss);
}

template <int G, int V>
static size_t sgd_group_smem_bytes() { return sizeof(float) * (size_t)(8 * (32 / G)) * sgd_group_smem_floats_per_worker<G, V>(); }

// resident workers of the group kernel for a layout (sets the unit size at staging and the in-flight estimate at launch)
template <int G, int V>
static int sgd_group_resident_workers_gv(lrk_handle_s* h, int* ctas_out) {
    const size_t smem = sgd_group_smem_bytes<G, V>();
    cudaFuncSetAttribute(sgd_group_epoch_kernel<G, V, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(sgd_group_epoch_kernel<G, V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sgd_group_epoch_kernel<G, V, true>, 256, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    if (ctas_out) *ctas_out = per_sm * h->sm_count;
    return per_sm * h->sm_count * 8 * (32 / G);
}
static bool sgd_group_layout_supported(const lrk_handle_s* h) { return h->V == 1 && (h->G == 8 || h->G == 16 || h->G == 32); }
// LRK_SGD_GROUP=0 keeps the item-run-tile stream kernel of sgd.cuh (A/B probe)
static bool lrk_use_group_kernel(const lrk_handle_s* h) {
    static const char* env = getenv("LRK_SGD_GROUP");
    if (env && atoi(env) == 0) return false;
    return lrk_is_rating_model(h) && h->cfg.update_mode == LRK_UPDATE_ATOMIC && sgd_group_layout_supported(h);
}
static int sgd_group_resident_workers(lrk_handle_s* h, int* ctas_out) {
    switch (h->G) {
        case 8: return sgd_group_resident_workers_gv<8, 1>(h, ctas_out);
        case 16: return sgd_group_resident_workers_gv<16, 1>(h, ctas_out);
        default: return sgd_group_resident_workers_gv<32, 1>(h, ctas_out);
    }
}

template <int G, int V>
static int sgd_group_launch_gv(lrk_handle_s* h, SgdGroupParams& gp, int64_t n_ratings, int conc_div) {
    int ctas = 0;
    const int workers_full = sgd_group_resident_workers_gv<G, V>(h, &ctas);
    (void)workers_full;
    constexpr int WPC = 8 * (32 / G);
    int64_t grid = ctas;
    const int64_t need = (gp.n_units + WPC - 1) / WPC;
    if (need < grid) grid = need;
    if (conc_div > 1) grid /= conc_div;                      // rollback safeguard: fewer units in flight
    if (grid < 1) grid = 1;
    const double workers = (double)std::min<int64_t>(grid * WPC, gp.n_units);
    gp.inflight_frac = (float)(workers * (double)LRK_GROUP_HOLD / (double)(n_ratings > 0 ? n_ratings : 1));
    const size_t smem = sgd_group_smem_bytes<G, V>();
    if (h->cfg.model == LRK_MODEL_BIASEDMF) sgd_group_epoch_kernel<G, V, true><<<(unsigned)grid, 256, smem, h->stream>>>(gp);
    else sgd_group_epoch_kernel<G, V, false><<<(unsigned)grid, 256, smem, h->stream>>>(gp);
    LRK_LAUNCH_CHECK(h);
    return LRK_OK;
}
static int sgd_group_launch(lrk_handle_s* h, SgdGroupParams& gp, int64_t n_ratings, int conc_div) {
    switch (h->G) {
        case 8: return sgd_group_launch_gv<8, 1>(h, gp, n_ratings, conc_div);
        case 16: return sgd_group_launch_gv<16, 1>(h, gp, n_ratings, conc_div);
        case 32: return sgd_group_launch_gv<32, 1>(h, gp, n_ratings, conc_div);
        default: return lrk_fail(h, LRK_ERR_INVALID, "sgd_group_launch", "unsupported factor layout", __FILE__, __LINE__);
    }
}
