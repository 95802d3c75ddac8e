// Exact fp64 prediction and top-N kernels (sm_100a).  They reproduce the reference arithmetic
// bit for bit when given identical factors:
//   predict      recommender/MatrixFactorizationRecommender.java:104-106, cf/rating/BiasedMFRecommender.java:118-120
//   dot order    math/structure/DenseVector.java:104-111 (f = 0..k-1, from 0.0, separate multiply and add:
//                __dmul_rn/__dadd_rn keep nvcc from contracting them into an FMA, which Java never does)
//   top-N        recommender/MatrixRecommender.java:153-201 + util/Lists.java:416-468:
//                java.util.PriorityQueue min-heap of size N fed in ascending item order, an item
//                replaces the minimum only if strictly greater under Double.compareTo, result is
//                the heap ARRAY order followed by a stable descending sort.
// The heap walk is inherently sequential per user, so one warp owns a user: lanes score 32
// consecutive items in parallel (each lane a full left-to-right dot), then the warp replays the
// 32 offers in item order against the heap kept in shared memory.
// This kernel is the exact path for every catalogue and the fallback of the tensor-core
// candidate path (topn_tc.cuh) when its exactness certificate fails.
#pragma once
#include "lrk_common.cuh"

// Double.compareTo: numeric order, except -0.0 < +0.0 (NaN never reaches the heap)
__device__ __forceinline__ int jcompare(double a, double b) {
    if (a < b) return -1;
    if (a > b) return 1;
    const long long x = __double_as_longlong(a), y = __double_as_longlong(b);
    return x == y ? 0 : (x < y ? -1 : 1);
}

__device__ __forceinline__ double dot_lr_f64(const double* __restrict__ a, const double* __restrict__ b, int k) {
    double r = 0.0;
    for (int f = 0; f < k; ++f) r = __dadd_rn(r, __dmul_rn(b[f], a[f]));
    return r;
}

// predict(u,i) for explicit pairs
__global__ void predict_pairs_kernel(const double* __restrict__ P, const double* __restrict__ Q,
                                     const double* __restrict__ bu, const double* __restrict__ bi, double mu, int biased,
                                     int k, const int32_t* __restrict__ us, const int32_t* __restrict__ is, int64_t n,
                                     double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int32_t u = us[t], i = is[t];
    double d = dot_lr_f64(P + (int64_t)u * k, Q + (int64_t)i * k, k);
    if (biased) d = __dadd_rn(__dadd_rn(__dadd_rn(d, bu[u]), bi[i]), mu);
    out[t] = d;
}

// recommendRating(DataSet): bounded prediction per test entry (MatrixRecommender.java:230-248,272-284)
// + per-block partial sums of squared / absolute error (RMSEEvaluator.java:55, MAEEvaluator.java:56).
__global__ void eval_rating_kernel(const double* __restrict__ P, const double* __restrict__ Q,
                                   const double* __restrict__ bu, const double* __restrict__ bi, double mu, int biased,
                                   int k, int32_t U, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                   const double* __restrict__ val, double min_rate, double max_rate,
                                   double* __restrict__ pred_out, double* __restrict__ part_se, double* __restrict__ part_ae) {
    __shared__ double s_se[256], s_ae[256];
    const int64_t nnz = rowptr[U];
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double se = 0.0, ae = 0.0;
    if (e < nnz) {
        int32_t lo = 0, hi = U;   // largest u with rowptr[u] <= e
        while (hi - lo > 1) { const int32_t m = (lo + hi) >> 1; if (rowptr[m] <= e) lo = m; else hi = m; }
        const int32_t u = lo, i = col[e];
        double p = dot_lr_f64(P + (int64_t)u * k, Q + (int64_t)i * k, k);
        if (biased) p = __dadd_rn(__dadd_rn(__dadd_rn(p, bu[u]), bi[i]), mu);
        if (p > max_rate) p = max_rate; else if (p < min_rate) p = min_rate;
        if (p != p) p = mu;
        if (pred_out) pred_out[e] = p;
        const double d = val[e] - p;
        se = d * d; ae = fabs(d);
    }
    s_se[threadIdx.x] = se; s_ae[threadIdx.x] = ae;
    __syncthreads();
    for (int m = 128; m >= 1; m >>= 1) {
        if ((int)threadIdx.x < m) { s_se[threadIdx.x] += s_se[threadIdx.x + m]; s_ae[threadIdx.x] += s_ae[threadIdx.x + m]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part_se[blockIdx.x] = s_se[0]; part_ae[blockIdx.x] = s_ae[0]; }
}
__global__ void eval_rating_final_kernel(const double* __restrict__ part_se, const double* __restrict__ part_ae, int nb,
                                         int64_t n, double* __restrict__ out2) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double se = 0.0, ae = 0.0;
        for (int b = 0; b < nb; ++b) { se += part_se[b]; ae += part_ae[b]; }
        out2[0] = n > 0 ? sqrt(se / (double)n) : 0.0;
        out2[1] = n > 0 ? ae / (double)n : 0.0;
    }
}

// ---- java.util.PriorityQueue (JDK 8) on a shared-memory array, executed by one lane ---------
struct HeapRef {
    double* v;
    int32_t* key;
};
__device__ __forceinline__ void heap_sift_up(HeapRef h, int kpos, double xv, int32_t xk) {
    while (kpos > 0) {
        const int parent = (kpos - 1) >> 1;
        if (jcompare(xv, h.v[parent]) >= 0) break;
        h.v[kpos] = h.v[parent]; h.key[kpos] = h.key[parent];
        kpos = parent;
    }
    h.v[kpos] = xv; h.key[kpos] = xk;
}
__device__ __forceinline__ void heap_sift_down(HeapRef h, int size, int kpos, double xv, int32_t xk) {
    const int half = size >> 1;
    while (kpos < half) {
        int child = (kpos << 1) + 1;
        const int right = child + 1;
        if (right < size && jcompare(h.v[child], h.v[right]) > 0) child = right;
        if (jcompare(xv, h.v[child]) <= 0) break;
        h.v[kpos] = h.v[child]; h.key[kpos] = h.key[child];
        kpos = child;
    }
    h.v[kpos] = xv; h.key[kpos] = xk;
}
// Lists.java:436-441: poll() then add(entry) on a full heap of `size` elements
__device__ __forceinline__ void heap_replace_min(HeapRef h, int size, double xv, int32_t xk) {
    const int s = size - 1;
    const double lv = h.v[s]; const int32_t lk = h.key[s];
    if (s != 0) heap_sift_down(h, s, 0, lv, lk);
    if (s == 0) { h.v[0] = xv; h.key[0] = xk; } else heap_sift_up(h, s, xv, xk);
}

#define TOPN_WARPS 8
#define TOPN_UPW 4                      // users per warp
#define TOPN_UPB (TOPN_WARPS * TOPN_UPW)  // users per block

// smem layout (doubles first): Qs[32][k+1] | Ps[UPB][k] | hv[UPB][topn] | hk[UPB][topn] (int32)
static inline size_t topn_exact_smem(int k, int topn) {
    return sizeof(double) * ((size_t)32 * (k + 1) + (size_t)TOPN_UPB * k + (size_t)TOPN_UPB * topn) +
           sizeof(int32_t) * (size_t)TOPN_UPB * topn;
}

__global__ void __launch_bounds__(TOPN_WARPS * 32) topn_exact_kernel(
    const double* __restrict__ P, const double* __restrict__ Q, const double* __restrict__ bu,
    const double* __restrict__ bi, double mu, int biased, int k, int32_t I,
    const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int exclude_train,
    const int32_t* __restrict__ users, int32_t nq, int topn,
    int32_t* __restrict__ out_items, double* __restrict__ out_scores, int32_t* __restrict__ out_counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Qs = reinterpret_cast<double*>(smem_raw);
    double* Ps = Qs + (size_t)32 * (k + 1);
    double* hv_all = Ps + (size_t)TOPN_UPB * k;
    int32_t* hk_all = reinterpret_cast<int32_t*>(hv_all + (size_t)TOPN_UPB * topn);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qs_ld = k + 1;
    const int32_t c0 = blockIdx.x * TOPN_UPB;   // first query slot of this block

    // stage the block's user rows
    for (int idx = threadIdx.x; idx < TOPN_UPB * k; idx += blockDim.x) {
        const int s = idx / k, f = idx - s * k;
        const int32_t c = c0 + s;
        double v = 0.0;
        if (c < nq) { const int32_t u = users ? users[c] : c; v = P[(int64_t)u * k + f]; }
        Ps[idx] = v;
    }

    int32_t uu[TOPN_UPW];
    int64_t tp[TOPN_UPW], tend[TOPN_UPW];
    int hsize[TOPN_UPW];
    double ubias[TOPN_UPW];
#pragma unroll
    for (int j = 0; j < TOPN_UPW; ++j) {
        const int32_t c = c0 + warp * TOPN_UPW + j;
        uu[j] = c < nq ? (users ? users[c] : c) : -1;
        tp[j] = 0; tend[j] = 0; hsize[j] = 0; ubias[j] = 0.0;
        if (uu[j] >= 0) {
            if (exclude_train && rowptr) { tp[j] = rowptr[uu[j]]; tend[j] = rowptr[uu[j] + 1]; }
            if (biased) ubias[j] = bu[uu[j]];
        }
    }

    for (int32_t base = 0; base < I; base += 32) {
        __syncthreads();   // previous tile fully consumed (also covers the Ps staging on the first trip)
        for (int idx = threadIdx.x; idx < 32 * k; idx += blockDim.x) {
            const int r = idx / k, f = idx - r * k;
            const int32_t it = base + r;
            Qs[r * qs_ld + f] = it < I ? Q[(int64_t)it * k + f] : 0.0;
        }
        // train columns that can fall inside this tile (issued early, consumed after the dot loop)
        int32_t tc[TOPN_UPW];
#pragma unroll
        for (int j = 0; j < TOPN_UPW; ++j) tc[j] = (tp[j] + lane < tend[j]) ? __ldg(col + tp[j] + lane) : 0x7fffffff;
        __syncthreads();

        const int32_t item = base + lane;
        double acc[TOPN_UPW];
#pragma unroll
        for (int j = 0; j < TOPN_UPW; ++j) acc[j] = 0.0;
        const double* qrow = Qs + lane * qs_ld;
        const double* prow = Ps + (size_t)(warp * TOPN_UPW) * k;
        for (int f = 0; f < k; ++f) {
            const double q = qrow[f];
#pragma unroll
            for (int j = 0; j < TOPN_UPW; ++j) acc[j] = __dadd_rn(acc[j], __dmul_rn(q, prow[j * k + f]));
        }
        const double ibias = (biased && item < I) ? bi[item] : 0.0;

#pragma unroll
        for (int j = 0; j < TOPN_UPW; ++j) {
            if (uu[j] < 0) continue;   // warp-uniform
            double score = acc[j];
            if (biased) score = __dadd_rn(__dadd_rn(__dadd_rn(score, ubias[j]), ibias), mu);
            // train mask for [base, base+32): MatrixRecommender.java:170-174
            const bool in_win = tc[j] < base + 32;
            const uint32_t mybit = in_win ? (1u << (tc[j] - base)) : 0u;
            const uint32_t tmask = __reduce_or_sync(0xffffffffu, mybit);
            tp[j] += __popc(__ballot_sync(0xffffffffu, in_win));
            const bool valid = item < I && !((tmask >> lane) & 1u) && !(score != score);
            HeapRef h{hv_all + (size_t)(warp * TOPN_UPW + j) * topn, hk_all + (size_t)(warp * TOPN_UPW + j) * topn};
            uint32_t pend = __ballot_sync(0xffffffffu, valid);
            // Lists.java:431-433 : the first k entries are added unconditionally
            while (pend != 0u && hsize[j] < topn) {
                const int b = __ffs(pend) - 1;
                pend &= pend - 1;
                const double xv = __shfl_sync(0xffffffffu, score, b);
                if (lane == 0) heap_sift_up(h, hsize[j], xv, base + b);
                hsize[j]++;
                __syncwarp();
            }
            if (pend != 0u) {
                // Lists.java:434-441 : strictly greater than the current minimum replaces it
                const double hmin = h.v[0];
                pend &= __ballot_sync(0xffffffffu, valid && jcompare(score, hmin) > 0);
                while (pend != 0u) {
                    const int b = __ffs(pend) - 1;
                    pend &= pend - 1;
                    const double xv = __shfl_sync(0xffffffffu, score, b);
                    if (lane == 0 && jcompare(xv, h.v[0]) > 0) heap_replace_min(h, topn, xv, base + b);
                    __syncwarp();
                }
            }
        }
    }

    // new ArrayList<>(heap) then Collections.sort descending, stable (Lists.java:442-443,458-467)
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < TOPN_UPW; ++j) {
            if (uu[j] < 0) continue;
            const int32_t c = c0 + warp * TOPN_UPW + j;
            double* hv = hv_all + (size_t)(warp * TOPN_UPW + j) * topn;
            int32_t* hk = hk_all + (size_t)(warp * TOPN_UPW + j) * topn;
            const int n = hsize[j];
            for (int a = 1; a < n; ++a) {
                const double xv = hv[a]; const int32_t xk = hk[a];
                int b = a - 1;
                while (b >= 0 && jcompare(hv[b], xv) < 0) { hv[b + 1] = hv[b]; hk[b + 1] = hk[b]; --b; }
                hv[b + 1] = xv; hk[b + 1] = xk;
            }
            for (int t = 0; t < topn; ++t) {
                out_items[(int64_t)c * topn + t] = t < n ? hk[t] : -1;
                out_scores[(int64_t)c * topn + t] = t < n ? hv[t] : 0.0;
            }
            out_counts[c] = n;
        }
    }
}

// exact fp64 top-N for `nq` query slots whose user ids are in d_users (device) or NULL (= 0..nq-1);
// results to device buffers
static int topn_exact_launch(lrk_handle_s* h, const int32_t* d_users, int32_t nq, int topn, int exclude_train,
                             int32_t* d_items, double* d_scores, int32_t* d_counts) {
    const size_t smem = topn_exact_smem(h->k, topn);
    LRK_CUDA(h, cudaFuncSetAttribute(topn_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = lrk_ceil_div(nq, TOPN_UPB);
    topn_exact_kernel<<<grid, TOPN_WARPS * 32, smem, h->stream>>>(
        h->P64, h->Q64, h->bu64, h->bi64, h->mu, h->cfg.model == LRK_MODEL_BIASEDMF, h->k, h->I,
        h->d_rowptr, h->d_col, exclude_train, d_users, nq, topn, d_items, d_scores, d_counts);
    LRK_LAUNCH_CHECK(h);
    return LRK_OK;
}

