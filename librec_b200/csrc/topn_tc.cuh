// Tensor-core top-N (sm_100a): recommendRank()'s user x item score sweep
// (recommender/MatrixRecommender.java:153-201) as a tcgen05 / TMA bf16 GEMM with the selection fused
// into the epilogue, followed by an exact fp64 re-score -- so the lists that leave the library are
// still bit-identical to the reference (util/Lists.java:416-468 semantics).
//
//   1. operands: A = queried user factors, B = item factors, both K-major bf16, K padded to 64;
//      BiasedMF folds the item bias into two extra K columns (hi/lo bf16 split against 1.0 in A).
//   2. topn_tc_kernel (persistent, warp-specialised, 1 CTA/SM, 320 threads):
//        warp 0  TMA producer   cp.async.bulk.tensor.2d, SWIZZLE_128B, mbarrier ring
//        warp 1  MMA issuer     tcgen05.mma.cta_group::1.kind::f16, M=128 N=128 K=16; two user
//                               sub-tiles (256 users) share every item tile -> halves L2->SM traffic;
//                               accumulators double-buffered in all 512 TMEM columns
//        warps 2-9 epilogue     tcgen05.ld 32x32b.x32: one thread owns one user row; a score is
//                               appended to the row's candidate list only if it beats the row's running
//                               threshold tau (1 FMNMX per score + 1 compare per 8 scores in the common
//                               case); train items are masked on the rare append path; full lists are
//                               compacted warp-cooperatively to the best K' (tau rises).
//      Invariant: every item that is NOT in a row's candidate list has approximate score <= tau_row
//      (or is a train item / NaN / out of range).
//   3. topn_tc_rescore_kernel: exact fp64 scores of the candidates in Java's summation order, top-N,
//      and an exactness certificate: N-th exact score > tau + error bound of the bf16 sweep, and no
//      exact ties among the first N+1.  Rows that fail are re-done by the exact kernel (topn_exact.cuh).
#pragma once
#include "lrk_common.cuh"
#include "topn_exact.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cudaTypedefs.h>
#include <cmath>
#include <algorithm>

#define TC_TILE_M 128          // users per accumulator
#define TC_UT 2                // user sub-tiles per CTA
#define TC_TILE_N 128          // items per MMA tile
#define TC_KB 64               // bf16 elements per 128-byte swizzle row
#define TC_KEEP 32             // K': candidates kept per (row, chunk) -- a min-heap in shared memory
#define TC_PEND 16             // pending hits parked per row before all lanes of the warp offer them together
#define TC_SLOTS (TC_KEEP + TC_PEND)
#define TC_CAP TC_KEEP         // candidate slots per (row, chunk) in global memory
#define TC_ROWS (TC_TILE_M * TC_UT)   // user rows per CTA
#define TC_THREADS 320
#define TC_MAX_CHUNKS 16

struct TcState {
    __nv_bfloat16* Bq = nullptr;      // [I x Kp]
    int Kp = 0;
    bool valid = false;
    double qnorm_max = 0.0, bi_max = 0.0;
    unsigned long long* d_stats = nullptr;   // [0] max ||q||^2 bits, [1] max |bi| bits
    PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    void* work = nullptr;                    // per-call scratch (user operand, candidate lists, ...), grown on demand
    size_t work_bytes = 0;
};

static inline TcState* tc_state(lrk_handle_s* h) {
    if (!h->tc) h->tc = new TcState();
    return (TcState*)h->tc;
}
static inline void topn_tc_release(lrk_handle_s* h) {
    TcState* s = (TcState*)h->tc;
    if (!s) return;
    cudaFree(s->Bq); cudaFree(s->d_stats); cudaFree(s->work);
    delete s;
    h->tc = nullptr;
}
static inline void topn_tc_invalidate(lrk_handle_s* h) {
    if (h->tc) ((TcState*)h->tc)->valid = false;
}
static inline int tc_kp(lrk_handle_s* h) {
    const int kaug = h->k + (h->cfg.model == LRK_MODEL_BIASEDMF ? 2 : 0);
    return ((kaug + TC_KB - 1) / TC_KB) * TC_KB;
}
static inline bool topn_tc_profitable(lrk_handle_s* h, int32_t nq, int topn) {
    return topn <= TC_KEEP / 2 && tc_kp(h) <= 128 && h->I >= 8192 && nq >= 512;
}

// ------------------------------------------------------------------------------------------------
// operand builders
// ------------------------------------------------------------------------------------------------
__global__ void tc_build_items_kernel(const double* __restrict__ Q, const double* __restrict__ bi, int biased, int k, int Kp,
                                      int32_t I, __nv_bfloat16* __restrict__ out, unsigned long long* __restrict__ stats) {
    const int32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= I) return;
    double n2 = 0.0;
    for (int f = lane; f < Kp; f += 32) {
        float v = 0.f;
        if (f < k) { const double q = Q[(int64_t)i * k + f]; n2 += q * q; v = (float)q; }
        else if (biased && f == k) v = __bfloat162float(__float2bfloat16_rn((float)bi[i]));
        else if (biased && f == k + 1) { const float b = (float)bi[i]; v = b - __bfloat162float(__float2bfloat16_rn(b)); }
        out[(int64_t)i * Kp + f] = __float2bfloat16_rn(v);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, m);
    if (lane == 0) {
        atomicMax(stats, (unsigned long long)__double_as_longlong(n2));
        if (biased) atomicMax(stats + 1, (unsigned long long)__double_as_longlong(fabs(bi[i])));
    }
}
__global__ void tc_build_users_kernel(const double* __restrict__ P, int biased, int k, int Kp, const int32_t* __restrict__ users,
                                      int32_t nq, __nv_bfloat16* __restrict__ out, double* __restrict__ pnorm) {
    const int32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= nq) return;
    const int32_t u = users ? users[c] : c;
    double n2 = 0.0;
    for (int f = lane; f < Kp; f += 32) {
        float v = 0.f;
        if (f < k) { const double p = P[(int64_t)u * k + f]; n2 += p * p; v = (float)p; }
        else if (biased && (f == k || f == k + 1)) v = 1.f;
        out[(int64_t)c * Kp + f] = __float2bfloat16_rn(v);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, m);
    if (lane == 0) pnorm[c] = sqrt(n2);
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);      // start address, 16-byte units
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, K-major both, N=128, M=128
#define TC_IDESC ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_TILE_N >> 3) << 17) | ((uint32_t)(TC_TILE_M >> 4) << 24))

struct TcParams {
    int32_t nq, I;
    int n_chunks, tiles_per_chunk, total_tiles, num_kb, stages;
    int exclude_train;
    const int64_t* __restrict__ rowptr;
    const int32_t* __restrict__ col;
    const int32_t* __restrict__ users;
    float* cand_score;      // [nq_pad][n_chunks][TC_CAP]
    int32_t* cand_item;
    int32_t* cand_cnt;      // [nq_pad][n_chunks]
    float* cand_tau;
};

// Offers (x, item) to a row's min-heap of TC_KEEP packed {score bits, item} entries in shared memory
// (slot stride TC_ROWS) after masking train items (MatrixRecommender.java:170-174: binary search in the
// sorted CSR row).  Returns the row's new threshold in the low word and "inserted" in bit 32.
__device__ __noinline__ unsigned long long tc_hit(uint2* ent, int size, float tau, float x, int32_t item,
                                                  const int32_t* __restrict__ col_row, int tlen) {
    if (tlen > 0) {
        int lo = 0, hi_ = tlen;
        while (lo < hi_) { const int mm = (lo + hi_) >> 1; if (__ldg(col_row + mm) < item) lo = mm + 1; else hi_ = mm; }
        if (lo < tlen && __ldg(col_row + lo) == item) return (unsigned long long)__float_as_uint(tau);
    }
    const uint2 me = make_uint2(__float_as_uint(x), (uint32_t)item);
    if (size < TC_KEEP) {
        int pos = size;
        while (pos > 0) {
            const int par = (pos - 1) >> 1;
            const uint2 pe = ent[par * TC_ROWS];
            if (!(x < __uint_as_float(pe.x))) break;
            ent[pos * TC_ROWS] = pe;
            pos = par;
        }
        ent[pos * TC_ROWS] = me;
        tau = (size + 1 == TC_KEEP) ? __uint_as_float(ent[0].x) : -INFINITY;
    } else {
        int pos = 0;
        for (;;) {
            int ch = 2 * pos + 1;
            if (ch >= TC_KEEP) break;
            uint2 ce = ent[ch * TC_ROWS];
            if (ch + 1 < TC_KEEP) { const uint2 re = ent[(ch + 1) * TC_ROWS]; if (__uint_as_float(re.x) < __uint_as_float(ce.x)) { ce = re; ++ch; } }
            if (!(__uint_as_float(ce.x) < x)) break;
            ent[pos * TC_ROWS] = ce;
            pos = ch;
        }
        ent[pos * TC_ROWS] = me;
        tau = __uint_as_float(ent[0].x);
    }
    return (1ull << 32) | (unsigned long long)__float_as_uint(tau);
}

// raw TMEM load of 32 accumulator columns of this thread's row (no wait)
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(TC_THREADS, 1)
topn_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    // carve: 1024-aligned operand tiles, then barriers
    unsigned char* base = (unsigned char*)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t tile_bytes = TC_TILE_M * TC_KB * 2;               // 16 KB: 128 rows x 128 B
    unsigned char* smA = base;                                       // [UT][num_kb] tiles
    unsigned char* smB = smA + (size_t)TC_UT * p.num_kb * tile_bytes; // [stages][num_kb] tiles
    uint2* heap_ent = (uint2*)(smB + (size_t)p.stages * p.num_kb * tile_bytes);   // [TC_SLOTS][TC_ROWS] {score bits, item}: heap, then pending
    uint64_t* bars = (uint64_t*)(heap_ent + TC_SLOTS * TC_ROWS);
    // barrier map: [0..S) b_full, [S..2S) b_empty, 2S a_full, 2S+1 a_empty, 2S+2.. tmem_full[2], tmem_empty[2]
    const int S = p.stages;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 6);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto BAR = [&](int i) { return smem_u32(bars + i); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(BAR(s), 1); mbar_init(BAR(S + s), 1); }
        mbar_init(BAR(2 * S), 1); mbar_init(BAR(2 * S + 1), 1);
        mbar_init(BAR(2 * S + 2), 1); mbar_init(BAR(2 * S + 3), 1);
        mbar_init(BAR(2 * S + 4), 8); mbar_init(BAR(2 * S + 5), 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m_tiles = (p.nq + TC_TILE_M * TC_UT - 1) / (TC_TILE_M * TC_UT);
    const int num_units = m_tiles * p.n_chunks;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int st = 0; uint32_t ph = 0, a_ph = 0;
            for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
                const int mt = unit / p.n_chunks, ch = unit - mt * p.n_chunks;
                mbar_wait(BAR(2 * S + 1), a_ph ^ 1);                       // A slot free
                mbar_expect_tx(BAR(2 * S), TC_UT * p.num_kb * tile_bytes);
                for (int a = 0; a < TC_UT; ++a)
                    for (int kb = 0; kb < p.num_kb; ++kb)
                        tma_load_2d(smem_u32(smA + (size_t)(a * p.num_kb + kb) * tile_bytes), &tmA, kb * TC_KB,
                                    (mt * TC_UT + a) * TC_TILE_M, BAR(2 * S));
                a_ph ^= 1;
                const int t0 = ch * p.tiles_per_chunk;
                const int t1 = min(p.total_tiles, t0 + p.tiles_per_chunk);
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(BAR(S + st), ph ^ 1);                        // B slot free
                    mbar_expect_tx(BAR(st), p.num_kb * tile_bytes);
                    for (int kb = 0; kb < p.num_kb; ++kb)
                        tma_load_2d(smem_u32(smB + (size_t)(st * p.num_kb + kb) * tile_bytes), &tmB, kb * TC_KB, t * TC_TILE_N, BAR(st));
                    if (++st == S) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            int st = 0; uint32_t ph = 0, a_ph = 0; int as = 0; uint32_t as_ph = 0;
            for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
                const int mt = unit / p.n_chunks, ch = unit - mt * p.n_chunks;
                (void)mt;
                mbar_wait(BAR(2 * S), a_ph);                               // A landed
                a_ph ^= 1;
                const int t0 = ch * p.tiles_per_chunk;
                const int t1 = min(p.total_tiles, t0 + p.tiles_per_chunk);
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(BAR(2 * S + 4 + as), as_ph ^ 1);             // accumulator stage drained
                    mbar_wait(BAR(st), ph);                                // B landed
                    tc_fence_after();
                    for (int a = 0; a < TC_UT; ++a) {
                        const uint32_t d = tmem_base + (uint32_t)(as * 256 + a * TC_TILE_N);
                        for (int kb = 0; kb < p.num_kb; ++kb) {
                            const uint32_t a_addr = smem_u32(smA + (size_t)(a * p.num_kb + kb) * tile_bytes);
                            const uint32_t b_addr = smem_u32(smB + (size_t)(st * p.num_kb + kb) * tile_bytes);
#pragma unroll
                            for (int k4 = 0; k4 < TC_KB / 16; ++k4)
                                tc_mma_f16(d, tc_smem_desc(a_addr + k4 * 32), tc_smem_desc(b_addr + k4 * 32), TC_IDESC,
                                           (kb | k4) != 0 ? 1u : 0u);
                        }
                    }
                    tc_commit(BAR(S + st));                                // B slot reusable once these MMAs retire
                    tc_commit(BAR(2 * S + 2 + as));                        // accumulators ready for the epilogue
                    if (++st == S) { st = 0; ph ^= 1; }
                    if (++as == 2) { as = 0; as_ph ^= 1; }
                }
                tc_commit(BAR(2 * S + 1));                                 // A slot reusable
            }
        }
    } else {
        // ================= epilogue: one thread == one user row =================
        const int ew = warp - 2;
        const int a = ew >> 2;                 // user sub-tile
        const int q = warp & 3;                // TMEM lane quarter this warp may read
        const int row_in_cta = a * TC_TILE_M + q * 32 + lane;
        int as = 0; uint32_t as_ph = 0;
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
            const int mt = unit / p.n_chunks, ch = unit - mt * p.n_chunks;
            const int32_t c = mt * (TC_TILE_M * TC_UT) + row_in_cta;
            const bool valid = c < p.nq;
            const int t0 = ch * p.tiles_per_chunk;
            const int t1 = min(p.total_tiles, t0 + p.tiles_per_chunk);
            const int32_t i1 = min(p.I, t1 * TC_TILE_N);
            uint2* ent = heap_ent + row_in_cta;                                  // this row's heap (slot stride TC_ROWS)
            const uint32_t pend_base = smem_u32(ent + TC_KEEP * TC_ROWS);       // this row's pending slots (shared-space address)
            const int32_t* col_row = p.col;
            int tlen = 0;
            if (valid && p.exclude_train) {
                const int32_t u = p.users ? p.users[c] : c;
                const int64_t rb = p.rowptr[u], re = p.rowptr[u + 1];
                col_row = p.col + rb; tlen = (int)(re - rb);
            }
            float tau = -INFINITY;
            int cnt = 0, pend = 0;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * TC_TILE_N);
            // Common path per 32-column slab: 31 FMNMX + 4 warp votes, no divergence.  A group of 8 columns is
            // inspected (8 predicated appends into the row's pending slots) only if SOME lane of the warp has a
            // score above its threshold there.  Pending hits are offered to the heaps by all 32 lanes in
            // lock-step rounds (TC_FLUSH), so the expensive part (train mask + heap sift) runs with many active
            // lanes.  The code is deliberately compact (rolled slab loop, one noinline tc_hit): the first
            // versions of this epilogue were instruction-cache bound.
#define TC_OFFER(X, ITEM)                                                                                          \
            do {                                                                                                   \
                const unsigned long long r_ = tc_hit(ent, cnt, tau, (X), (ITEM), col_row, tlen);                   \
                tau = __uint_as_float((uint32_t)r_);                                                               \
                if ((r_ >> 32) && cnt < TC_KEEP) ++cnt;                                                            \
            } while (0)
#define TC_FLUSH()                                                                                                 \
            do {                                                                                                   \
                for (int rr_ = 0; __any_sync(0xffffffffu, rr_ < pend); ++rr_) {                                    \
                    if (rr_ < pend) {                                                                              \
                        const uint2 pe_ = ent[(TC_KEEP + rr_) * TC_ROWS];                                          \
                        const float fx_ = __uint_as_float(pe_.x);                                                  \
                        if (fx_ > tau) TC_OFFER(fx_, (int32_t)pe_.y);                                              \
                    }                                                                                              \
                }                                                                                                  \
                pend = 0;                                                                                          \
            } while (0)
#define TC_CHECK1(R, J)                                                                                            \
            do {                                                                                                   \
                const float x_ = __uint_as_float(R[J]);                                                            \
                if (valid && x_ > tau && item0 + (J) < i1) {                                                       \
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(pend_base + (uint32_t)pend * (TC_ROWS * 8)), \
                                 "r"(R[J]), "r"(item0 + (J)) : "memory");                                          \
                    ++pend;                                                                                        \
                }                                                                                                  \
            } while (0)
#define TC_MAX8(R, O) fmaxf(fmaxf(fmaxf(__uint_as_float(R[O]), __uint_as_float(R[O + 1])), fmaxf(__uint_as_float(R[O + 2]), __uint_as_float(R[O + 3]))), \
                            fmaxf(fmaxf(__uint_as_float(R[O + 4]), __uint_as_float(R[O + 5])), fmaxf(__uint_as_float(R[O + 6]), __uint_as_float(R[O + 7]))))
#define TC_GROUP(R, O)                                                                                             \
            do {                                                                                                   \
                if (__any_sync(0xffffffffu, valid && TC_MAX8(R, O) > tau)) {                                       \
                    if (__any_sync(0xffffffffu, pend > TC_PEND - 8)) TC_FLUSH();                                   \
                    TC_CHECK1(R, O); TC_CHECK1(R, O + 1); TC_CHECK1(R, O + 2); TC_CHECK1(R, O + 3);                \
                    TC_CHECK1(R, O + 4); TC_CHECK1(R, O + 5); TC_CHECK1(R, O + 6); TC_CHECK1(R, O + 7);            \
                }                                                                                                  \
            } while (0)
#define TC_PROCESS(R, CB)                                                                                          \
            do {                                                                                                   \
                const int32_t item0 = n0 + (CB) * 32;                                                              \
                TC_GROUP(R, 0); TC_GROUP(R, 8); TC_GROUP(R, 16); TC_GROUP(R, 24);                                  \
            } while (0)
            for (int t = t0; t < t1; ++t) {
                mbar_wait(BAR(2 * S + 2 + as), as_ph);
                tc_fence_after();
                const int32_t n0 = t * TC_TILE_N;
                const uint32_t tcol = trow + (uint32_t)(as * 256);
                uint32_t ra[32], rb_[32];
                tc_ld32_nowait(tcol, ra);
                tc_ld_wait();
#pragma unroll 1
                for (int pair = 0; pair < 2; ++pair) {
                    tc_ld32_nowait(tcol + (uint32_t)(pair * 64 + 32), rb_);
                    TC_PROCESS(ra, pair * 2);
                    tc_ld_wait();
                    if (pair == 0) tc_ld32_nowait(tcol + 64, ra);
                    else {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(BAR(2 * S + 4 + as));      // this warp has drained the stage
                    }
                    TC_PROCESS(rb_, pair * 2 + 1);
                    if (pair == 0) tc_ld_wait();
                }
                if (++as == 2) { as = 0; as_ph ^= 1; }
            }
            TC_FLUSH();
#undef TC_PROCESS
#undef TC_GROUP
#undef TC_MAX8
#undef TC_CHECK1
#undef TC_FLUSH
#undef TC_OFFER
            if (valid) {
                float* cs = p.cand_score + ((size_t)c * p.n_chunks + ch) * TC_CAP;
                int32_t* ci = p.cand_item + ((size_t)c * p.n_chunks + ch) * TC_CAP;
                for (int e = 0; e < cnt; ++e) { const uint2 he = ent[e * TC_ROWS]; cs[e] = __uint_as_float(he.x); ci[e] = (int32_t)he.y; }
                p.cand_cnt[(size_t)c * p.n_chunks + ch] = cnt;
                p.cand_tau[(size_t)c * p.n_chunks + ch] = tau;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// exact re-score + certificate, one warp per query slot
// ------------------------------------------------------------------------------------------------
#define TC_RS_WARPS 4
__global__ void __launch_bounds__(TC_RS_WARPS * 32) topn_tc_rescore_kernel(
    const double* __restrict__ P, const double* __restrict__ Q, const double* __restrict__ bu, const double* __restrict__ bi,
    double mu, int biased, int k, const int32_t* __restrict__ users, int32_t nq, int n_chunks, int topn,
    const float* __restrict__ cand_score, const int32_t* __restrict__ cand_item, const int32_t* __restrict__ cand_cnt,
    const float* __restrict__ cand_tau, const double* __restrict__ pnorm, double qnorm_max, double bi_max, double c_err,
    int32_t* __restrict__ out_items, double* __restrict__ out_scores, int32_t* __restrict__ out_counts,
    int32_t* __restrict__ fail_slots, int32_t* __restrict__ fail_users, int* __restrict__ fail_count) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int max_cand = n_chunks * TC_CAP;
    double* sv = reinterpret_cast<double*>(rs_smem) + (size_t)warp * max_cand;
    int32_t* si = reinterpret_cast<int32_t*>(reinterpret_cast<double*>(rs_smem) + (size_t)TC_RS_WARPS * max_cand) + (size_t)warp * max_cand;
    const int32_t c = blockIdx.x * TC_RS_WARPS + warp;
    if (c >= nq) return;
    const int32_t u = users ? users[c] : c;
    // gather candidates of all chunks
    int M = 0;
    float tau_max = -INFINITY;
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int cnt = cand_cnt[(size_t)c * n_chunks + ch];
        tau_max = fmaxf(tau_max, cand_tau[(size_t)c * n_chunks + ch]);
        for (int e = lane; e < cnt; e += 32) si[M + e] = cand_item[((size_t)c * n_chunks + ch) * TC_CAP + e];
        M += cnt;
    }
    __syncwarp();
    // exact scores in Java's order (DenseVector.java:104-111; BiasedMFRecommender.java:119)
    const double* pu = P + (int64_t)u * k;
    for (int e = lane; e < M; e += 32) {
        const int32_t it = si[e];
        double d = dot_lr_f64(pu, Q + (int64_t)it * k, k);
        if (biased) d = __dadd_rn(__dadd_rn(__dadd_rn(d, bu[u]), bi[it]), mu);
        sv[e] = d;
    }
    __syncwarp();
    bool ok = M >= topn;
    // selection of the best topn+1 by (Double.compareTo desc); a tie anywhere in that prefix fails the row
    double prev = 0.0;
    const int want = min(M, topn + 1);
    for (int r = 0; ok && r < want; ++r) {
        double bv = -INFINITY; int be = -1;
        for (int e = lane; e < M; e += 32) {
            const double x = sv[e];
            if (si[e] >= 0 && (be < 0 || jcompare(x, bv) > 0)) { bv = x; be = e; }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, m);
            const int oe = __shfl_xor_sync(0xffffffffu, be, m);
            if (oe >= 0 && (be < 0 || jcompare(ov, bv) > 0 || (jcompare(ov, bv) == 0 && oe < be))) { bv = ov; be = oe; }
        }
        if (be < 0 || bv != bv) { ok = false; break; }
        if (r > 0 && jcompare(bv, prev) == 0) { ok = false; break; }      // exact tie: heap-order semantics needed
        if (r < topn && lane == 0) { out_items[(int64_t)c * topn + r] = si[be]; out_scores[(int64_t)c * topn + r] = bv; }
        if (r == topn - 1) {
            // certificate: nothing outside the candidate lists can reach the N-th exact score
            const double ub = biased ? bu[u] : 0.0;
            const double scale = pnorm[c] * qnorm_max + bi_max;
            const double bound = c_err * scale + 1e-12 * (scale + fabs(ub) + fabs(mu) + fabs((double)tau_max));
            const double reach = (double)tau_max + bound + (biased ? (ub + mu) : 0.0);
            if (!(bv > reach)) { ok = false; break; }
        }
        prev = bv;
        __syncwarp();
        if (lane == 0) si[be] = -1 - si[be];     // taken
        __syncwarp();
    }
    if (lane == 0) {
        if (ok) out_counts[c] = topn;
        else {
            const int pos = atomicAdd(fail_count, 1);
            fail_slots[pos] = c; fail_users[pos] = u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
static int tc_make_map(lrk_handle_s* h, TcState* s, CUtensorMap* map, const void* gptr, int Kp, int64_t rows) {
    if (!s->encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        LRK_CUDA(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess)
            return lrk_fail(h, LRK_ERR_CUDA, "cuTensorMapEncodeTiled", "driver entry point not available", __FILE__, __LINE__);
        s->encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    }
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {TC_KB, TC_TILE_M};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = s->encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return lrk_fail(h, LRK_ERR_CUDA, "cuTensorMapEncodeTiled", "encode failed", __FILE__, __LINE__);
    return LRK_OK;
}

static int topn_tc_run(lrk_handle_s* h, const int32_t* d_users, int32_t nq, int topn, int exclude_train,
                       int32_t* d_items, double* d_scores, int32_t* d_counts) {
    TcState* s = tc_state(h);
    cudaStream_t st = h->stream;
    const bool biased = h->cfg.model == LRK_MODEL_BIASEDMF;
    const int Kp = tc_kp(h);
    const int num_kb = Kp / TC_KB;
    if (num_kb > 2 || topn > TC_KEEP / 2)
        return lrk_fail(h, LRK_ERR_INVALID, "lrk_topn", "tensor-core path supports k (+2 for BiasedMF) <= 128 and topn <= 16", __FILE__, __LINE__);
    // ---- item operand (cached until the factors change)
    if (!s->valid || s->Kp != Kp) {
        if (s->Bq) { cudaFree(s->Bq); s->Bq = nullptr; }
        LRK_CUDA(h, cudaMalloc((void**)&s->Bq, sizeof(__nv_bfloat16) * (size_t)h->I * Kp));
        if (!s->d_stats) LRK_CUDA(h, cudaMalloc((void**)&s->d_stats, 16));
        LRK_CUDA(h, cudaMemsetAsync(s->d_stats, 0, 16, st));
        tc_build_items_kernel<<<lrk_ceil_div(h->I, 8), 256, 0, st>>>(h->Q64, h->bi64, biased, h->k, Kp, h->I, s->Bq, s->d_stats);
        LRK_LAUNCH_CHECK(h);
        unsigned long long stats[2];
        LRK_CUDA(h, cudaMemcpyAsync(stats, s->d_stats, 16, cudaMemcpyDeviceToHost, st));
        LRK_CUDA(h, cudaStreamSynchronize(st));
        double n2, bm;
        memcpy(&n2, &stats[0], 8); memcpy(&bm, &stats[1], 8);
        s->qnorm_max = sqrt(n2); s->bi_max = biased ? bm : 0.0;
        s->Kp = Kp; s->valid = true;
    }
    // ---- work decomposition
    const int total_tiles = lrk_ceil_div(h->I, TC_TILE_N);
    const int m_tiles = lrk_ceil_div(nq, TC_TILE_M * TC_UT);
    int n_chunks = lrk_ceil_div(h->sm_count, m_tiles);
    n_chunks = std::max(1, std::min(n_chunks, std::min(TC_MAX_CHUNKS, std::max(1, total_tiles / 16))));
    const int tiles_per_chunk = lrk_ceil_div(total_tiles, n_chunks);
    n_chunks = lrk_ceil_div(total_tiles, tiles_per_chunk);
    const int64_t nq_pad = (int64_t)m_tiles * TC_TILE_M * TC_UT;
    const size_t tile_bytes = (size_t)TC_TILE_M * TC_KB * 2;
    const size_t a_bytes = (size_t)TC_UT * num_kb * tile_bytes, b_stage = (size_t)num_kb * tile_bytes;
    const size_t heap_bytes = (size_t)TC_SLOTS * TC_ROWS * 8;                  // per-row min-heaps + pending slots
    int stages = (int)((225 * 1024 + 512 - heap_bytes - a_bytes) / b_stage);
    stages = std::max(2, std::min(stages, 6));
    const size_t smem = 1024 + a_bytes + (size_t)stages * b_stage + heap_bytes + (2 * stages + 6) * 8 + 16;
    // ---- query operand + scratch
    int32_t *fi = nullptr, *fc = nullptr; double* fs = nullptr;
    int rc = LRK_OK;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_aq = up(sizeof(__nv_bfloat16) * (size_t)nq_pad * Kp), b_pn = up(sizeof(double) * (size_t)nq_pad);
    const size_t b_cs = up(sizeof(float) * (size_t)nq_pad * n_chunks * TC_CAP), b_ci = up(sizeof(int32_t) * (size_t)nq_pad * n_chunks * TC_CAP);
    const size_t b_cc = up(sizeof(int32_t) * (size_t)nq_pad * n_chunks), b_ct = up(sizeof(float) * (size_t)nq_pad * n_chunks);
    const size_t b_fl = up(sizeof(int32_t) * (size_t)nq);
    const size_t need = b_aq + b_pn + b_cs + b_ci + b_cc + b_ct + 2 * b_fl + 256;
    if (s->work_bytes < need) {
        if (s->work) { cudaFree(s->work); s->work = nullptr; s->work_bytes = 0; }
        LRK_CUDA(h, cudaMalloc(&s->work, need));
        s->work_bytes = need;
    }
    char* w = (char*)s->work;
    __nv_bfloat16* Aq = (__nv_bfloat16*)w; w += b_aq;
    double* pnorm = (double*)w; w += b_pn;
    float* cscore = (float*)w; w += b_cs;
    int32_t* citem = (int32_t*)w; w += b_ci;
    int32_t* ccnt = (int32_t*)w; w += b_cc;
    float* ctau = (float*)w; w += b_ct;
    int32_t* fail_slots = (int32_t*)w; w += b_fl;
    int32_t* fail_users = (int32_t*)w; w += b_fl;
    int* fail_count = (int*)w;
    cudaError_t e = cudaMemsetAsync(fail_count, 0, sizeof(int), st);
    if (e == cudaSuccess && nq_pad > nq) e = cudaMemsetAsync(Aq + (size_t)nq * Kp, 0, sizeof(__nv_bfloat16) * (size_t)(nq_pad - nq) * Kp, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(ccnt, 0, sizeof(int32_t) * (size_t)nq_pad * n_chunks, st);
    int nfail = 0;
    do {
        if (e != cudaSuccess) break;
        tc_build_users_kernel<<<lrk_ceil_div(nq, 8), 256, 0, st>>>(h->P64, biased, h->k, Kp, d_users, nq, Aq, pnorm);
        h->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        CUtensorMap tmA, tmB;
        if ((rc = tc_make_map(h, s, &tmA, Aq, Kp, nq_pad))) break;
        if ((rc = tc_make_map(h, s, &tmB, s->Bq, Kp, h->I))) break;
        TcParams p;
        memset(&p, 0, sizeof p);
        p.nq = nq; p.I = h->I; p.n_chunks = n_chunks; p.tiles_per_chunk = tiles_per_chunk; p.total_tiles = total_tiles;
        p.num_kb = num_kb; p.stages = stages; p.exclude_train = exclude_train ? 1 : 0;
        p.rowptr = h->d_rowptr; p.col = h->d_col; p.users = d_users;
        p.cand_score = cscore; p.cand_item = citem; p.cand_cnt = ccnt; p.cand_tau = ctau;
        if ((e = cudaFuncSetAttribute(topn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) break;
        const int grid = std::min(h->sm_count, m_tiles * n_chunks);
        topn_tc_kernel<<<grid, TC_THREADS, smem, st>>>(tmA, tmB, p);
        h->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        // ---- exact re-score + certificate
        const double c_err = ldexp(1.0, -7) * (1.0 + ldexp(1.0, -6)) + (double)(Kp + 8) * ldexp(1.0, -22);
        const size_t rs_smem_bytes = (size_t)TC_RS_WARPS * n_chunks * TC_CAP * (sizeof(double) + sizeof(int32_t));
        if ((e = cudaFuncSetAttribute(topn_tc_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes)) != cudaSuccess) break;
        topn_tc_rescore_kernel<<<lrk_ceil_div(nq, TC_RS_WARPS), TC_RS_WARPS * 32, rs_smem_bytes, st>>>(
            h->P64, h->Q64, h->bu64, h->bi64, h->mu, biased, h->k, d_users, nq, n_chunks, topn, cscore, citem, ccnt, ctau, pnorm,
            s->qnorm_max, s->bi_max, c_err, d_items, d_scores, d_counts, fail_slots, fail_users, fail_count);
        h->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(&nfail, fail_count, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
        if (nfail > 0) {
            // rows without a certificate: item-parallel exact fp64 path (heap replay only for exact ties)
            if ((e = cudaMalloc((void**)&fi, sizeof(int32_t) * (size_t)nfail * topn)) != cudaSuccess) break;
            if ((e = cudaMalloc((void**)&fs, sizeof(double) * (size_t)nfail * topn)) != cudaSuccess) break;
            if ((e = cudaMalloc((void**)&fc, sizeof(int32_t) * (size_t)nfail)) != cudaSuccess) break;
            if ((rc = topn_exact_parallel_launch(h, fail_users, nfail, topn, exclude_train, fi, fs, fc))) break;
            topn_scatter_kernel<<<lrk_ceil_div((int64_t)nfail * topn, 256), 256, 0, st>>>(fail_slots, nfail, topn, fi, fs, fc,
                                                                                       d_items, d_scores, d_counts);
            h->launches++;
            if ((e = cudaGetLastError()) != cudaSuccess) break;
            if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
        }
    } while (0);
    if (e == cudaSuccess && rc == LRK_OK) e = cudaStreamSynchronize(st);
    cudaFree(fi); cudaFree(fs); cudaFree(fc);
    if (rc) return rc;
    LRK_CUDA(h, e);
    h->topn_fast_users = nq - nfail;
    h->topn_fallback_users = nfail;
    return LRK_OK;
}
