// Staging kernels: flat CSR (what the Java shim flattens SequentialAccessSparseMatrix into,
// math/structure/SequentialAccessSparseMatrix.java:23, RowSequentialAccessSparseMatrix.java:19)
// -> device CSR + shuffled COO stream; DenseMatrix double[][] <-> padded fp32 working rows.
#pragma once
#include "lrk_common.cuh"
#include <cub/cub.cuh>

__device__ __forceinline__ uint32_t lrk_hash32(uint64_t x) {   // splitmix64 finaliser, high word
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (uint32_t)(x >> 32);
}

// entry e -> (row, shuffle key).  Row found by binary search in rowptr (robust to skewed rows).
__global__ void coo_expand_kernel(const int64_t* __restrict__ rowptr, int32_t U, int64_t nnz, uint64_t seed,
                                  int32_t* __restrict__ row_of, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    int32_t lo = 0, hi = U;
    while (hi - lo > 1) { const int32_t m = (lo + hi) >> 1; if (rowptr[m] <= e) lo = m; else hi = m; }
    row_of[e] = lo;
    keys[e] = lrk_hash32((uint64_t)e ^ (seed * 0xD6E8FEB86659FD93ull));
    idx[e] = (uint32_t)e;
}

__global__ void coo_gather_kernel(const uint32_t* __restrict__ perm, const int32_t* __restrict__ row_of,
                                  const int32_t* __restrict__ col, const double* __restrict__ val, int64_t nnz,
                                  int32_t* __restrict__ su, int32_t* __restrict__ si, float* __restrict__ sr) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nnz) return;
    const uint32_t e = perm[t];
    su[t] = row_of[e]; si[t] = col[e]; sr[t] = (float)val[e];
}

// 0 = ok; bit0: rowptr not monotone / bad ends, bit1: column out of range, bit2: row not strictly ascending
__global__ void csr_validate_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int32_t U,
                                    int32_t I, int64_t nnz, int* __restrict__ flags) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < U) { if (rowptr[t + 1] < rowptr[t]) atomicOr(flags, 1); }
    if (t == 0) { if (rowptr[0] != 0 || rowptr[U] != nnz) atomicOr(flags, 1); }
    if (t < nnz) {
        const int32_t c = col[t];
        if (c < 0 || c >= I) atomicOr(flags, 2);
    }
}
__global__ void csr_validate_rows_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                         const int32_t* __restrict__ row_of, int64_t nnz, int* __restrict__ flags) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t + 1 < nnz && row_of[t] == row_of[t + 1] && col[t] >= col[t + 1]) atomicOr(flags, 4);
}

// double[rows][k] -> float[rows][ld] (zero padded), and back
__global__ void f64_to_f32_rows_kernel(const double* __restrict__ src, float* __restrict__ dst, int64_t rows, int k, int ld) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * ld) return;
    const int64_t r = t / ld; const int f = (int)(t - r * ld);
    dst[t] = f < k ? (float)src[r * k + f] : 0.f;
}
__global__ void f32_to_f64_rows_kernel(const float* __restrict__ src, double* __restrict__ dst, int64_t rows, int k, int ld) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * k) return;
    const int64_t r = t / k; const int f = (int)(t - r * k);
    dst[t] = (double)src[r * ld + f];
}

// Host driver: device CSR (already resident: d_rowptr, d_col) + host values -> shuffled COO stream.
// The shuffle is a stable radix sort of the entries by a 32-bit hash of (entry index, seed).
static int stage_coo_from_csr(lrk_handle_s* h, const int64_t* d_rowptr, const int32_t* d_col, const double* h_val,
                              int32_t U, int32_t I, int64_t nnz, int32_t* su, int32_t* si, float* sr, bool validate) {
    cudaStream_t st = h->stream;
    if (nnz == 0) { LRK_CUDA(h, cudaStreamSynchronize(st)); return LRK_OK; }
    size_t tmp_bytes = 0;
    LRK_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                                (uint32_t*)nullptr, (int)nnz, 0, 32, st));
    const size_t n = (size_t)nnz;
    LrkScratch sc;
    int rc = lrk_scratch_begin(h, n * (8 + 4 * 5) + tmp_bytes + 16 * 256, &sc);
    if (rc) return rc;
    double* d_val = sc.take<double>(n);
    int32_t* row_of = sc.take<int32_t>(n);
    uint32_t *keys = sc.take<uint32_t>(n), *idx = sc.take<uint32_t>(n), *keys2 = sc.take<uint32_t>(n), *perm = sc.take<uint32_t>(n);
    void* tmp = sc.take<char>(tmp_bytes ? tmp_bytes : 1);
    int* d_flags = sc.take<int>(1);
    if (!d_val || !row_of || !keys || !idx || !keys2 || !perm || !tmp || !d_flags)
        return lrk_fail(h, LRK_ERR_NOMEM, "stage_coo_from_csr", "scratch arena too small", __FILE__, __LINE__);
    int flags = 0;
    LRK_CUDA(h, cudaMemcpyAsync(d_val, h_val, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    LRK_CUDA(h, cudaMemsetAsync(d_flags, 0, sizeof(int), st));
    const int nb = lrk_ceil_div(nnz, 256);
    coo_expand_kernel<<<nb, 256, 0, st>>>(d_rowptr, U, nnz, h->cfg.seed, row_of, keys, idx);
    LRK_LAUNCH_CHECK(h);
    if (validate) {
        const int64_t m = nnz > U ? nnz : U;
        csr_validate_kernel<<<lrk_ceil_div(m, 256), 256, 0, st>>>(d_rowptr, d_col, U, I, nnz, d_flags);
        LRK_LAUNCH_CHECK(h);
        csr_validate_rows_kernel<<<nb, 256, 0, st>>>(d_rowptr, d_col, row_of, nnz, d_flags);
        LRK_LAUNCH_CHECK(h);
    }
    LRK_CUDA(h, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, idx, perm, (int)nnz, 0, 32, st));
    coo_gather_kernel<<<nb, 256, 0, st>>>(perm, row_of, d_col, d_val, nnz, su, si, sr);
    LRK_LAUNCH_CHECK(h);
    LRK_CUDA(h, cudaMemcpyAsync(&flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    LRK_CUDA(h, cudaStreamSynchronize(st));
    if (flags & 1) return lrk_fail(h, LRK_ERR_INVALID, "lrk_set_train_csr", "rowptr is not a monotone prefix sum ending at nnz", __FILE__, __LINE__);
    if (flags & 2) return lrk_fail(h, LRK_ERR_INVALID, "lrk_set_train_csr", "column index out of range", __FILE__, __LINE__);
    if (flags & 4) return lrk_fail(h, LRK_ERR_INVALID, "lrk_set_train_csr", "columns must be strictly ascending inside a row", __FILE__, __LINE__);
    return LRK_OK;
}
