// spmv.cu -- K5/K6 of SURVEY.md §2: application of the fixed-row-length CSR operators.
//
// Replaces the SparseMatrixCSC products of the time-stepping RHS (examples/adv_diff_test.jl:151-152):
//   D*u                      -> spmv_kernel            (coalesced value/index streams, sub-warp team per row)
//   (a*Dxx + a*Dyy - ...)*u  -> spmv_multi_kernel      (all operators share ONE colind: one gather of u
//                                                       serves every matrix, no scalar*sparse temporaries)
//   E' * v                   -> spmv_t_csc_kernel      (deterministic, over a lazily built CSC copy: no atomics)
// Bytes per row (algorithmic): 12 n + 16 for one matrix (8n values + 4n int32 indices + x + y).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "spmv_rows.cuh"

namespace {

template <int TPR, int NMAT, int VEC, int ITERS>
__global__ void __launch_bounds__(256) spmv_multi_kernel(int64_t M, int n, const int32_t* __restrict__ colind, SpmvMats m,
                                                         const double* __restrict__ x, SpmvEpilogue ep, double* __restrict__ y) {
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    spmv_rows<TPR, NMAT, VEC, ITERS, false>(warp, 0, M, n, colind, m, x, nullptr, 0, ep, y);
}

// y[c] = alpha * sum_{entries e in column c} vals[e] * v[row(e)] + beta * y[c]: 8 lanes per column, entries of a column are
// contiguous in the CSC view.  Owned operators read a CSC copy of the values (t_vals: three coalesced streams + the gather of
// v, as the forward product); borrowed ones gather the caller's values through t_src.
template <int KU>
__global__ void spmv_t_csc_kernel(int64_t N, const int32_t* __restrict__ t_ptr, const int32_t* __restrict__ t_row,
                                  const double* __restrict__ t_vals, const int32_t* __restrict__ t_src, const double* __restrict__ vals,
                                  const double* __restrict__ v, double alpha, double beta, double* __restrict__ y) {
    const int lane = threadIdx.x & 7;
    const int64_t col = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
    double acc0 = 0.0, acc1 = 0.0;
    if (col < N) {
        const int b = t_ptr[col], e = t_ptr[col + 1];
        for (int i0 = b + lane; i0 < e; i0 += 8 * KU) {  // KU entries per lane in flight, all loads before the first use
            int r[KU];
            double w[KU], x[KU];
#pragma unroll
            for (int k = 0; k < KU; ++k) {
                const int i = i0 + 8 * k;
                const bool in = i < e;
                r[k] = in ? t_row[i] : -1;
                w[k] = in ? (t_vals ? __ldcs(t_vals + i) : __ldg(vals + t_src[i])) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < KU; ++k) x[k] = r[k] >= 0 ? __ldg(v + r[k]) : 0.0;
#pragma unroll
            for (int k = 0; k < KU; k += 2) {
                acc0 = fma(w[k], x[k], acc0);
                acc1 = fma(w[k + 1], x[k + 1], acc1);
            }
        }
    }
    double acc = acc0 + acc1;
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (col < N && lane == 0) y[col] = beta == 0.0 ? alpha * acc : alpha * acc + beta * y[col];
}

__global__ void transpose_rows_kernel(const int32_t* __restrict__ t_src, int64_t nnz, int n, int32_t* __restrict__ t_row) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nnz) t_row[i] = t_src[i] / n;
}

__global__ void transpose_vals_kernel(const int32_t* __restrict__ t_src, const double* __restrict__ vals, int64_t nnz, double* __restrict__ t_vals) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nnz) t_vals[i] = vals[t_src[i]];
}

__global__ void iota_kernel2(int* p, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int)i;
}

__global__ void seg_start_kernel2(const int* __restrict__ key_sorted, int64_t n, int nseg, int* __restrict__ start) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i > n) return;
    int prev = i == 0 ? -1 : key_sorted[i - 1];
    int cur = i == n ? nseg : key_sorted[i];
    for (int c = prev + 1; c <= cur; ++c) start[c] = (int)i;
}

__global__ void gather_kernel(const double* __restrict__ src, const int32_t* __restrict__ index, int64_t count, double* __restrict__ dst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < count) dst[i] = src[index[i]];
}

__global__ void scatter_add_kernel(const double* __restrict__ src, const int32_t* __restrict__ index, int64_t count, double* __restrict__ dst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < count) atomicAdd(dst + index[i], src[i]);
}

template <int TPR, int VEC, int ITERS>
int launch_multi(rbffd_operator* op, int nm, const SpmvMats& m, const double* x, const SpmvEpilogue& ep, double* y) {
    rbffd_context* ctx = op->ctx;
    const int rows_per_block = 256 / TPR;
    const int grid = ceil_div_i64(op->M, rows_per_block);
    cudaStream_t st = ctx->stream;
    switch (nm) {
        case 1: spmv_multi_kernel<TPR, 1, VEC, ITERS><<<grid, 256, 0, st>>>(op->M, op->n, op->colind, m, x, ep, y); break;
        case 2: spmv_multi_kernel<TPR, 2, VEC, ITERS><<<grid, 256, 0, st>>>(op->M, op->n, op->colind, m, x, ep, y); break;
        case 3: spmv_multi_kernel<TPR, 3, VEC, ITERS><<<grid, 256, 0, st>>>(op->M, op->n, op->colind, m, x, ep, y); break;
        case 4: spmv_multi_kernel<TPR, 4, VEC, ITERS><<<grid, 256, 0, st>>>(op->M, op->n, op->colind, m, x, ep, y); break;
        case 5: spmv_multi_kernel<TPR, 5, VEC, ITERS><<<grid, 256, 0, st>>>(op->M, op->n, op->colind, m, x, ep, y); break;
        default: spmv_multi_kernel<TPR, 6, VEC, ITERS><<<grid, 256, 0, st>>>(op->M, op->n, op->colind, m, x, ep, y); break;
    }
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

// row length -> (team size, vector width, unrolled iterations); TPR*VEC*ITERS >= n
int dispatch_multi(rbffd_operator* op, int nm, const SpmvMats& m, const double* x, const SpmvEpilogue& ep, double* y) {
    const int n = op->n;
    bool even = (n % 2) == 0;            // rows start 16-byte aligned when the planes are (n*8 bytes per row)
    for (int i = 0; i < nm; ++i) even = even && (reinterpret_cast<uintptr_t>(m.v[i]) % 16 == 0);
    even = even && (reinterpret_cast<uintptr_t>(op->colind) % 8 == 0);
    if (even) {
        if (n <= 8) return launch_multi<4, 2, 1>(op, nm, m, x, ep, y);
        if (n <= 16) return launch_multi<4, 2, 2>(op, nm, m, x, ep, y);
        if (n <= 32) return launch_multi<8, 2, 2>(op, nm, m, x, ep, y);
        if (n <= 48) return launch_multi<8, 2, 3>(op, nm, m, x, ep, y);
        if (n <= 64) return launch_multi<8, 2, 4>(op, nm, m, x, ep, y);
        if (n <= 128) return launch_multi<16, 2, 4>(op, nm, m, x, ep, y);
        if (n <= 256) return launch_multi<32, 2, 4>(op, nm, m, x, ep, y);
    } else {
        if (n <= 8) return launch_multi<4, 1, 2>(op, nm, m, x, ep, y);
        if (n <= 16) return launch_multi<4, 1, 4>(op, nm, m, x, ep, y);
        if (n <= 32) return launch_multi<8, 1, 4>(op, nm, m, x, ep, y);
        if (n <= 64) return launch_multi<8, 1, 8>(op, nm, m, x, ep, y);
        if (n <= 128) return launch_multi<16, 1, 8>(op, nm, m, x, ep, y);
        if (n <= 256) return launch_multi<32, 1, 8>(op, nm, m, x, ep, y);
    }
    RBFFD_FAIL(op->ctx, RBFFD_ERR_UNSUPPORTED, "spmv: row length %d > 256", n);
}

}  // namespace

// y = sum_i coef[i] * D[which[i]] * x + beta * y ; terms are fused six at a time over the shared pattern
int rbffd_spmv_multi_impl(rbffd_operator* op, int nterms, const int32_t* which, const double* coef, const double* x,
                          double beta, double* y) {
    rbffd_context* ctx = op->ctx;
    if (nterms < 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "spmv: need at least one term");
    for (int i = 0; i < nterms; ++i)
        if (which[i] < 0 || which[i] >= op->nmat) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "spmv: matrix index %d out of range (nmat=%d)", which[i], op->nmat);
    if (op->M == 0) return RBFFD_OK;
    const size_t stride = (size_t)op->M * op->n;
    for (int i0 = 0; i0 < nterms; i0 += SPMV_MAXMAT) {
        const int nm = std::min(SPMV_MAXMAT, nterms - i0);
        SpmvMats m{};
        for (int i = 0; i < nm; ++i) { m.v[i] = op->vals + stride * which[i0 + i]; m.c[i] = coef[i0 + i]; }
        const SpmvEpilogue ep{i0 == 0 ? beta : 1.0, nullptr, 0.0, 0.0, 0.0};
        RBFFD_TRY(dispatch_multi(op, nm, m, x, ep, y));
    }
    return RBFFD_OK;
}

// out = a * u + b * (x + dt * sum_i coef[i] * D[which[i]] * x): one SSP-RK stage as ONE launch (rows = nodes, M == N, <= 6 terms)
int rbffd_spmv_stage_impl(rbffd_operator* op, int nterms, const int32_t* which, const double* coef, const double* x,
                          double a, const double* u, double b, double dt, double* out) {
    rbffd_context* ctx = op->ctx;
    if (nterms < 1 || nterms > SPMV_MAXMAT) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "spmv_stage: 1..%d terms (combine the operators first)", SPMV_MAXMAT);
    if (op->M != op->N) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "spmv_stage: the stage update needs rows = nodes (M == N)");
    if (out == x || out == u) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "spmv_stage: out must not alias x or u (other rows still gather from them)");
    for (int i = 0; i < nterms; ++i)
        if (which[i] < 0 || which[i] >= op->nmat) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "spmv_stage: matrix index %d out of range (nmat=%d)", which[i], op->nmat);
    if (op->M == 0) return RBFFD_OK;
    const size_t stride = (size_t)op->M * op->n;
    SpmvMats m{};
    for (int i = 0; i < nterms; ++i) { m.v[i] = op->vals + stride * which[i]; m.c[i] = coef[i]; }
    const SpmvEpilogue ep{0.0, u, a, b, dt};
    return dispatch_multi(op, nterms, m, x, ep, out);
}

int rbffd_build_transpose(rbffd_operator* op) {
    if (op->t_ptr) return RBFFD_OK;
    rbffd_context* ctx = op->ctx;
    cudaStream_t st = ctx->stream;
    const int64_t nnz = op->M * op->n;
    if (nnz > 0x7fffffff) RBFFD_FAIL(ctx, RBFFD_ERR_UNSUPPORTED, "transpose view needs nnz < 2^31 per shard");
    DevBuf<int> ident, keys_sorted, src, ptr, rowid;
    CUDA_TRY(ctx, ident.alloc(nnz, st));
    CUDA_TRY(ctx, keys_sorted.alloc(nnz, st));
    CUDA_TRY(ctx, src.alloc(nnz, st));
    CUDA_TRY(ctx, ptr.alloc(op->N + 1, st));
    iota_kernel2<<<ceil_div_i64(nnz, 256), 256, 0, st>>>(ident.p, nnz);
    int bits = 1;
    while ((1ll << bits) < op->N) ++bits;
    size_t tmp_bytes = 0;
    CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, op->colind, keys_sorted.p, ident.p, src.p, (int)nnz, 0, bits, st));
    DevBuf<unsigned char> tmp;
    CUDA_TRY(ctx, tmp.alloc(tmp_bytes, st));
    CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, op->colind, keys_sorted.p, ident.p, src.p, (int)nnz, 0, bits, st));
    seg_start_kernel2<<<ceil_div_i64(nnz + 1, 256), 256, 0, st>>>(keys_sorted.p, nnz, (int)op->N, ptr.p);
    CUDA_TRY(ctx, rowid.alloc(nnz, st));
    transpose_rows_kernel<<<ceil_div_i64(std::max<int64_t>(nnz, 1), 256), 256, 0, st>>>(src.p, nnz, op->n, rowid.p);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    op->t_ptr = ptr.release();
    op->t_src = src.release();
    op->t_row = rowid.release();
    return RBFFD_OK;
}

int rbffd_spmv_t_impl(rbffd_operator* op, int which, double alpha, const double* v, double beta, double* y) {
    rbffd_context* ctx = op->ctx;
    if (which < 0 || which >= op->nmat) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "spmv_t: matrix index %d out of range", which);
    RBFFD_TRY(rbffd_build_transpose(op));
    if (op->N == 0) return RBFFD_OK;
    const double* vals = op->vals + (size_t)op->M * op->n * which;
    const int64_t nnz = op->M * (int64_t)op->n;
    const double* tv = nullptr;
    if (!op->borrowed && nnz > 0) {                      // values in column order, built once per matrix
        if (op->t_vals.empty()) op->t_vals.assign(op->nmat, nullptr);
        if (!op->t_vals[which]) {
            double* p = nullptr;
            CUDA_TRY(ctx, cudaMalloc((void**)&p, sizeof(double) * (size_t)nnz));
            transpose_vals_kernel<<<ceil_div_i64(nnz, 256), 256, 0, ctx->stream>>>(op->t_src, vals, nnz, p);
            KLAUNCH(ctx);
            op->t_vals[which] = p;
        }
        tv = op->t_vals[which];
    }
    // columns hold about n entries: two per lane cover n <= 16 ... 40 in one or two trips, four per lane the larger stencils
    if (op->n <= 40) spmv_t_csc_kernel<2><<<ceil_div_i64(op->N, 256 / 8), 256, 0, ctx->stream>>>(op->N, op->t_ptr, op->t_row, tv, op->t_src, vals, v, alpha, beta, y);
    else spmv_t_csc_kernel<4><<<ceil_div_i64(op->N, 256 / 8), 256, 0, ctx->stream>>>(op->N, op->t_ptr, op->t_row, tv, op->t_src, vals, v, alpha, beta, y);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

int rbffd_gather_impl(rbffd_context* ctx, const double* src, const int32_t* index, int64_t count, double* dst) {
    if (count == 0) return RBFFD_OK;
    gather_kernel<<<ceil_div_i64(count, 256), 256, 0, ctx->stream>>>(src, index, count, dst);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

int rbffd_scatter_add_impl(rbffd_context* ctx, const double* src, const int32_t* index, int64_t count, double* dst) {
    if (count == 0) return RBFFD_OK;
    scatter_add_kernel<<<ceil_div_i64(count, 256), 256, 0, ctx->stream>>>(src, index, count, dst);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}
