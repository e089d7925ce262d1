// bc.cu -- K7 of SURVEY.md §2: the ghost-node updates of the semidiscretisation RHS.
//
// Replaces examples/adv_diff_test.jl:118-141 (set-up: W_g = D_b[bc_b, ghost_b], its dense inverse, and
// W_int = D_b[bc_b, all columns except ghost_b]) and :162-176 (per RHS call:
//     u[ghost_b] = -inv(W_g) * (W_int * u[non-ghost_b])      for the "solve" boundaries,
//     u[bc_b] = u[ghost_b] = value                           for the Dirichlet boundary).
// The set-up is cold (once per operator; the small dense inverse is formed on the host exactly as the
// reference does with inv(Array(w_bc_g))).  The per-call part runs on the device: one masked row product per
// boundary node straight from the fixed-row CSR (no sliced copies), then one small dense GEMV.
#include <cmath>
#include <cstdlib>
#include <new>
#include <vector>

#include "common.cuh"

struct rbffd_bc {
    rbffd_operator* op = nullptr;
    int nb = 0;
    std::vector<int> kind, which, nbc, off;     // per boundary
    std::vector<double> value;
    int32_t* bc_idx = nullptr;      // [sum nbc] device
    int32_t* ghost_idx = nullptr;   // [sum nbc] device
    int8_t* ghost_of = nullptr;     // [N] device: 1 + boundary id of a ghost node, else 0
    double* winv = nullptr;         // concatenated dense inverses (row-major), device
    std::vector<size_t> winv_off;
    double* tbuf = nullptr;         // [max nbc] device
};

namespace {

// t[i] = sum_j [ghost_of[col] != tag] * vals[row_i][j] * u[col]     (W_int * u[non-ghost], adv_diff_test.jl:162)
__global__ void bc_rows_kernel(int nbc, int n, const int32_t* __restrict__ rows, const int32_t* __restrict__ colind,
                               const double* __restrict__ vals, const int8_t* __restrict__ ghost_of, int tag,
                               const double* __restrict__ u, double* __restrict__ t) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nbc) return;
    const int64_t base = (int64_t)rows[i] * n;
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) {
        const int c = colind[base + j];
        if (ghost_of[c] != tag) acc += vals[base + j] * u[c];
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) t[i] = acc;
}

// u[ghost[i]] = -sum_k winv[i][k] * t[k]
__global__ void bc_solve_kernel(int nbc, const double* __restrict__ winv, const double* __restrict__ t,
                                const int32_t* __restrict__ ghost, double* __restrict__ u) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nbc) return;
    double acc = 0.0;
    for (int k = lane; k < nbc; k += 32) acc += winv[(size_t)i * nbc + k] * t[k];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) u[ghost[i]] = -acc;
}

__global__ void bc_set_kernel(int nbc, const int32_t* __restrict__ a, const int32_t* __restrict__ b, double value, double* __restrict__ u) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nbc) { u[a[i]] = value; u[b[i]] = value; }
}

// All boundaries in ONE launch (one CTA): the boundaries depend on each other through corner stencils, so they run one after
// the other inside the CTA with block-wide barriers in between; per boundary the arithmetic is that of bc_rows_kernel /
// bc_solve_kernel (one warp per boundary node, lanes stride 32), so the result is bit-identical to the launch-per-kernel path.
constexpr int BC_FUSED_MAX_B = 16, BC_FUSED_MAX_NBC = 1024, BC_FUSED_THREADS = 512;
struct BcFused {
    int nb;
    int kind[BC_FUSED_MAX_B], cnt[BC_FUSED_MAX_B], off[BC_FUSED_MAX_B];
    const double* vals[BC_FUSED_MAX_B];
    const double* winv[BC_FUSED_MAX_B];
    double value[BC_FUSED_MAX_B];
};
__global__ void __launch_bounds__(BC_FUSED_THREADS) bc_fused_kernel(BcFused f, int n, const int32_t* __restrict__ bc_idx, const int32_t* __restrict__ ghost_idx,
                                                                    const int32_t* __restrict__ colind, const int8_t* __restrict__ ghost_of, double* u) {
    __shared__ double t[BC_FUSED_MAX_NBC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = BC_FUSED_THREADS / 32;
    for (int b = 0; b < f.nb; ++b) {
        const int cnt = f.cnt[b];
        const int32_t* rows = bc_idx + f.off[b];
        const int32_t* gh = ghost_idx + f.off[b];
        if (f.kind[b] == 1) {
            const double* vals = f.vals[b];
            for (int i = warp; i < cnt; i += nwarps) {
                const int64_t base = (int64_t)rows[i] * n;
                double acc = 0.0;
                for (int j = lane; j < n; j += 32) {
                    const int c = colind[base + j];
                    if (ghost_of[c] != b + 1) acc += vals[base + j] * u[c];
                }
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) t[i] = acc;
            }
            __syncthreads();
            const double* winv = f.winv[b];
            for (int i = warp; i < cnt; i += nwarps) {
                double acc = 0.0;
                for (int k = lane; k < cnt; k += 32) acc += winv[(size_t)i * cnt + k] * t[k];
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) u[gh[i]] = -acc;
            }
        } else {
            for (int i = threadIdx.x; i < cnt; i += BC_FUSED_THREADS) { u[rows[i]] = f.value[b]; u[gh[i]] = f.value[b]; }
        }
        __syncthreads();             // this boundary's ghost values are visible to the next one (same CTA: no fence needed beyond the barrier)
    }
}

// dense inverse by Gauss-Jordan with partial pivoting (host, cold path); returns false if singular
bool invert_dense(std::vector<double>& A, int n) {
    std::vector<double> B((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) B[(size_t)i * n + i] = 1.0;
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = std::fabs(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) if (std::fabs(A[(size_t)i * n + k]) > best) { best = std::fabs(A[(size_t)i * n + k]); p = i; }
        if (!(best > 0.0)) return false;
        if (p != k) for (int j = 0; j < n; ++j) { std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]); std::swap(B[(size_t)k * n + j], B[(size_t)p * n + j]); }
        const double r = 1.0 / A[(size_t)k * n + k];
        for (int j = 0; j < n; ++j) { A[(size_t)k * n + j] *= r; B[(size_t)k * n + j] *= r; }
        for (int i = 0; i < n; ++i) {
            if (i == k) continue;
            const double l = A[(size_t)i * n + k];
            if (l == 0.0) continue;
            for (int j = 0; j < n; ++j) { A[(size_t)i * n + j] -= l * A[(size_t)k * n + j]; B[(size_t)i * n + j] -= l * B[(size_t)k * n + j]; }
        }
    }
    A.swap(B);
    return true;
}

}  // namespace

extern "C" {

int rbffd_bc_destroy(rbffd_bc* bc) {
    if (!bc) return RBFFD_OK;
    if (bc->op && bc->op->ctx) { cudaSetDevice(bc->op->ctx->device); cudaStreamSynchronize(bc->op->ctx->stream); }
    cudaFree(bc->bc_idx); cudaFree(bc->ghost_idx); cudaFree(bc->ghost_of); cudaFree(bc->winv); cudaFree(bc->tbuf);
    delete bc;
    return RBFFD_OK;
}

int rbffd_bc_create(rbffd_operator* op, int32_t nb, const int32_t* kind, const int32_t* which, const double* value,
                    const int64_t* ptr, const int64_t* bc_idx, const int64_t* ghost_idx, int32_t index_base, rbffd_bc** out) {
    if (!op) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = op->ctx;
    if (!out || nb < 1 || !kind || !which || !value || !ptr || !bc_idx || !ghost_idx) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "bc_create: NULL pointer");
    *out = nullptr;
    if (op->M != op->N) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "bc_create: needs square operators");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t total = ptr[nb] - ptr[0];
    rbffd_bc* bc = new (std::nothrow) rbffd_bc();
    if (!bc) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "out of host memory");
    bc->op = op; bc->nb = nb;
    std::vector<int32_t> hb(total), hg(total);
    std::vector<int8_t> gof(op->N, 0);
    int maxn = 1;
    for (int b = 0; b < nb; ++b) {
        const int64_t lo = ptr[b] - ptr[0], cnt = ptr[b + 1] - ptr[b];
        bc->kind.push_back(kind[b]); bc->which.push_back(which[b]); bc->value.push_back(value[b]);
        bc->nbc.push_back((int)cnt); bc->off.push_back((int)lo);
        if (kind[b] == 1 && (which[b] < 0 || which[b] >= op->nmat)) { rbffd_bc_destroy(bc); RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "bc_create: matrix index out of range"); }
        for (int64_t i = 0; i < cnt; ++i) {
            const int64_t r = bc_idx[lo + i] - index_base, g = ghost_idx[lo + i] - index_base;
            if (r < 0 || r >= op->N || g < 0 || g >= op->N) { rbffd_bc_destroy(bc); RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "bc_create: node index out of range"); }
            hb[lo + i] = (int32_t)r; hg[lo + i] = (int32_t)g;
            gof[g] = (int8_t)(b + 1);
        }
        maxn = std::max(maxn, (int)cnt);
    }
    // W_g = D_b[bc_b, ghost_b] from the device CSR rows, inverted on the host (adv_diff_test.jl:126-141)
    std::vector<double> winv_all;
    const int n = op->n;
    for (int b = 0; b < nb; ++b) {
        bc->winv_off.push_back(winv_all.size());
        if (kind[b] != 1) continue;
        const int cnt = bc->nbc[b], lo = bc->off[b];
        std::vector<int32_t> hc((size_t)cnt * n);
        std::vector<double> hv((size_t)cnt * n);
        const double* vals = op->vals + (size_t)op->M * n * which[b];
        for (int i = 0; i < cnt; ++i) {
            cudaError_t e = cudaMemcpyAsync(hc.data() + (size_t)i * n, op->colind + (size_t)hb[lo + i] * n, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(hv.data() + (size_t)i * n, vals + (size_t)hb[lo + i] * n, sizeof(double) * n, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) { rbffd_bc_destroy(bc); RBFFD_FAIL(ctx, RBFFD_ERR_CUDA, "bc_create: %s", cudaGetErrorString(e)); }
        }
        cudaStreamSynchronize(st);
        std::vector<int> local(op->N, -1);
        for (int i = 0; i < cnt; ++i) local[hg[lo + i]] = i;
        std::vector<double> W((size_t)cnt * cnt, 0.0);
        for (int i = 0; i < cnt; ++i)
            for (int j = 0; j < n; ++j) {
                const int c = hc[(size_t)i * n + j];
                if (gof[c] == b + 1 && local[c] >= 0) W[(size_t)i * cnt + local[c]] += hv[(size_t)i * n + j];
            }
        if (!invert_dense(W, cnt)) { rbffd_bc_destroy(bc); RBFFD_FAIL(ctx, RBFFD_ERR_SINGULAR, "bc_create: ghost block of boundary %d is singular", b); }
        winv_all.insert(winv_all.end(), W.begin(), W.end());
    }
    cudaError_t e = cudaMalloc((void**)&bc->bc_idx, sizeof(int32_t) * std::max<int64_t>(total, 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&bc->ghost_idx, sizeof(int32_t) * std::max<int64_t>(total, 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&bc->ghost_of, std::max<int64_t>(op->N, 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&bc->winv, sizeof(double) * std::max<size_t>(winv_all.size(), 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&bc->tbuf, sizeof(double) * maxn);
    if (e == cudaSuccess) e = cudaMemcpyAsync(bc->bc_idx, hb.data(), sizeof(int32_t) * total, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(bc->ghost_idx, hg.data(), sizeof(int32_t) * total, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(bc->ghost_of, gof.data(), op->N, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && !winv_all.empty()) e = cudaMemcpyAsync(bc->winv, winv_all.data(), sizeof(double) * winv_all.size(), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { rbffd_bc_destroy(bc); RBFFD_FAIL(ctx, RBFFD_ERR_CUDA, "bc_create: %s", cudaGetErrorString(e)); }
    *out = bc;
    return RBFFD_OK;
}

// boundaries are applied in the order they were given (the reference: right, left(Dirichlet), top, bottom; :162-176)
int rbffd_bc_apply_device(rbffd_bc* bc, double* u) {
    if (!bc) return RBFFD_ERR_INVALID;
    rbffd_operator* op = bc->op;
    rbffd_context* ctx = op->ctx;
    if (!u) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "bc_apply: NULL pointer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int maxn = 0;
    for (int b = 0; b < bc->nb; ++b) maxn = std::max(maxn, bc->nbc[b]);
    static const bool no_fuse = [] { const char* e = getenv("RBFFD_BC_FUSED"); return e && atoi(e) == 0; }();
    if (bc->nb <= BC_FUSED_MAX_B && maxn <= BC_FUSED_MAX_NBC && !no_fuse) {
        BcFused f{};
        f.nb = bc->nb;
        for (int b = 0; b < bc->nb; ++b) {
            f.kind[b] = bc->kind[b]; f.cnt[b] = bc->nbc[b]; f.off[b] = bc->off[b]; f.value[b] = bc->value[b];
            f.vals[b] = bc->kind[b] == 1 ? op->vals + (size_t)op->M * op->n * bc->which[b] : nullptr;
            f.winv[b] = bc->winv + bc->winv_off[b];
        }
        bc_fused_kernel<<<1, BC_FUSED_THREADS, 0, st>>>(f, op->n, bc->bc_idx, bc->ghost_idx, op->colind, bc->ghost_of, u);
        KLAUNCH(ctx);
        CUDA_TRY(ctx, cudaGetLastError());
        return RBFFD_OK;
    }
    for (int b = 0; b < bc->nb; ++b) {
        const int cnt = bc->nbc[b];
        if (cnt == 0) continue;
        const int32_t* rows = bc->bc_idx + bc->off[b];
        const int32_t* gh = bc->ghost_idx + bc->off[b];
        if (bc->kind[b] == 1) {
            const double* vals = op->vals + (size_t)op->M * op->n * bc->which[b];
            bc_rows_kernel<<<(cnt * 32 + 127) / 128, 128, 0, st>>>(cnt, op->n, rows, op->colind, vals, bc->ghost_of, b + 1, u, bc->tbuf);
            bc_solve_kernel<<<(cnt * 32 + 127) / 128, 128, 0, st>>>(cnt, bc->winv + bc->winv_off[b], bc->tbuf, gh, u);
            KLAUNCH(ctx); KLAUNCH(ctx);
        } else {
            bc_set_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(cnt, rows, gh, bc->value[b], u);
            KLAUNCH(ctx);
        }
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

int rbffd_bc_apply_host(rbffd_bc* bc, double* u) {
    if (!bc) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = bc->op->ctx;
    if (!u) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "bc_apply: NULL pointer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf<double> d;
    CUDA_TRY(ctx, d.alloc(bc->op->N, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d.p, u, sizeof(double) * bc->op->N, cudaMemcpyHostToDevice, ctx->stream));
    RBFFD_TRY(rbffd_bc_apply_device(bc, d.p));
    CUDA_TRY(ctx, cudaMemcpyAsync(u, d.p, sizeof(double) * bc->op->N, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RBFFD_OK;
}

}  // extern "C"
