// api.cu -- the extern "C" surface of librbffd.so (declared in include/rbffd.h).
#include <cmath>
#include <cstring>
#include <new>

#include <chrono>
#include <cstdlib>
#include <thread>
#include <emmintrin.h>
#include <sched.h>

#include "common.cuh"

int rbffd_spmv_multi_impl(rbffd_operator* op, int nterms, const int32_t* which, const double* coef, const double* x,
                          double beta, double* y);
int rbffd_spmv_t_impl(rbffd_operator* op, int which, double alpha, const double* v, double beta, double* y);
int rbffd_spmv_stage_impl(rbffd_operator* op, int nterms, const int32_t* which, const double* coef, const double* x,
                          double a, const double* u, double b, double dt, double* out);
int rbffd_gather_impl(rbffd_context* ctx, const double* src, const int32_t* index, int64_t count, double* dst);
int rbffd_scatter_add_impl(rbffd_context* ctx, const double* src, const int32_t* index, int64_t count, double* dst);

namespace {

thread_local std::string g_err_noctx;

__global__ void i32_to_i64_kernel(const int32_t* __restrict__ in, int64_t n, int64_t base, int64_t* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int64_t)in[i] + base;
}

__global__ void i32_add_base_kernel(const int32_t* __restrict__ in, int64_t n, int32_t base, int32_t* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + base;
}

__global__ void i64_to_i32_kernel(const int64_t* __restrict__ in, int64_t n, int64_t base, int32_t* __restrict__ out, int64_t hi, int* bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        int64_t v = in[i] - base;
        if (v < 0 || v >= hi) *bad = 1;
        out[i] = (int32_t)v;
    }
}

__global__ void sqrt_kernel(double* p, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = sqrt(p[i]);
}

__device__ __forceinline__ double lattice_uniform(uint64_t seed, uint64_t lin, int axis) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (lin * 3ull + (uint64_t)axis + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * 0x1.0p-53;
}

__global__ void lattice_kernel(int dim, int64_t g, uint64_t seed, int64_t first, int64_t count, double* __restrict__ X) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int64_t lin = first + t;
    int64_t c[3] = {lin % g, (lin / g) % g, lin / (g * g)};
    if (dim == 2) c[1] = lin / g;
    for (int a = 0; a < dim; ++a) {
        double u = lattice_uniform(seed, (uint64_t)lin, a);
        X[t * dim + a] = ((double)c[a] + 0.5 + 0.5 * (u - 0.5)) / (double)g;
    }
}

int copy_indices_to_host(rbffd_context* ctx, const int32_t* dev, int64_t count, int64_t base, int64_t* host) {
    if (!host || count == 0) return RBFFD_OK;
    DevBuf<int64_t> wide;
    CUDA_TRY(ctx, wide.alloc(count, ctx->stream));
    i32_to_i64_kernel<<<ceil_div_i64(count, 256), 256, 0, ctx->stream>>>(dev, count, base, wide.p);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(host, wide.p, sizeof(int64_t) * count, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RBFFD_OK;
}

}  // namespace

extern "C" {

int rbffd_version(void) { return 100; }

int rbffd_create(int device, rbffd_context** out) {
    if (!out) return RBFFD_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_err_noctx = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); librbffd has no CPU fallback";
        return RBFFD_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        g_err_noctx = "device ordinal out of range";
        return RBFFD_ERR_INVALID;
    }
    rbffd_context* ctx = new (std::nothrow) rbffd_context();
    if (!ctx) return RBFFD_ERR_INVALID;
    ctx->device = device;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->chunk_ev[i], cudaEventDisableTiming);
    if (e == cudaSuccess) {
        ctx->owned_stream = ctx->stream;
        for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    }
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->hflags, 16 * sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&ctx->hflags_dev, ctx->hflags, 0);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e == cudaSuccess) {
        // keep freed temporaries in the pool: repeated generate calls do not hit the driver allocator
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thresh = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
        }
    }
    if (e != cudaSuccess) {
        g_err_noctx = std::string("CUDA error during rbffd_create: ") + cudaGetErrorString(e);
        delete ctx;
        return RBFFD_ERR_CUDA;
    }
    *out = ctx;
    return RBFFD_OK;
}

int rbffd_destroy(rbffd_context* ctx) {
    if (!ctx) return RBFFD_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    dev_block_cache().clear();                    // parked temporaries go back to the pool before their streams disappear
    for (int i = 0; i < 8; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 4; ++i) if (ctx->chunk_ev[i]) cudaEventDestroy(ctx->chunk_ev[i]);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->hflags) cudaFreeHost(ctx->hflags);
    if (ctx->stage_i32) cudaFreeHost(ctx->stage_i32);
    if (ctx->owned_stream) cudaStreamDestroy(ctx->owned_stream);
    delete ctx;
    return RBFFD_OK;
}

const char* rbffd_last_error(const rbffd_context* ctx) { return ctx ? ctx->err.c_str() : g_err_noctx.c_str(); }

int rbffd_set_stream(rbffd_context* ctx, void* cuda_stream) {
    if (!ctx) return RBFFD_ERR_INVALID;
    cudaSetDevice(ctx->device);
    ctx->stream = (cudaStream_t)cuda_stream;      // borrowed; the context's own stream is kept for rbffd_reset_stream
    return RBFFD_OK;
}

int rbffd_get_stream(rbffd_context* ctx, void** cuda_stream) {
    if (!ctx || !cuda_stream) return RBFFD_ERR_INVALID;
    *cuda_stream = (void*)ctx->stream;
    return RBFFD_OK;
}

int rbffd_reset_stream(rbffd_context* ctx) {
    if (!ctx) return RBFFD_ERR_INVALID;
    ctx->stream = ctx->owned_stream;
    return RBFFD_OK;
}

int rbffd_synchronize(rbffd_context* ctx) {
    if (!ctx) return RBFFD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RBFFD_OK;
}

long long rbffd_launch_count(const rbffd_context* ctx) { return ctx ? ctx->launches : 0; }

int rbffd_timings(rbffd_context* ctx, double* ms, int n) {
    if (!ctx || !ms) return RBFFD_ERR_INVALID;
    for (int i = 0; i < n && i < 8; ++i) ms[i] = ctx->timings[i];
    return RBFFD_OK;
}

int rbffd_knn_device(rbffd_context* ctx, const double* X, int64_t N, int32_t dim, const double* Q, int64_t NQ, int32_t k,
                     const int32_t* xgroup, const int32_t* qgroup, int32_t* idx_out, double* d2_out) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (!X || !idx_out) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "knn: NULL pointer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const bool q_is_x = (Q == nullptr) || (Q == X && NQ == N && qgroup == xgroup);
    return rbffd_knn_impl(ctx, X, N, dim, Q, NQ, k, xgroup, qgroup, q_is_x, idx_out, d2_out);
}

int rbffd_calculateneighbors_host(rbffd_context* ctx, const double* X, int64_t N, const double* Y, int64_t M,
                                  int32_t dim, int32_t n, const int32_t* xgroup, int32_t index_base,
                                  int64_t* idxs_x, int64_t* idxs_y_x, double* dists_x, double* dists_y_x) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (!X || N < 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "calculateneighbors: X is empty");
    if (dim < 1 || dim > 3) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "dim must be 1..3");
    if (index_base != 0 && index_base != 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "index_base must be 0 or 1");
    if (!Y) { Y = X; M = N; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf<double> dX, dY, d2x, d2y;
    DevBuf<int32_t> dG, dS, dC;
    CUDA_TRY(ctx, dX.alloc((size_t)N * dim, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(dX.p, X, sizeof(double) * N * dim, cudaMemcpyHostToDevice, st));
    const double* dYp = dX.p;
    if (Y != X) {
        CUDA_TRY(ctx, dY.alloc((size_t)M * dim, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(dY.p, Y, sizeof(double) * M * dim, cudaMemcpyHostToDevice, st));
        dYp = dY.p;
    }
    if (xgroup) {
        CUDA_TRY(ctx, dG.alloc(N, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(dG.p, xgroup, sizeof(int32_t) * N, cudaMemcpyHostToDevice, st));
    }
    const bool want_x = idxs_x || dists_x, want_y = idxs_y_x || dists_y_x;
    if (want_x) { CUDA_TRY(ctx, dS.alloc((size_t)N * n, st)); if (dists_x) CUDA_TRY(ctx, d2x.alloc((size_t)N * n, st)); }
    if (want_y) { CUDA_TRY(ctx, dC.alloc(M, st)); if (dists_y_x) CUDA_TRY(ctx, d2y.alloc(M, st)); }
    RBFFD_TRY(rbffd_stencils_impl(ctx, dX.p, N, dim, dYp, M, n, xgroup ? dG.p : nullptr, want_x ? dS.p : nullptr,
                                  dists_x ? d2x.p : nullptr, want_y ? dC.p : nullptr, dists_y_x ? d2y.p : nullptr));
    if (dists_x) {
        sqrt_kernel<<<ceil_div_i64(N * n, 256), 256, 0, st>>>(d2x.p, N * n);
        CUDA_TRY(ctx, cudaMemcpyAsync(dists_x, d2x.p, sizeof(double) * N * n, cudaMemcpyDeviceToHost, st));
    }
    if (dists_y_x) {
        sqrt_kernel<<<ceil_div_i64(M, 256), 256, 0, st>>>(d2y.p, M);
        CUDA_TRY(ctx, cudaMemcpyAsync(dists_y_x, d2y.p, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
    }
    if (idxs_x) RBFFD_TRY(copy_indices_to_host(ctx, dS.p, N * n, index_base, idxs_x));
    if (idxs_y_x) RBFFD_TRY(copy_indices_to_host(ctx, dC.p, M, index_base, idxs_y_x));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return RBFFD_OK;
}

int rbffd_stencils_device(rbffd_context* ctx, const double* X, int64_t N, int32_t dim, const double* Y, int64_t M,
                          int32_t n, const int32_t* xgroup, int32_t* stencils_out, int32_t* center_out) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (!X) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "stencils: NULL pointer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!Y) { Y = X; M = N; }
    return rbffd_stencils_impl(ctx, X, N, dim, Y, M, n, xgroup, stencils_out, nullptr, center_out, nullptr);
}

int rbffd_operator_from_device(rbffd_context* ctx, int64_t M, int64_t N, int32_t n, int32_t nmat, const int32_t* colind,
                               const double* vals, rbffd_operator** out) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (!out || !colind || !vals) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "operator_from_device: NULL pointer");
    *out = nullptr;
    if (M < 0 || N < 1 || n < 1 || nmat < 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "operator_from_device: bad sizes");
    rbffd_operator* op = new (std::nothrow) rbffd_operator();
    if (!op) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "out of host memory");
    op->ctx = ctx; op->M = M; op->N = N; op->n = n; op->nmat = nmat;
    op->colind = const_cast<int32_t*>(colind);
    op->vals = const_cast<double*>(vals);
    op->borrowed = true;
    *out = op;
    return RBFFD_OK;
}

int rbffd_weights_device(rbffd_context* ctx, const rbffd_options* opts, const double* X, int64_t N, const double* Y,
                         int64_t M, const int32_t* stencils, int64_t NS, const int32_t* center, int32_t* colind_out, double* vals_out) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (!X || !stencils || !colind_out || !vals_out) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "weights: NULL pointer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!Y) { Y = X; M = N; }
    if (NS <= 0) NS = N;
    return rbffd_weights_impl(ctx, opts, X, N, Y, M, stencils, NS, center, colind_out, vals_out);
}

namespace {
__global__ void init_deferred_kernel(int* f) { f[threadIdx.x] = (threadIdx.x & 7) == 0 ? 0x7fffffff : 0; }
__global__ void pack_deferred_kernel(const int* f, int nchunks, int* out) {
    if ((int)threadIdx.x < nchunks) { out[2 * threadIdx.x] = f[8 * threadIdx.x]; out[2 * threadIdx.x + 1] = f[8 * threadIdx.x + 4]; }
}
}  // namespace

int rbffd_generate_operator_host(rbffd_context* ctx, const rbffd_options* opts, const double* X, int64_t N,
                                 const double* Y, int64_t M, const int32_t* xgroup, int64_t* colind_out, double* vals_out) {
    if (!ctx) return RBFFD_ERR_INVALID;
    RBFFD_TRY(rbffd_validate_options(ctx, opts));
    if (!X || N < 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "generate_operator: X is empty");
    if (!colind_out || !vals_out) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "generate_operator: NULL output buffer");
    if (!Y) { Y = X; M = N; }
    if (opts->variant != 0 && (Y != X || M != N)) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "the legacy collocated variant takes X only (pass Y = NULL)");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int dim = opts->dim, n = opts->n;
    DevBuf<double> dX, dY;
    DevBuf<int32_t> dG;
    CUDA_TRY(ctx, dX.alloc((size_t)N * dim, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(dX.p, X, sizeof(double) * N * dim, cudaMemcpyHostToDevice, st));
    const double* dYp = dX.p;
    if (Y != X) {
        CUDA_TRY(ctx, dY.alloc((size_t)M * dim, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(dY.p, Y, sizeof(double) * M * dim, cudaMemcpyHostToDevice, st));
        dYp = dY.p;
    }
    if (xgroup) {
        CUDA_TRY(ctx, dG.alloc(N, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(dG.p, xgroup, sizeof(int32_t) * N, cudaMemcpyHostToDevice, st));
    }
    // Collocated rows with pinned output buffers run as a pipeline: the column indices are known as soon as
    // the neighbour search is done (they ARE the stencils), so they leave right away on a second stream under the weight
    // solve; the rows are solved in a few shrinking chunks queued back to back, each followed on the copy stream by its
    // values; status words are collected per chunk on the device and read once at the end.  Anything else takes the
    // one-shot path below.
    cudaPointerAttributes pa_c, pa_v;
    const bool pinned = cudaPointerGetAttributes(&pa_c, colind_out) == cudaSuccess && pa_c.type == cudaMemoryTypeHost &&
                        cudaPointerGetAttributes(&pa_v, vals_out) == cudaSuccess && pa_v.type == cudaMemoryTypeHost;
    cudaGetLastError();
    constexpr int MAXCH = 8;
    // RBFFD_HOST_CHUNKS (2..8) / RBFFD_HOST_CHUNK_RATIO: tuning knobs of the row-chunk pipeline (default: the measured optimum)
    static const int nchunks_env = [] { const char* e = getenv("RBFFD_HOST_CHUNKS"); const int v = e ? atoi(e) : 0; return v >= 2 && v <= MAXCH ? v : 0; }();
    static const double ratio_env = [] { const char* e = getenv("RBFFD_HOST_CHUNK_RATIO"); const double v = e ? atof(e) : 0.0; return v > 0.05 && v <= 1.0 ? v : 0.62; }();
    const int nchunks = nchunks_env ? nchunks_env : 5;
    if (pinned && Y == X && M == N && M >= 64 * nchunks && !opts->sort_columns) {
        const bool trace = getenv("RBFFD_TRACE") != nullptr;
        const auto t_begin = std::chrono::steady_clock::now();
        auto lap = [&](const char* what) {
            if (trace) fprintf(stderr, "[rbffd trace] %-28s %8.3f ms\n", what,
                               std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
        };
        for (int i = 0; i < 8; ++i) ctx->timings[i] = 0.0;
        const int nops = opts->nops;
        DevBuf<int32_t> stencils;
        DevBuf<int64_t> c64;
        CUDA_TRY(ctx, stencils.alloc((size_t)N * n, st));
        RBFFD_TRY(rbffd_stencils_impl(ctx, dX.p, N, dim, dX.p, N, n, xgroup ? dG.p : nullptr, stencils.p, nullptr, nullptr, nullptr));
        // Row chunks shrink geometrically: big chunks first (few launch tails), small ones last (the copy of the last
        // chunk is the only one that nothing overlaps).
        int64_t cbeg[MAXCH + 1];
        {
            double frac[MAXCH] = {0.38, 0.28, 0.18, 0.11, 0.05, 0, 0, 0};
            if (nchunks_env) {                                  // geometric: frac[k] ~ ratio^k
                double tot = 0.0, w = 1.0;
                for (int k = 0; k < nchunks; ++k) { frac[k] = w; tot += w; w *= ratio_env; }
                for (int k = 0; k < nchunks; ++k) frac[k] /= tot;
            }
            double acc = 0.0;
            cbeg[0] = 0;
            for (int k = 0; k < nchunks; ++k) {
                acc += frac[k];
                cbeg[k + 1] = k + 1 == nchunks ? M : std::min<int64_t>(M, (int64_t)(acc * (double)M) / 32 * 32);
            }
        }
        int64_t ch = 0;
        for (int k = 0; k < nchunks; ++k) ch = std::max(ch, cbeg[k + 1] - cbeg[k]);
        DevBuf<int32_t> c32;                       // the weight kernels write the pattern too; here it is the stencil array itself
        DevBuf<double> vb;                         // every chunk has its own slice: the solve never waits for a copy
        DevBuf<int> dflags;                        // status words of every chunk, inspected once at the end
        CUDA_TRY(ctx, c32.alloc((size_t)ch * n, st));
        CUDA_TRY(ctx, vb.alloc((size_t)M * n * nops, st));
        CUDA_TRY(ctx, dflags.alloc(8 * MAXCH, st));
        // Declared after every temporary the copy stream reads, so it runs BEFORE their destructors on every exit path
        // (error returns included): the stream-ordered frees on `st` must not overtake D2H copies still in flight.
        struct DrainCopies { cudaStream_t s; ~DrainCopies() { cudaStreamSynchronize(s); } } drain_copies{ctx->copy_stream};
        init_deferred_kernel<<<1, 8 * MAXCH, 0, st>>>(dflags.p);
        KLAUNCH(ctx);
        // The pattern crosses PCIe as int32 (half the bytes of the caller's int64) into a pinned staging buffer, slice by
        // slice; host threads widen every slice into colind_out as soon as its copy has landed, under the value copies.
        // RBFFD_HOST_WIDEN=0 (or fewer than 8 hardware threads per local rank) widens on the device and ships int64.
        const char* hw_env = getenv("RBFFD_HOST_WIDEN");
        const char* wt_env = getenv("RBFFD_WIDEN_THREADS");
        const char* lw_env = getenv("LOCAL_WORLD_SIZE");            // one process per GPU: share the host cores
        const unsigned hc = std::thread::hardware_concurrency();
        const unsigned lw = lw_env ? (unsigned)std::max(1, atoi(lw_env)) : 1u;
        // hardware threads THIS process may use: its affinity mask when the launcher bound it to the cores next to its GPU
        // (rb.bind_to_gpu_numa), else an equal share of the host
        unsigned avail = hc / (2 * lw);
        {
            cpu_set_t set;
            CPU_ZERO(&set);
            if (sched_getaffinity(0, sizeof(set), &set) == 0) {
                const unsigned bound = (unsigned)CPU_COUNT(&set);
                if (bound > 0 && bound < hc) avail = bound > 1 ? bound - 1 : 1;      // bound process: all its threads but the caller's
            }
        }
        const int T = wt_env ? std::max(1, atoi(wt_env)) : (int)std::max(1u, std::min(8u, avail));
        // Several ranks on one host share its ingest bandwidth (measured at N = 2: 16.3 ms per call when every rank ships the
        // int64 pattern, 11.1 ms with the int32 pattern widened by 6 host threads per rank), so the int32 transfer is the
        // default at any rank count: the widening threads are capped by the cores of the rank, never replaced by an int64 copy.
        const bool want32 = opts->index_width == 32;
        bool host_widen = !want32 && (hw_env ? atoi(hw_env) != 0 : true);
        constexpr int NSL = 8;
        struct Wideners {                          // joined on every exit path (the threads only wait for queued copies)
            std::vector<std::thread> th;
            cudaEvent_t ev[NSL] = {};
            std::vector<cudaEvent_t> chunk_done;
            void finish() {
                for (auto& t : th) if (t.joinable()) t.join();
                th.clear();
                for (int k = 0; k < NSL; ++k) if (ev[k]) { cudaEventDestroy(ev[k]); ev[k] = nullptr; }
                for (auto e : chunk_done) cudaEventDestroy(e);
                chunk_done.clear();
            }
            ~Wideners() { finish(); }
        } wd;
        std::vector<std::thread>& wideners = wd.th;
        cudaEvent_t* sl_ev = wd.ev;
        const int64_t total = N * (int64_t)n;
        if (host_widen && ctx->stage_i32_count < (size_t)total) {
            if (ctx->stage_i32) cudaFreeHost(ctx->stage_i32);
            ctx->stage_i32 = nullptr; ctx->stage_i32_count = 0;
            if (cudaHostAlloc(&ctx->stage_i32, sizeof(int32_t) * (size_t)total, cudaHostAllocDefault) == cudaSuccess) ctx->stage_i32_count = (size_t)total;
            else { cudaGetLastError(); ctx->stage_i32 = nullptr; host_widen = false; }     // no pinned memory left: widen on the device
        }
        DevBuf<int32_t> c32b;
        if (want32) {
            // the caller keeps int32 indices (SparseMatrixCSC{Float64,Int32}): the stencils go straight into its buffer
            const int32_t* srcp = stencils.p;
            if (opts->index_base != 0) {
                CUDA_TRY(ctx, c32b.alloc((size_t)total, st));
                i32_add_base_kernel<<<ceil_div_i64(total, 256), 256, 0, st>>>(stencils.p, total, opts->index_base, c32b.p);
                KLAUNCH(ctx);
                srcp = c32b.p;
            }
            CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[2], st));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[2], 0));
            CUDA_TRY(ctx, cudaMemcpyAsync(reinterpret_cast<int32_t*>(colind_out), srcp, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, ctx->copy_stream));
        } else if (host_widen) {
            CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[2], st));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[2], 0));
            const int64_t sl = ((total + NSL - 1) / NSL + 63) / 64 * 64;
            for (int k = 0; k < NSL; ++k) {
                const int64_t b0 = std::min<int64_t>(total, k * sl), b1 = std::min<int64_t>(total, b0 + sl);
                CUDA_TRY(ctx, cudaEventCreateWithFlags(&sl_ev[k], cudaEventDisableTiming | cudaEventBlockingSync));
                if (b1 > b0) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage_i32 + b0, stencils.p + b0, sizeof(int32_t) * (b1 - b0), cudaMemcpyDeviceToHost, ctx->copy_stream));
                CUDA_TRY(ctx, cudaEventRecord(sl_ev[k], ctx->copy_stream));
            }
            const int32_t* src = ctx->stage_i32;
            const int64_t base = opts->index_base;
            const int dev = ctx->device;
            for (int t = 0; t < T; ++t)
                wideners.emplace_back([=]() {
                    cudaSetDevice(dev);
                    for (int k = 0; k < NSL; ++k) {
                        const int64_t b0 = std::min<int64_t>(total, k * sl), b1 = std::min<int64_t>(total, b0 + sl);
                        cudaEventSynchronize(sl_ev[k]);
                        const int64_t len = b1 - b0, per = (len + T - 1) / T;
                        const int64_t e0 = b0 + std::min<int64_t>(len, t * per), e1 = b0 + std::min<int64_t>(len, (t + 1) * per);
                        // streaming stores: the int64 pattern is written once and not read here (no read-for-ownership)
                        for (int64_t e = e0; e < e1; ++e) _mm_stream_si64(reinterpret_cast<long long*>(colind_out + e), (long long)src[e] + base);
                    }
                    _mm_sfence();
                });
        } else {
            CUDA_TRY(ctx, c64.alloc((size_t)N * n, st));
            i32_to_i64_kernel<<<ceil_div_i64(N * n, 256), 256, 0, st>>>(stencils.p, N * n, opts->index_base, c64.p);
            KLAUNCH(ctx);
            CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[2], st));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[2], 0));
            CUDA_TRY(ctx, cudaMemcpyAsync(colind_out, c64.p, sizeof(int64_t) * N * n, cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
        lap("search + colind D2H queued");
        double t_weights = 0.0;
        int rc = RBFFD_OK;
        ctx->trusted_stencils = true;              // produced by our own search: no range check per chunk
        ctx->collocated_rows = true;               // Y == X: row r0 + i sits at the centre of stencil r0 + i
        ctx->deferred_flags = dflags.p;            // no host synchronisation per chunk: the kernels queue back to back
        struct Restore { rbffd_context* c; ~Restore() { c->deferred_flags = nullptr; c->trusted_stencils = false; c->collocated_rows = false; } } restore{ctx};
        std::vector<cudaEvent_t>& cev = wd.chunk_done;
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], st));
        const bool dbg_noship = getenv("RBFFD_DEBUG_NOSHIP") != nullptr;     // timing experiments only: results stay on the device
        auto ship = [&](int k) -> cudaError_t {    // values of chunk k -> host, on the copy stream
            if (dbg_noship) return cudaSuccess;
            const int64_t r0 = cbeg[k], cnt = cbeg[k + 1] - r0;
            const double* vchunk = vb.p + (size_t)r0 * n * nops;
            for (int o = 0; o < nops; ++o) {
                cudaError_t ce = cudaMemcpyAsync(vals_out + ((size_t)o * M + r0) * n, vchunk + (size_t)o * cnt * n, sizeof(double) * cnt * n,
                                                 cudaMemcpyDeviceToHost, ctx->copy_stream);
                if (ce != cudaSuccess) return ce;
            }
            return cudaSuccess;
        };
        for (int k = 0; k < nchunks && rc == RBFFD_OK; ++k) {
            const int64_t r0 = cbeg[k], cnt = cbeg[k + 1] - r0;
            if (cnt <= 0) continue;
            ctx->deferred_slot = k;
            rc = rbffd_weights_impl(ctx, opts, dX.p, N, dX.p + r0 * dim, cnt, stencils.p + r0 * n, cnt, nullptr, c32.p,
                                    vb.p + (size_t)r0 * n * nops);
            if (rc != RBFFD_OK) break;
            cudaEvent_t ev;
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            cev.push_back(ev);
            CUDA_TRY(ctx, cudaEventRecord(ev, st));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ev, 0));
            CUDA_TRY(ctx, ship(k));
        }
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], st));
        lap("all chunks queued");
        ctx->deferred_flags = nullptr;
        int hf[2 * MAXCH];
        if (rc == RBFFD_OK) {
            DevBuf<int> packed;
            CUDA_TRY(ctx, packed.alloc(2 * nchunks, st));
            pack_deferred_kernel<<<1, 32, 0, st>>>(dflags.p, nchunks, packed.p);
            KLAUNCH(ctx);
            CUDA_TRY(ctx, rbffd_fetch_flags(ctx, packed.p, 2 * nchunks, hf));     // synchronises `st`: every chunk is solved
            lap("all chunks solved");
            float ms = 0.f;
            CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
            t_weights = ms;
            for (int k = 0; k < nchunks && rc == RBFFD_OK; ++k) {
                if (hf[2 * k] == 0x7fffffff && hf[2 * k + 1] == 0) continue;
                // a stencil of this chunk was refused by the null-space kernel (or is singular): redo the chunk through the
                // synchronous path, which falls back to the pivoted kernels and reports singular nodes
                const int64_t r0 = cbeg[k], cnt = cbeg[k + 1] - r0;
                rc = rbffd_weights_impl(ctx, opts, dX.p, N, dX.p + r0 * dim, cnt, stencils.p + r0 * n, cnt, nullptr, c32.p,
                                        vb.p + (size_t)r0 * n * nops);
                if (rc != RBFFD_OK) break;
                CUDA_TRY(ctx, cudaEventRecord(ctx->chunk_ev[3], st));
                CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[3], 0));
                CUDA_TRY(ctx, ship(k));
            }
        }
        ctx->trusted_stencils = false;
        cudaError_t e = cudaStreamSynchronize(ctx->copy_stream);
        lap("copies drained");
        wd.finish();
        lap("pattern widened");
        if (trace) fprintf(stderr, "[rbffd trace] weight kernels (device time, all chunks) %8.3f ms\n", t_weights);
        ctx->timings[3] = t_weights;
        // the stream-ordered temporaries are freed on `st`: make sure the copy stream is done with them first
        if (rc != RBFFD_OK) return rc;
        CUDA_TRY(ctx, e);
        return RBFFD_OK;
    }
    rbffd_operator* op = nullptr;
    RBFFD_TRY(rbffd_operator_generate(ctx, opts, dX.p, N, dYp, M, xgroup ? dG.p : nullptr, &op));
    int rc = RBFFD_OK;
    if (opts->index_width == 32) {
        const int64_t nnz = op->M * op->n;
        DevBuf<int32_t> tmp;
        cudaError_t e = tmp.alloc((size_t)std::max<int64_t>(nnz, 1), st);
        if (e == cudaSuccess && nnz > 0) {
            i32_add_base_kernel<<<ceil_div_i64(nnz, 256), 256, 0, st>>>(op->colind, nnz, opts->index_base, tmp.p);
            KLAUNCH(ctx);
            e = cudaMemcpyAsync(reinterpret_cast<int32_t*>(colind_out), tmp.p, sizeof(int32_t) * (size_t)nnz, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(vals_out, op->vals, sizeof(double) * (size_t)nnz * op->nmat, cudaMemcpyDeviceToHost, st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { ctx->err = std::string("CUDA error in generate_operator (int32 output): ") + cudaGetErrorString(e); rc = RBFFD_ERR_CUDA; }
    } else {
        rc = rbffd_operator_to_host(op, opts->index_base, colind_out, vals_out);
    }
    rbffd_operator_destroy(op);
    return rc;
}

int rbffd_operator_generate(rbffd_context* ctx, const rbffd_options* opts, const double* X, int64_t N, const double* Y,
                            int64_t M, const int32_t* xgroup, rbffd_operator** out) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (!out) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "operator_generate: NULL output handle");
    *out = nullptr;
    RBFFD_TRY(rbffd_validate_options(ctx, opts));
    if (!X || N < 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "operator_generate: X is empty");
    if (!Y) { Y = X; M = N; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = opts->n;
    for (int i = 0; i < 8; ++i) ctx->timings[i] = 0.0;
    DevBuf<int32_t> stencils, center;
    CUDA_TRY(ctx, stencils.alloc((size_t)N * n, st));
    CUDA_TRY(ctx, center.alloc(M, st));
    RBFFD_TRY(rbffd_stencils_impl(ctx, X, N, opts->dim, Y, M, n, xgroup, stencils.p, nullptr, center.p, nullptr));
    rbffd_operator* op = new (std::nothrow) rbffd_operator();
    if (!op) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "out of host memory");
    op->ctx = ctx; op->M = M; op->N = N; op->n = n; op->nmat = opts->nops;
    cudaError_t e = cudaMalloc((void**)&op->colind, sizeof(int32_t) * (size_t)std::max<int64_t>(M * n, 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&op->vals, sizeof(double) * (size_t)std::max<int64_t>(M * n, 1) * opts->nops);
    if (e != cudaSuccess) { rbffd_operator_destroy(op); RBFFD_FAIL(ctx, RBFFD_ERR_CUDA, "cudaMalloc of the operator failed: %s", cudaGetErrorString(e)); }
    int rc = rbffd_weights_impl(ctx, opts, X, N, Y, M, stencils.p, N, center.p, op->colind, op->vals);
    if (rc != RBFFD_OK) { rbffd_operator_destroy(op); return rc; }
    *out = op;
    return RBFFD_OK;
}

int rbffd_operator_generate_host(rbffd_context* ctx, const rbffd_options* opts, const double* X, int64_t N, const double* Y, int64_t M,
                                 const int32_t* xgroup, rbffd_operator** out) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (!out) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "operator_generate_host: NULL output handle");
    *out = nullptr;
    RBFFD_TRY(rbffd_validate_options(ctx, opts));
    if (!X || N < 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "operator_generate_host: X is empty");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int dim = opts->dim;
    DevBuf<double> dX, dY;
    DevBuf<int32_t> dG;
    CUDA_TRY(ctx, dX.alloc((size_t)N * dim, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(dX.p, X, sizeof(double) * N * dim, cudaMemcpyHostToDevice, st));
    if (Y && Y != X) {
        if (M < 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "operator_generate_host: Y is empty");
        CUDA_TRY(ctx, dY.alloc((size_t)M * dim, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(dY.p, Y, sizeof(double) * M * dim, cudaMemcpyHostToDevice, st));
    }
    if (xgroup) {
        CUDA_TRY(ctx, dG.alloc(N, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(dG.p, xgroup, sizeof(int32_t) * N, cudaMemcpyHostToDevice, st));
    }
    const bool two = Y && Y != X;
    int rc = rbffd_operator_generate(ctx, opts, dX.p, N, two ? dY.p : nullptr, two ? M : N, xgroup ? dG.p : nullptr, out);
    cudaStreamSynchronize(st);
    return rc;
}

int rbffd_device_malloc(rbffd_context* ctx, int64_t bytes, void** ptr) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (!ptr || bytes < 0) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "device_malloc: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMalloc(ptr, (size_t)std::max<int64_t>(bytes, 1)));
    return RBFFD_OK;
}

int rbffd_device_free(rbffd_context* ctx, void* ptr) {
    if (!ctx) return RBFFD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ptr) CUDA_TRY(ctx, cudaFree(ptr));
    return RBFFD_OK;
}

int rbffd_device_upload(rbffd_context* ctx, void* dst, const void* src, int64_t bytes) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "device_upload: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (bytes > 0) CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RBFFD_OK;
}

int rbffd_device_download(rbffd_context* ctx, void* dst, const void* src, int64_t bytes) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "device_download: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (bytes > 0) CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RBFFD_OK;
}

int rbffd_operator_from_host(rbffd_context* ctx, int64_t M, int64_t N, int32_t n, int32_t nmat, const int64_t* colind,
                             int32_t index_base, const double* vals, rbffd_operator** out) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (!out || !colind || !vals) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "operator_from_host: NULL pointer");
    *out = nullptr;
    if (M < 0 || N < 1 || n < 1 || nmat < 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "operator_from_host: bad sizes");
    if (N > 0x7fffffff) RBFFD_FAIL(ctx, RBFFD_ERR_UNSUPPORTED, "column count exceeds the int32 device index range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    rbffd_operator* op = new (std::nothrow) rbffd_operator();
    if (!op) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "out of host memory");
    op->ctx = ctx; op->M = M; op->N = N; op->n = n; op->nmat = nmat;
    const size_t nnz = (size_t)std::max<int64_t>(M * n, 1);
    cudaError_t e = cudaMalloc((void**)&op->colind, sizeof(int32_t) * nnz);
    if (e == cudaSuccess) e = cudaMalloc((void**)&op->vals, sizeof(double) * nnz * nmat);
    if (e != cudaSuccess) { rbffd_operator_destroy(op); RBFFD_FAIL(ctx, RBFFD_ERR_CUDA, "cudaMalloc of the operator failed: %s", cudaGetErrorString(e)); }
    DevBuf<int64_t> wide;
    DevBuf<int> bad;
    CUDA_TRY(ctx, wide.alloc(nnz, st));
    CUDA_TRY(ctx, bad.alloc(1, st));
    CUDA_TRY(ctx, cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    CUDA_TRY(ctx, cudaMemcpyAsync(wide.p, colind, sizeof(int64_t) * M * n, cudaMemcpyHostToDevice, st));
    if (M * n > 0) i64_to_i32_kernel<<<ceil_div_i64(M * n, 256), 256, 0, st>>>(wide.p, M * n, index_base, op->colind, N, bad.p);
    CUDA_TRY(ctx, cudaMemcpyAsync(op->vals, vals, sizeof(double) * M * n * nmat, cudaMemcpyHostToDevice, st));
    int h_bad = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (h_bad) { rbffd_operator_destroy(op); RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "operator_from_host: column index out of range"); }
    *out = op;
    return RBFFD_OK;
}

int rbffd_operator_destroy(rbffd_operator* op) {
    if (!op) return RBFFD_OK;
    if (op->ctx) { cudaSetDevice(op->ctx->device); cudaStreamSynchronize(op->ctx->stream); }
    if (!op->borrowed) { cudaFree(op->colind); cudaFree(op->vals); }
    cudaFree(op->t_ptr); cudaFree(op->t_src); cudaFree(op->t_row); cudaFree(op->work); cudaFree(op->work2);
    for (double* p : op->t_vals) cudaFree(p);
    delete op;
    return RBFFD_OK;
}

int rbffd_operator_info(const rbffd_operator* op, int64_t* M, int64_t* N, int32_t* n, int32_t* nmat) {
    if (!op) return RBFFD_ERR_INVALID;
    if (M) *M = op->M;
    if (N) *N = op->N;
    if (n) *n = op->n;
    if (nmat) *nmat = op->nmat;
    return RBFFD_OK;
}

int rbffd_operator_pointers(const rbffd_operator* op, int32_t which, const int32_t** colind, const double** vals) {
    if (!op || which < 0 || which >= op->nmat) return RBFFD_ERR_INVALID;
    if (colind) *colind = op->colind;
    if (vals) *vals = op->vals + (size_t)op->M * op->n * which;
    return RBFFD_OK;
}

int rbffd_operator_to_host(rbffd_operator* op, int32_t index_base, int64_t* colind_out, double* vals_out) {
    if (!op) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = op->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t nnz = op->M * op->n;
    if (vals_out && nnz > 0)
        CUDA_TRY(ctx, cudaMemcpyAsync(vals_out, op->vals, sizeof(double) * nnz * op->nmat, cudaMemcpyDeviceToHost, ctx->stream));
    RBFFD_TRY(copy_indices_to_host(ctx, op->colind, nnz, index_base, colind_out));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RBFFD_OK;
}

int rbffd_spmv_device(rbffd_operator* op, int32_t which, double alpha, const double* x, double beta, double* y) {
    if (!op) return RBFFD_ERR_INVALID;
    CUDA_TRY(op->ctx, cudaSetDevice(op->ctx->device));
    return rbffd_spmv_multi_impl(op, 1, &which, &alpha, x, beta, y);
}

int rbffd_spmv_multi_device(rbffd_operator* op, int32_t nterms, const int32_t* which, const double* coef, const double* x, double* y) {
    if (!op) return RBFFD_ERR_INVALID;
    CUDA_TRY(op->ctx, cudaSetDevice(op->ctx->device));
    return rbffd_spmv_multi_impl(op, nterms, which, coef, x, 0.0, y);
}

namespace {
struct CombineArgs { const double* v[8]; double c[8]; int nt; };
__global__ void __launch_bounds__(256) combine_values_kernel(CombineArgs a, int64_t count, double* __restrict__ out, bool vec) {
    const int64_t e = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 2;
    if (vec && e + 1 < count) {                         // every base pointer is 16-byte aligned: 16-byte streaming loads
        double2 acc = make_double2(0.0, 0.0);
        for (int k = 0; k < a.nt; ++k) {
            const double2 v = __ldcs(reinterpret_cast<const double2*>(a.v[k] + e));
            acc.x = fma(a.c[k], v.x, acc.x);
            acc.y = fma(a.c[k], v.y, acc.y);
        }
        *reinterpret_cast<double2*>(out + e) = acc;
    } else {
        for (int64_t f = e; f < count && f < e + 2; ++f) {
            double acc = 0.0;
            for (int k = 0; k < a.nt; ++k) acc = fma(a.c[k], a.v[k][f], acc);
            out[f] = acc;
        }
    }
}
}  // namespace

int rbffd_operator_combine_device(rbffd_operator* op, int32_t nterms, const int32_t* which, const double* coef, double* vals_out) {
    if (!op) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = op->ctx;
    if (nterms < 1 || nterms > 8 || !which || !coef || !vals_out) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "operator_combine: 1..8 terms and non-NULL arrays");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CombineArgs a;
    a.nt = nterms;
    const int64_t count = op->M * (int64_t)op->n;
    for (int k = 0; k < nterms; ++k) {
        if (which[k] < 0 || which[k] >= op->nmat) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "operator_combine: matrix index %d out of range", which[k]);
        a.v[k] = op->vals + (size_t)which[k] * count;
        a.c[k] = coef[k];
    }
    if (count == 0) return RBFFD_OK;
    bool vec = (reinterpret_cast<uintptr_t>(vals_out) & 15) == 0;
    for (int k = 0; k < nterms; ++k) vec = vec && (reinterpret_cast<uintptr_t>(a.v[k]) & 15) == 0;
    combine_values_kernel<<<ceil_div_i64((count + 1) / 2, 256), 256, 0, ctx->stream>>>(a, count, vals_out, vec);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

int rbffd_spmv_t_device(rbffd_operator* op, int32_t which, double alpha, const double* v, double beta, double* y) {
    if (!op) return RBFFD_ERR_INVALID;
    CUDA_TRY(op->ctx, cudaSetDevice(op->ctx->device));
    return rbffd_spmv_t_impl(op, which, alpha, v, beta, y);
}

static int host_apply(rbffd_operator* op, int64_t nin, int64_t nout, const double* hin, double beta, double* hout,
                      int (*fn)(rbffd_operator*, const double*, double*, void*), void* arg) {
    rbffd_context* ctx = op->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf<double> din, dout;
    CUDA_TRY(ctx, din.alloc(nin, st));
    CUDA_TRY(ctx, dout.alloc(nout, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(din.p, hin, sizeof(double) * nin, cudaMemcpyHostToDevice, st));
    if (beta != 0.0) CUDA_TRY(ctx, cudaMemcpyAsync(dout.p, hout, sizeof(double) * nout, cudaMemcpyHostToDevice, st));
    RBFFD_TRY(fn(op, din.p, dout.p, arg));
    CUDA_TRY(ctx, cudaMemcpyAsync(hout, dout.p, sizeof(double) * nout, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return RBFFD_OK;
}

struct SpmvArg { int which; double alpha, beta; };

int rbffd_spmv_host(rbffd_operator* op, int32_t which, double alpha, const double* x, double beta, double* y) {
    if (!op) return RBFFD_ERR_INVALID;
    if (!x || !y) RBFFD_FAIL(op->ctx, RBFFD_ERR_INVALID, "spmv: NULL pointer");
    SpmvArg a{which, alpha, beta};
    return host_apply(op, op->N, op->M, x, beta, y, [](rbffd_operator* o, const double* in, double* out, void* p) {
        SpmvArg* s = (SpmvArg*)p;
        return rbffd_spmv_multi_impl(o, 1, &s->which, &s->alpha, in, s->beta, out);
    }, &a);
}

int rbffd_spmv_t_host(rbffd_operator* op, int32_t which, double alpha, const double* v, double beta, double* y) {
    if (!op) return RBFFD_ERR_INVALID;
    if (!v || !y) RBFFD_FAIL(op->ctx, RBFFD_ERR_INVALID, "spmv_t: NULL pointer");
    SpmvArg a{which, alpha, beta};
    return host_apply(op, op->M, op->N, v, beta, y, [](rbffd_operator* o, const double* in, double* out, void* p) {
        SpmvArg* s = (SpmvArg*)p;
        return rbffd_spmv_t_impl(o, s->which, s->alpha, in, s->beta, out);
    }, &a);
}

namespace {
// max |E - I| over the entries of one value plane (rows = nodes): decides whether E' * w may be replaced by w
__global__ void identity_defect_kernel(const int32_t* __restrict__ colind, const double* __restrict__ vals, int64_t M, int n, unsigned long long* out) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double d = 0.0;
    if (e < M * n) {
        const int64_t row = e / n;
        const double v = vals[e] - (colind[e] == row ? 1.0 : 0.0);
        d = v == v ? fabs(v) : __longlong_as_double(0x7ff0000000000000ll);
    }
    for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0 && d > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(d));     // non-negative doubles order like integers
}

// 1: matrix iE of a square operator equals the identity to 1e-8: at collocated rows E = I exactly, what the weight solve returns
// deviates from it by its rounding noise only (eps * cond(A_i): 1e-11 ... 1e-9 at the configurations of BASELINE.json)
int e_is_identity(rbffd_operator* op, int iE, bool* yes) {
    rbffd_context* ctx = op->ctx;
    *yes = false;
    if (op->M != op->N || iE < 0 || iE >= op->nmat) return RBFFD_OK;
    if (op->e_identity_which == iE && op->e_identity >= 0) { *yes = op->e_identity == 1; return RBFFD_OK; }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CUDA_TRY(ctx, cudaStreamIsCapturing(ctx->stream, &cap));
    if (cap != cudaStreamCaptureStatusNone) return RBFFD_OK;            // cannot synchronise inside a capture: take the general path
    DevBuf<unsigned long long> d;
    CUDA_TRY(ctx, d.alloc(1, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(d.p, 0, sizeof(unsigned long long), ctx->stream));
    const int64_t nnz = op->M * (int64_t)op->n;
    identity_defect_kernel<<<ceil_div_i64(std::max<int64_t>(nnz, 1), 256), 256, 0, ctx->stream>>>(op->colind, op->vals + (size_t)nnz * iE, op->M, op->n, d.p);
    KLAUNCH(ctx);
    unsigned long long h = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&h, d.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    double defect;
    memcpy(&defect, &h, sizeof(defect));
    op->e_identity_which = iE;
    op->e_identity = defect <= 1e-8 ? 1 : 0;
    *yes = op->e_identity == 1;
    return RBFFD_OK;
}

// terms of the interior line of cons_sys with E = I: alpha*(Dxx + Dyy) - ux*Dx - uy*Dy - gamma*(Dxk + Dyk); zero coefficients dropped
int advdiff_terms(const rbffd_advdiff_params* prm, int32_t* which, double* coef) {
    int nt = 0;
    if (prm->alpha != 0.0) { which[nt] = prm->iDxx; coef[nt++] = prm->alpha; which[nt] = prm->iDyy; coef[nt++] = prm->alpha; }
    if (prm->ux != 0.0) { which[nt] = prm->iDx; coef[nt++] = -prm->ux; }
    if (prm->uy != 0.0) { which[nt] = prm->iDy; coef[nt++] = -prm->uy; }
    if (prm->iDxk >= 0 && prm->iDyk >= 0 && prm->gamma != 0.0) { which[nt] = prm->iDxk; coef[nt++] = -prm->gamma; which[nt] = prm->iDyk; coef[nt++] = -prm->gamma; }
    return nt;
}

// element-wise, so out may alias u or x
__global__ void stage_update_kernel(int64_t N, double a, const double* u, double b, const double* x, double dt, const double* du, double* out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) out[i] = a * u[i] + b * (x[i] + dt * du[i]);
}
}  // namespace

int rbffd_rhs_advdiff_device(rbffd_operator* op, const rbffd_advdiff_params* prm, const double* u, double* du) {
    if (!op) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = op->ctx;
    if (!prm || !u || !du) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "rhs_advdiff: NULL pointer");
    if (op->M != op->N) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "rhs_advdiff: the semidiscretisation needs square operators (M == N)");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (prm->flags & RBFFD_ADVDIFF_COLLOCATED) {
        // collocated rows: E = I, so the whole line is ONE pass over the shared pattern with up to six value planes
        bool ident = false;
        RBFFD_TRY(e_is_identity(op, prm->iE, &ident));
        if (ident) {
            int32_t which[6];
            double coef[6];
            const int nt = advdiff_terms(prm, which, coef);
            if (nt == 0) { CUDA_TRY(ctx, cudaMemsetAsync(du, 0, sizeof(double) * op->N, ctx->stream)); return RBFFD_OK; }
            return rbffd_spmv_multi_impl(op, nt, which, coef, u, 0.0, du);
        }
    }
    if (!op->work) CUDA_TRY(ctx, cudaMalloc((void**)&op->work, sizeof(double) * (size_t)std::max<int64_t>(op->M, 1)));
    // w = alpha*Dxx*u + alpha*Dyy*u - ux*Dx*u - uy*Dy*u  (zero coefficients are dropped: no traffic for them)
    int32_t which[4];
    double coef[4];
    int nt = 0;
    if (prm->alpha != 0.0) { which[nt] = prm->iDxx; coef[nt++] = prm->alpha; which[nt] = prm->iDyy; coef[nt++] = prm->alpha; }
    if (prm->ux != 0.0) { which[nt] = prm->iDx; coef[nt++] = -prm->ux; }
    if (prm->uy != 0.0) { which[nt] = prm->iDy; coef[nt++] = -prm->uy; }
    if (nt == 0) CUDA_TRY(ctx, cudaMemsetAsync(op->work, 0, sizeof(double) * op->M, ctx->stream));
    else RBFFD_TRY(rbffd_spmv_multi_impl(op, nt, which, coef, u, 0.0, op->work));
    // du = E' * w
    RBFFD_TRY(rbffd_spmv_t_impl(op, prm->iE, 1.0, op->work, 0.0, du));
    // du -= gamma * (Dxk + Dyk) * u
    if (prm->iDxk >= 0 && prm->iDyk >= 0 && prm->gamma != 0.0) {
        int32_t wk[2] = {prm->iDxk, prm->iDyk};
        double ck[2] = {-prm->gamma, -prm->gamma};
        RBFFD_TRY(rbffd_spmv_multi_impl(op, 2, wk, ck, u, 1.0, du));
    }
    return RBFFD_OK;
}

int rbffd_spmv_stage_device(rbffd_operator* op, int32_t nterms, const int32_t* which, const double* coef, const double* x,
                            double a, const double* u, double b, double dt, double* out) {
    if (!op) return RBFFD_ERR_INVALID;
    if (!which || !coef || !x || !u || !out) RBFFD_FAIL(op->ctx, RBFFD_ERR_INVALID, "spmv_stage: NULL pointer");
    CUDA_TRY(op->ctx, cudaSetDevice(op->ctx->device));
    return rbffd_spmv_stage_impl(op, nterms, which, coef, x, a, u, b, dt, out);
}

int rbffd_rhs_advdiff_stage_device(rbffd_operator* op, const rbffd_advdiff_params* prm, const double* x, double a, const double* u,
                                   double b, double dt, double* out) {
    if (!op) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = op->ctx;
    if (!prm || !x || !u || !out) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "rhs_advdiff_stage: NULL pointer");
    if (op->M != op->N) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "rhs_advdiff_stage: the semidiscretisation needs square operators (M == N)");
    if (out == x || out == u) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "rhs_advdiff_stage: out must not alias x or u");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (prm->flags & RBFFD_ADVDIFF_COLLOCATED) {
        bool ident = false;
        RBFFD_TRY(e_is_identity(op, prm->iE, &ident));
        int32_t which[6];
        double coef[6];
        const int nt = advdiff_terms(prm, which, coef);
        if (ident && nt > 0) return rbffd_spmv_stage_impl(op, nt, which, coef, x, a, u, b, dt, out);
    }
    if (!op->work2) CUDA_TRY(ctx, cudaMalloc((void**)&op->work2, sizeof(double) * (size_t)std::max<int64_t>(op->N, 1)));
    RBFFD_TRY(rbffd_rhs_advdiff_device(op, prm, x, op->work2));
    stage_update_kernel<<<ceil_div_i64(std::max<int64_t>(op->N, 1), 256), 256, 0, ctx->stream>>>(op->N, a, u, b, x, dt, op->work2, out);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

int rbffd_stage_update_device(rbffd_context* ctx, int64_t N, double a, const double* u, double b, const double* x, double dt,
                              const double* du, double* out) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (N < 0 || !u || !x || !du || !out) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "stage_update: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (N == 0) return RBFFD_OK;
    stage_update_kernel<<<ceil_div_i64(N, 256), 256, 0, ctx->stream>>>(N, a, u, b, x, dt, du, out);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

int rbffd_rhs_advdiff_host(rbffd_operator* op, const rbffd_advdiff_params* prm, const double* u, double* du) {
    if (!op) return RBFFD_ERR_INVALID;
    if (!prm || !u || !du) RBFFD_FAIL(op->ctx, RBFFD_ERR_INVALID, "rhs_advdiff: NULL pointer");
    return host_apply(op, op->N, op->N, u, 0.0, du, [](rbffd_operator* o, const double* in, double* out, void* p) {
        return rbffd_rhs_advdiff_device(o, (const rbffd_advdiff_params*)p, in, out);
    }, (void*)prm);
}

int rbffd_gather_device(rbffd_context* ctx, const double* src, const int32_t* index, int64_t count, double* dst) {
    if (!ctx) return RBFFD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return rbffd_gather_impl(ctx, src, index, count, dst);
}

int rbffd_scatter_add_device(rbffd_context* ctx, const double* src, const int32_t* index, int64_t count, double* dst) {
    if (!ctx) return RBFFD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return rbffd_scatter_add_impl(ctx, src, index, count, dst);
}

int rbffd_jittered_lattice_device(rbffd_context* ctx, int32_t dim, int64_t g, uint64_t seed, int64_t first, int64_t count, double* X_out) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (dim < 2 || dim > 3 || g < 1 || first < 0 || count < 0 || !X_out) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "jittered_lattice: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return RBFFD_OK;
    lattice_kernel<<<ceil_div_i64(count, 256), 256, 0, ctx->stream>>>(dim, g, seed, first, count, X_out);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

}  // extern "C"
