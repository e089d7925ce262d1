// common.cuh -- context, error plumbing and small device helpers shared by the librbffd.so kernels.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <algorithm>
#include <chrono>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/rbffd.h"

struct rbffd_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t owned_stream = nullptr;  // created by rbffd_create; `stream` points at it until rbffd_set_stream borrows another
    cudaStream_t copy_stream = nullptr;   // D2H of finished row chunks overlaps the weight kernel of the next chunk
    cudaEvent_t chunk_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;
    double timings[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int sm_count = 148;
    int max_smem_optin = 0;
    int* hflags = nullptr;       // 16 ints of mapped pinned host memory: status words are published with plain stores from a
    int* hflags_dev = nullptr;   // tiny kernel, never by a D2H memcpy (that would queue behind bulk D2H traffic on the copy engine)
    bool trusted_stencils = false;   // set by internal callers whose stencils come from our own search (skips range checks)
    bool collocated_rows = false;    // set by internal callers whose row i is evaluated AT the centre of stencil i (eta == 0 exactly)
    long long launches = 0;      // hand-written kernels launched through this context (rbffd_launch_count)
    // deferred status checks (host entry point): when set, the weight path writes its status words of the current row
    // chunk to deferred_flags[8 * deferred_slot ..] ([0] singular node + 1, [4] null-space kernel refused) and never
    // synchronises; the caller inspects all chunks once at the end and redoes the rare chunk that needs the fallback
    int* deferred_flags = nullptr;
    int deferred_slot = 0;
    int32_t* stage_i32 = nullptr;    // pinned staging for the int32 pattern (host entry point widens it to the caller's int64)
    size_t stage_i32_count = 0;
};

struct rbffd_operator {
    rbffd_context* ctx = nullptr;
    int64_t M = 0, N = 0;
    int32_t n = 0, nmat = 0;
    int32_t* colind = nullptr;   // [M*n] device, 0-based, shared by all matrices
    double* vals = nullptr;      // [nmat][M*n] device
    // lazily built transpose pattern (for E' * v): CSC of the M x N pattern
    int32_t* t_ptr = nullptr;    // [N+1]
    int32_t* t_src = nullptr;    // [M*n] entry id (k*n+j) sorted by column
    int32_t* t_row = nullptr;    // [M*n] row of that entry (t_src / n)
    std::vector<double*> t_vals; // [nmat] values in column order, built on the first E'*v of a matrix (owned operators only:
                                 // a borrowed operator's values may be rewritten by the caller between applications)
    double* work = nullptr;      // [M] scratch
    double* work2 = nullptr;     // [N] scratch of the unfused stage path
    int e_identity = -1, e_identity_which = -1;   // cached answer of "is matrix e_identity_which the identity (1e-12)?"
    bool borrowed = false;       // colind/vals belong to the caller (rbffd_operator_from_device)
};

#define RBFFD_FAIL(ctx, code, ...)                                   \
    do {                                                             \
        char _b[512];                                                \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                       \
        if (ctx) (ctx)->err = _b;                                    \
        return (code);                                               \
    } while (0)

#define CUDA_TRY(ctx, expr)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            RBFFD_FAIL(ctx, RBFFD_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e),   \
                       __FILE__, __LINE__, #expr);                                                   \
    } while (0)

#define RBFFD_TRY(expr)                  \
    do {                                 \
        int _rc = (expr);                \
        if (_rc != RBFFD_OK) return _rc; \
    } while (0)

// Host-side free list in front of the stream-ordered pool.  cudaMallocAsync from a warm pool normally returns in microseconds,
// but on the boxes of this pool single calls were seen to block for 50-400 ms with the pool neither growing nor short of free
// memory (profiles/r02aq_alloc_stalls.txt) -- contention inside the driver that a library cannot fix but can stay away from:
// temporaries go back to this list instead of cudaFreeAsync and are handed out again to requests ON THE SAME STREAM (stream
// order makes that exactly as safe as free + malloc on that stream), so a repeated call sequence never enters the allocator.
// RBFFD_SCRATCH_CACHE_MB bounds the bytes held (default 16384: the scratch of the two-stage weight path alone is up to 12 GiB;
// 0 disables the list).
struct DevBlockCache {
    struct Blk { void* p; size_t bytes; cudaStream_t s; };
    std::mutex mu;
    std::vector<Blk> blks;
    size_t held = 0, cap = 0;
    DevBlockCache() {
        const char* e = getenv("RBFFD_SCRATCH_CACHE_MB");
        cap = (size_t)(e ? std::max(0ll, atoll(e)) : 16384ll) << 20;
    }
    void* take(size_t bytes, cudaStream_t s, size_t* got) {
        std::lock_guard<std::mutex> lk(mu);
        int best = -1;
        const size_t hi = bytes + bytes / 4 + (1u << 20);
        for (int i = 0; i < (int)blks.size(); ++i)
            if (blks[i].s == s && blks[i].bytes >= bytes && blks[i].bytes <= hi && (best < 0 || blks[i].bytes < blks[best].bytes)) best = i;
        if (best < 0) return nullptr;
        void* q = blks[best].p;
        *got = blks[best].bytes;
        held -= blks[best].bytes;
        blks[best] = blks.back();
        blks.pop_back();
        return q;
    }
    bool give(void* q, size_t bytes, cudaStream_t s) {
        std::lock_guard<std::mutex> lk(mu);
        if (held + bytes > cap) return false;
        blks.push_back({q, bytes, s});
        held += bytes;
        return true;
    }
    // every parked block goes back to the pool (rbffd_destroy: the streams they are parked under may be about to disappear).
    // cudaFree, not cudaFreeAsync: it needs no live stream and may not be issued while a stream capture is active anyway;
    // blocks parked under a borrowed stream that the caller switched away from (rbffd_set_stream) simply wait here until
    // that stream comes back or a context is destroyed.
    void clear() {
        std::lock_guard<std::mutex> lk(mu);
        for (const Blk& b : blks) cudaFree(b.p);
        blks.clear();
        held = 0;
    }
};
inline DevBlockCache& dev_block_cache() {
    static DevBlockCache* c = new DevBlockCache();       // never destroyed: no CUDA calls during static destruction
    return *c;
}

// stream-ordered temporary buffer (returned to the free list / the pool on scope exit, stream ordered)
template <typename T>
struct DevBuf {
    T* p = nullptr;
    cudaStream_t s = nullptr;
    size_t bytes = 0;            // size of the block behind p (>= the request when it came from the free list)
    bool parkable = false;       // allocated outside a stream capture
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    cudaError_t alloc(size_t count, cudaStream_t stream) {
        s = stream;
        if (count == 0) count = 1;
        const size_t want = count * sizeof(T);
        // inside a stream capture the allocation must become a node of the graph (a parked block would be baked into the graph
        // and handed to somebody else afterwards): no free list there
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        parkable = cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone;
        if (parkable) {
            if (void* q = dev_block_cache().take(want, stream, &bytes)) { p = reinterpret_cast<T*>(q); return cudaSuccess; }
        }
        bytes = want;
        // RBFFD_TRACE_ALLOC=1: report pool allocations that take the host more than 1 ms
        static const bool trace = [] { const char* e = getenv("RBFFD_TRACE_ALLOC"); return e && atoi(e) != 0; }();
        if (!trace) return cudaMallocAsync(reinterpret_cast<void**>(&p), want, stream);
        const auto t0 = std::chrono::steady_clock::now();
        const cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&p), want, stream);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 1.0) {
            uint64_t reserved = 0, used = 0;
            int dev = 0;
            cudaMemPool_t pool;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
            }
            fprintf(stderr, "[rbffd alloc] %.1f MB took %.2f ms on the host (pool after: reserved %.1f MB, used %.1f MB)\n",
                    want * 1e-6, ms, reserved * 1e-6, used * 1e-6);
        }
        return e;
    }
    T* release() { T* q = p; p = nullptr; return q; }
    void reset() { if (p && !(parkable && dev_block_cache().give(p, bytes, s))) cudaFreeAsync(p, s); p = nullptr; }
    ~DevBuf() { if (p && !(parkable && dev_block_cache().give(p, bytes, s))) cudaFreeAsync(p, s); }
};

static inline int ceil_div_i64(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
static __global__ void rbffd_publish_kernel(const int* __restrict__ src, int n, volatile int* dst) {
    if ((int)threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}
static __global__ void rbffd_set_flags_kernel(int* dst, int v0, int v1, int v2, int v3) {
    if (threadIdx.x == 0) { dst[0] = v0; dst[1] = v1; dst[2] = v2; dst[3] = v3; }
}
// device status words -> host, through mapped pinned memory (n <= 16); synchronises the context's stream
static inline cudaError_t rbffd_fetch_flags(rbffd_context* ctx, const int* dev, int n, int* out) {
    rbffd_publish_kernel<<<1, 32, 0, ctx->stream>>>(dev, n, ctx->hflags_dev);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) for (int i = 0; i < n; ++i) out[i] = ((volatile int*)ctx->hflags)[i];
    return e;
}
#endif

// ---- internal entry points shared between translation units ----------------------------------------
struct KnnGridPlan;   // knn.cu

int rbffd_knn_impl(rbffd_context* ctx, const double* X, int64_t N, int dim, const double* Q, int64_t NQ, int k,
                   const int32_t* xgroup, const int32_t* qgroup, bool q_is_x, int32_t* idx_out, double* d2_out);
// kNN of X among X (k = n) plus nearest X of every Y row (k = 1) sharing one binning pass.
int rbffd_stencils_impl(rbffd_context* ctx, const double* X, int64_t N, int dim, const double* Y, int64_t M, int n,
                        const int32_t* xgroup, int32_t* stencils, double* d2_x, int32_t* center, double* d2_y);
// stencils [NS][n] index into X (NX nodes); row k of the operators uses stencil center[k] (NULL: k, needs M == NS)
int rbffd_weights_impl(rbffd_context* ctx, const rbffd_options* opts, const double* X, int64_t NX,
                       const double* Y, int64_t M, const int32_t* stencils, int64_t NS, const int32_t* center,
                       int32_t* colind_out, double* vals_out);
#define KLAUNCH(ctx) ((ctx)->launches++)
int rbffd_validate_options(rbffd_context* ctx, const rbffd_options* o);
