// halo.cu -- NVLink peer-memory halo exchange for the sharded SpMV, without NCCL on the data path.
//
// Each rank keeps its field u = [halo_lo | owned | halo_hi] plus 4 flag words in a cudaMalloc'd buffer exported
// with CUDA IPC; the neighbours map it (cudaIpcOpenMemHandle) and STORE their boundary values straight into its halo
// regions over NVLink, then publish an epoch flag (release, system scope).  The receiver's stream only waits for the
// flags right before its boundary rows; interior rows run while the stores are in flight.  An ack flag written back
// after the boundary rows keeps a fast neighbour from overwriting a halo that is still being read.
//   push(epoch):  wait ack >= epoch-1 from each neighbour ; copy ; __threadfence_system ; ready = epoch   (one CTA per side)
//   wait(epoch):  spin until ready >= epoch from each neighbour
//   ack(epoch):   ack = epoch in each neighbour's flag block
// The reference has no distributed code; this replaces what a halo exchange of `u` before `D*u` would be.
#include "common.cuh"

namespace {

// flag block layout (uint32): [0] ready from lower neighbour, [1] ready from upper, [2] ack from lower, [3] ack from upper
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct PushSide {
    const double* src;      // my boundary values
    double* dst;            // neighbour's halo region (peer memory)
    int64_t count;
    const unsigned* my_ack; // written by the neighbour when it has consumed the previous epoch
    unsigned* peer_ready;   // neighbour's ready flag for my side
};

__global__ void halo_push_kernel(PushSide lo, PushSide hi, unsigned epoch) {
    const PushSide s = blockIdx.y == 0 ? lo : hi;
    if (s.count == 0) return;
    if (threadIdx.x == 0) {      // the neighbour must have consumed the previous epoch before its halo is overwritten
        while ((int)(ld_acquire_sys(s.my_ack) - (epoch - 1u)) < 0) { }
    }
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < s.count; i += (int64_t)gridDim.x * blockDim.x) s.dst[i] = s.src[i];
}

// one thread per side: after the copy kernel completed (stream order), publish the epoch
__global__ void halo_publish_kernel(unsigned* ready_lo, unsigned* ready_hi, unsigned epoch) {
    __threadfence_system();
    if (threadIdx.x == 0 && ready_lo) st_release_sys(ready_lo, epoch);
    if (threadIdx.x == 1 && ready_hi) st_release_sys(ready_hi, epoch);
}

__global__ void halo_wait_kernel(const unsigned* ready_lo, const unsigned* ready_hi, unsigned epoch) {
    if (threadIdx.x == 0 && ready_lo) while ((int)(ld_acquire_sys(ready_lo) - epoch) < 0) { }
    if (threadIdx.x == 1 && ready_hi) while ((int)(ld_acquire_sys(ready_hi) - epoch) < 0) { }
}

__global__ void halo_ack_kernel(unsigned* ack_lo, unsigned* ack_hi, unsigned epoch) {
    if (threadIdx.x == 0 && ack_lo) st_release_sys(ack_lo, epoch);
    if (threadIdx.x == 1 && ack_hi) st_release_sys(ack_hi, epoch);
}

}  // namespace

extern "C" {

int rbffd_ipc_alloc(rbffd_context* ctx, int64_t bytes, void** ptr, unsigned char* handle64) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (!ptr || !handle64 || bytes <= 0) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "ipc_alloc: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CUDA_TRY(ctx, cudaMalloc(ptr, (size_t)bytes));
    CUDA_TRY(ctx, cudaMemset(*ptr, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, *ptr));
    memcpy(handle64, &h, 64);
    return RBFFD_OK;
}

int rbffd_ipc_free(rbffd_context* ctx, void* ptr) {
    if (!ctx) return RBFFD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaFree(ptr));
    return RBFFD_OK;
}

int rbffd_ipc_open(rbffd_context* ctx, const unsigned char* handle64, void** peer_ptr) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (!handle64 || !peer_ptr) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "ipc_open: NULL pointer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CUDA_TRY(ctx, cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RBFFD_OK;
}

int rbffd_ipc_close(rbffd_context* ctx, void* peer_ptr) {
    if (!ctx) return RBFFD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaIpcCloseMemHandle(peer_ptr));
    return RBFFD_OK;
}

int rbffd_halo_push_device(rbffd_context* ctx, const rbffd_halo* h, uint32_t epoch) {
    if (!ctx || !h) return RBFFD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    PushSide lo{}, hi{};
    unsigned* my_flags = reinterpret_cast<unsigned*>(h->flags);
    if (h->peer_lo_u && h->count_to_lo > 0) {
        lo.src = h->u + h->n_lo;
        lo.dst = h->peer_lo_u + h->peer_lo_offset;
        lo.count = h->count_to_lo;
        lo.my_ack = my_flags + 2;
        lo.peer_ready = reinterpret_cast<unsigned*>(h->peer_lo_flags) + 1;   // I am the lower neighbour's UPPER side
    }
    if (h->peer_hi_u && h->count_to_hi > 0) {
        hi.src = h->u + h->n_lo + h->n_owned - h->count_to_hi;
        hi.dst = h->peer_hi_u + h->peer_hi_offset;
        hi.count = h->count_to_hi;
        hi.my_ack = my_flags + 3;
        hi.peer_ready = reinterpret_cast<unsigned*>(h->peer_hi_flags) + 0;   // I am the upper neighbour's LOWER side
    }
    if (lo.count == 0 && hi.count == 0) return RBFFD_OK;
    const int64_t mx = std::max(lo.count, hi.count);
    dim3 grid((unsigned)std::min<int64_t>((mx + 255) / 256, 64), 2);
    halo_push_kernel<<<grid, 256, 0, ctx->stream>>>(lo, hi, epoch);
    halo_publish_kernel<<<1, 32, 0, ctx->stream>>>(lo.count ? lo.peer_ready : nullptr, hi.count ? hi.peer_ready : nullptr, epoch);
    KLAUNCH(ctx); KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

int rbffd_halo_wait_device(rbffd_context* ctx, const rbffd_halo* h, uint32_t epoch) {
    if (!ctx || !h) return RBFFD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    unsigned* my_flags = reinterpret_cast<unsigned*>(h->flags);
    const unsigned* rl = (h->peer_lo_u && h->n_lo > 0) ? my_flags + 0 : nullptr;
    const unsigned* rh = (h->peer_hi_u && h->n_hi > 0) ? my_flags + 1 : nullptr;
    if (!rl && !rh) return RBFFD_OK;
    halo_wait_kernel<<<1, 32, 0, ctx->stream>>>(rl, rh, epoch);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

int rbffd_halo_ack_device(rbffd_context* ctx, const rbffd_halo* h, uint32_t epoch) {
    if (!ctx || !h) return RBFFD_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    // tell each neighbour that its data of this epoch has been consumed: the lower neighbour's ack-from-upper word, etc.
    unsigned* al = (h->peer_lo_u && h->n_lo > 0) ? reinterpret_cast<unsigned*>(h->peer_lo_flags) + 3 : nullptr;
    unsigned* ah = (h->peer_hi_u && h->n_hi > 0) ? reinterpret_cast<unsigned*>(h->peer_hi_flags) + 2 : nullptr;
    if (!al && !ah) return RBFFD_OK;
    halo_ack_kernel<<<1, 32, 0, ctx->stream>>>(al, ah, epoch);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

}  // extern "C"
