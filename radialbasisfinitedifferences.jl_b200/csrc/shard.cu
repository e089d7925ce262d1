// shard.cu -- spatial-block sharding of arbitrary node sets over the GPUs of one box, and the sharded operator application
// with its halo exchange FUSED into the SpMV launch (BASELINE.json north_star).  The reference has no
// distributed code (src/domains/domains.jl:7-8 holds commented-out includes only); what this replaces is the single-process
// `D*u` / `E'*v` of examples/adv_diff_test.jl:151-152 when the node set is split over several GPUs.
//
//   plan      rbffd_shard_plan_host         k-way coordinate-quantile blocks (b0 x b1 [x b2]) over arbitrary X
//   build     rbffd_shard_create_{host,device}  owned nodes + candidate halo -> exact kNN of the owned nodes (ties by GLOBAL id)
//                                           -> exactness proof -> halo = stencil closure -> local numbering
//                                           [interior owned | boundary owned | halo (by owner, by global id)], all on the device
//   wire      recv/send id lists, rbffd_shard_finalize / _connect   CUDA-IPC inbox [halo values | flags | reverse inbox]
//   apply     rbffd_shard_spmv_device       ONE launch: push CTAs (peer stores over NVLink + epoch flag), interior-row CTAs,
//                                           boundary-row CTAs (wait for the peers' flags, last one acknowledges); the epoch is a
//                                           device counter, so the launch is CUDA-graph capturable
//             rbffd_shard_spmv_t_device     E'*v with the reverse exchange (halo-column contributions added at their owners)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "spmv_rows.cuh"

int rbffd_spmv_t_impl(rbffd_operator* op, int which, double alpha, const double* v, double beta, double* y);

namespace {

constexpr int MAXP = RBFFD_SHARD_MAX_PEERS;
constexpr int PUSH_CHUNK = 4096;            // values one push CTA moves
// flag block of an inbox (uint32 words): who writes what
constexpr int F_READY = 0;                  // [src]  forward values of `src` have landed           (written by src)
constexpr int F_ACK = MAXP;                 // [dst]  `dst` has consumed my forward values            (written by dst)
constexpr int F_TREADY = 2 * MAXP;          // [src]  reverse contributions of `src` have landed      (written by src)
constexpr int F_TACK = 3 * MAXP;            // [dst]  `dst` has consumed my reverse contributions     (written by dst)
constexpr int F_WORDS = 4 * MAXP;

struct ShardCtl {                           // plain device memory of the owning rank
    unsigned epoch, tepoch;                 // completed forward / reverse exchanges
    unsigned bdone, all_done;
    unsigned timed_out;                     // a wait for a peer gave up (see spin_until)
    unsigned push_done[MAXP], tpush_done[MAXP];
};

struct ShardDev {                           // passed by value to the kernels
    int me, nsend, nrecv;
    int send_peer[MAXP];
    long long send_off[MAXP + 1];           // prefix of the send counts (slots in rank order)
    int chunk_off[MAXP + 1];                // prefix of the push-chunk counts
    double* send_dst[MAXP];                 // peer's inbox + where my values start in it
    unsigned* send_flags[MAXP];             // peer's flag block
    int recv_peer[MAXP];
    long long recv_off[MAXP + 1];
    int rchunk_off[MAXP + 1];
    double* recv_rdst[MAXP];                // peer's reverse inbox + where my contributions start in it
    unsigned* recv_flags[MAXP];
    const int32_t* send_idx;                // [send_off[nsend]] local owned ids, slot by slot
    double* inbox;                          // [n_halo]
    double* rinbox;                         // [send_off[nsend]]
    unsigned* flags;                        // my flag block
    ShardCtl* ctl;
    long long n_owned, n_int;
};

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Waits for a peer's epoch flag.  A peer that never arrives (crashed process, mismatched call sequence) must not hang the GPU:
// after SPIN_TIMEOUT_NS the wait gives up, records the fact in the control block (rbffd_shard_status reports it as an error)
// and the launch runs to completion with whatever the inbox holds.
constexpr unsigned long long SPIN_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void spin_until(const unsigned* flag, unsigned value, unsigned* timed_out) {
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while ((int)(ld_acquire_sys_u32(flag) - value) < 0) {
        if ((++spins & 1023u) == 0u) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > SPIN_TIMEOUT_NS) { *timed_out = 1u; return; }
        }
    }
}

// ---- sharded SpMV, halo exchange fused ------------------------------------------------------------------------------
// grid = [PB push CTAs | IB interior-row CTAs | BB boundary-row CTAs]; MODE 0 = fused exchange, 1 = no exchange (inbox filled
// by rbffd_shard_unpack_device)
template <int TPR, int NMAT, int VEC, int ITERS, int MODE>
__global__ void __launch_bounds__(256) shard_spmv_kernel(ShardDev sd, int n, const int32_t* __restrict__ colind, SpmvMats m,
                                                         const double* __restrict__ x, SpmvEpilogue epi, double* __restrict__ y, int PB, int IB) {
    __shared__ unsigned s_epoch;
    const int b = blockIdx.x;
    // only the push and the boundary-row CTAs take part in the exchange protocol (a few hundred of the grid's tens of
    // thousands): the interior-row CTAs neither read the epoch nor touch a counter
    const bool proto = MODE == 0 && (b < PB || b >= PB + IB);
    if (proto) {
        if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned*>(&sd.ctl->epoch) + 1u;
        __syncthreads();
    }
    const unsigned ep = proto ? s_epoch : 0u;
    if (b < PB) {
        // ---- push: this CTA moves one chunk of one peer's values ----
        int k = 0;
        while (b >= sd.chunk_off[k + 1]) ++k;
        const long long e0 = sd.send_off[k] + (long long)(b - sd.chunk_off[k]) * PUSH_CHUNK;
        const long long e1 = min(sd.send_off[k + 1], e0 + PUSH_CHUNK);
        if (threadIdx.x == 0) spin_until(sd.flags + F_ACK + sd.send_peer[k], ep - 1u, &sd.ctl->timed_out);     // the peer is done with the previous epoch
        __syncthreads();
        double* dst = sd.send_dst[k] - sd.send_off[k];
        for (long long i = e0 + threadIdx.x; i < e1; i += 256) dst[i] = x[sd.send_idx[i]];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned nch = (unsigned)(sd.chunk_off[k + 1] - sd.chunk_off[k]);
            if (atomicAdd(&sd.ctl->push_done[k], 1u) == nch - 1u) {           // last chunk of this peer: publish the epoch
                sd.ctl->push_done[k] = 0u;
                __threadfence_system();
                st_release_sys_u32(sd.send_flags[k] + F_READY + sd.me, ep);
            }
        }
    } else if (b < PB + IB) {
        const int64_t warp = ((int64_t)(b - PB) * 256 + threadIdx.x) >> 5;
        spmv_rows<TPR, NMAT, VEC, ITERS, false>(warp, 0, sd.n_int, n, colind, m, x, nullptr, 0, epi, y);
    } else {
        if (MODE == 0) {
            if ((int)threadIdx.x < sd.nrecv) spin_until(sd.flags + F_READY + sd.recv_peer[threadIdx.x], ep, &sd.ctl->timed_out);
            __syncthreads();
        }
        const int64_t warp = ((int64_t)(b - PB - IB) * 256 + threadIdx.x) >> 5;
        spmv_rows<TPR, NMAT, VEC, ITERS, true>(warp, sd.n_int, sd.n_owned, n, colind, m, x, sd.inbox, (int)sd.n_owned, epi, y);
        if (MODE == 0) {
            __syncthreads();                       // every halo value this CTA needs has been read
            if (threadIdx.x == 0) {
                const unsigned BB = gridDim.x - (unsigned)(PB + IB);
                if (atomicAdd(&sd.ctl->bdone, 1u) == BB - 1u) {               // last boundary CTA: the peers may overwrite my inbox
                    sd.ctl->bdone = 0u;
                    for (int r = 0; r < sd.nrecv; ++r) st_release_sys_u32(sd.recv_flags[r] + F_ACK + sd.me, ep);
                }
            }
        }
    }
    if (proto) {
        __syncthreads();
        // the last protocol CTA of the launch advances the epoch: every other one has read it by then
        if (threadIdx.x == 0 && atomicAdd(&sd.ctl->all_done, 1u) == gridDim.x - (unsigned)IB - 1u) {
            sd.ctl->all_done = 0u;
            *reinterpret_cast<volatile unsigned*>(&sd.ctl->epoch) = ep;
        }
    }
}

// ---- reverse exchange of the transposed product ----------------------------------------------------------------------
__global__ void __launch_bounds__(256) shard_tpush_kernel(ShardDev sd, const double* __restrict__ text) {
    __shared__ unsigned s_epoch;
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned*>(&sd.ctl->tepoch) + 1u;
    __syncthreads();
    const unsigned ep = s_epoch;
    const int b = blockIdx.x;
    int r = 0;
    while (b >= sd.rchunk_off[r + 1]) ++r;
    const long long e0 = sd.recv_off[r] + (long long)(b - sd.rchunk_off[r]) * PUSH_CHUNK;
    const long long e1 = min(sd.recv_off[r + 1], e0 + PUSH_CHUNK);
    if (threadIdx.x == 0) spin_until(sd.flags + F_TACK + sd.recv_peer[r], ep - 1u, &sd.ctl->timed_out);
    __syncthreads();
    double* dst = sd.recv_rdst[r] - sd.recv_off[r];
    for (long long i = e0 + threadIdx.x; i < e1; i += 256) dst[i] = text[sd.n_owned + i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned nch = (unsigned)(sd.rchunk_off[r + 1] - sd.rchunk_off[r]);
        if (atomicAdd(&sd.ctl->tpush_done[r], 1u) == nch - 1u) {
            sd.ctl->tpush_done[r] = 0u;
            __threadfence_system();
            st_release_sys_u32(sd.recv_flags[r] + F_TREADY + sd.me, ep);
        }
    }
}

__global__ void shard_tinit_kernel(const double* __restrict__ text, long long n_owned, double beta, double* __restrict__ y) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n_owned) y[i] = beta == 0.0 ? text[i] : text[i] + beta * y[i];
}

// contributions of send slot k (the nodes I ship to that peer forward): y[send_idx[i]] += rinbox[i]; one slot per launch so that
// a node needed by several peers receives its contributions in rank order
__global__ void __launch_bounds__(256) shard_tcombine_kernel(ShardDev sd, int k, double* __restrict__ y) {
    __shared__ unsigned s_epoch;
    if (threadIdx.x == 0) {
        s_epoch = *reinterpret_cast<volatile unsigned*>(&sd.ctl->tepoch) + 1u;
        spin_until(sd.flags + F_TREADY + sd.send_peer[k], s_epoch, &sd.ctl->timed_out);
    }
    __syncthreads();
    const long long i = sd.send_off[k] + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < sd.send_off[k + 1]) y[sd.send_idx[i]] += __ldcg(sd.rinbox + i);
}

__global__ void shard_tack_kernel(ShardDev sd) {
    const unsigned ep = *reinterpret_cast<volatile unsigned*>(&sd.ctl->tepoch) + 1u;
    if ((int)threadIdx.x < sd.nsend) st_release_sys_u32(sd.send_flags[threadIdx.x] + F_TACK + sd.me, ep);
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned*>(&sd.ctl->tepoch) = ep;
}

__global__ void shard_scatter_add_kernel(const double* __restrict__ src, const int32_t* __restrict__ idx, long long count, double* __restrict__ y) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < count) y[idx[i]] += src[i];          // the ids of one peer's list are distinct
}

__global__ void shard_pack_kernel(const double* __restrict__ x, const int32_t* __restrict__ idx, long long count, double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < count) out[i] = x[idx[i]];
}

// ---- shard construction ---------------------------------------------------------------------------------------------
__global__ void flag_owned_kernel(const int32_t* __restrict__ owner, int64_t nc, int rank, int* __restrict__ flag) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nc) flag[i] = owner[i] == rank ? 1 : 0;
}
__global__ void compact_kernel(const int* __restrict__ flag, const int* __restrict__ pos, int64_t nc, int32_t* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nc && flag[i]) out[pos[i]] = (int32_t)i;
}
__global__ void gather_coords_kernel(const double* __restrict__ X, const int32_t* __restrict__ ids, int64_t count, int dim, double* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < count)
        for (int a = 0; a < dim; ++a) out[i * dim + a] = X[(int64_t)ids[i] * dim + a];
}
struct Box { double lo[3], hi[3]; };
// per owned node: exactness proof, referenced-candidate marks, boundary-row flag
__global__ void check_mark_kernel(const int32_t* __restrict__ st, const double* __restrict__ d2, const double* __restrict__ Q, int64_t nq, int n,
                                  int dim, Box box, const int32_t* __restrict__ owner, int rank, int* __restrict__ ref, int* __restrict__ brow,
                                  int* __restrict__ bad) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const double r2 = d2[q * n + n - 1];
    bool ok = true;
    for (int a = 0; a < dim; ++a) {
        const double x = Q[q * dim + a];
        const double glo = x - box.lo[a], ghi = box.hi[a] - x;      // +infinity where nothing lies beyond the face
        ok = ok && glo > 0.0 && ghi > 0.0 && glo * glo > r2 && ghi * ghi > r2;
    }
    if (!ok) *bad = 1;
    int b = 0;
    for (int j = 0; j < n; ++j) {
        const int c = st[q * n + j];
        if (owner[c] != rank) { b = 1; ref[c] = 1; }
    }
    brow[q] = b;
}
__global__ void not_kernel(const int* __restrict__ in, int64_t count, int* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < count) out[i] = in[i] ? 0 : 1;
}
__global__ void new_row_kernel(const int* __restrict__ brow, const int* __restrict__ ipos, const int* __restrict__ bpos, int64_t nq, int n_int,
                               const int32_t* __restrict__ own_c, int32_t* __restrict__ new_of_q, int32_t* __restrict__ map_c) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int r = brow[q] ? n_int + bpos[q] : ipos[q];
    new_of_q[q] = r;
    map_c[own_c[q]] = r;
}
__global__ void gather_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ ids, int64_t count, int32_t* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < count) out[i] = src[ids[i]];
}
__global__ void halo_map_kernel(const int32_t* __restrict__ halo_c, int64_t nh, int32_t base, int32_t* __restrict__ map_c) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nh) map_c[halo_c[i]] = base + (int32_t)i;
}
__global__ void local_nodes_kernel(const double* __restrict__ Xc, const int64_t* __restrict__ gidc, int dim, const int32_t* __restrict__ own_c,
                                   const int32_t* __restrict__ new_of_q, int64_t nq, const int32_t* __restrict__ halo_c, int64_t nh,
                                   double* __restrict__ Xl, int64_t* __restrict__ gidl) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nq + nh) return;
    const int c = i < nq ? own_c[i] : halo_c[i - nq];
    const int64_t dst = i < nq ? new_of_q[i] : i;
    for (int a = 0; a < dim; ++a) Xl[dst * dim + a] = Xc[(int64_t)c * dim + a];
    gidl[dst] = gidc[c];
}
__global__ void remap_stencils_kernel(const int32_t* __restrict__ st, const int32_t* __restrict__ new_of_q, const int32_t* __restrict__ map_c,
                                      int64_t nq, int n, int32_t* __restrict__ out) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nq * n) return;
    const int64_t q = e / n;
    out[(int64_t)new_of_q[q] * n + (e - q * n)] = map_c[st[e]];
}

__device__ __forceinline__ double box_uniform(uint64_t seed, uint64_t lin, int axis) {      // same generator as lattice_kernel (api.cu)
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (lin * 3ull + (uint64_t)axis + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * 0x1.0p-53;
}
struct LatBox { long long lo[3], ext[3]; int blocks[3]; int have_blocks; };
__global__ void lattice_box_kernel(int dim, long long g, uint64_t seed, LatBox bx, long long count, double* __restrict__ X,
                                   int64_t* __restrict__ gid, int32_t* __restrict__ owner) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= count) return;
    // enumeration order == ascending global linear id: axis 0 fastest, the last axis slowest
    long long c[3] = {0, 0, 0}, rem = t;
    for (int a = 0; a < dim; ++a) { c[a] = bx.lo[a] + rem % bx.ext[a]; rem /= bx.ext[a]; }
    const long long lin = dim == 2 ? c[1] * g + c[0] : (c[2] * g + c[1]) * g + c[0];
    for (int a = 0; a < dim; ++a) {
        const double u = box_uniform(seed, (uint64_t)lin, a);
        X[t * dim + a] = ((double)c[a] + 0.5 + 0.5 * (u - 0.5)) / (double)g;
    }
    if (gid) gid[t] = lin;
    if (owner && bx.have_blocks) {
        int o = 0;
        for (int a = 0; a < dim; ++a) o = o * bx.blocks[a] + (int)((c[a] * bx.blocks[a]) / g);
        owner[t] = o;
    }
}

cudaError_t exclusive_scan(rbffd_context* ctx, const int* in, int* out, int64_t count) {
    size_t bytes = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)count, ctx->stream);
    if (e != cudaSuccess) return e;
    DevBuf<unsigned char> tmp;
    e = tmp.alloc(bytes, ctx->stream);
    if (e != cudaSuccess) return e;
    return cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)count, ctx->stream);
}

}  // namespace

struct rbffd_shard {
    rbffd_context* ctx = nullptr;
    int dim = 0, n = 0, rank = 0, nparts = 0;
    int64_t n_owned = 0, n_int = 0, n_halo = 0;
    double* X_local = nullptr;          // [n_owned + n_halo][dim]
    int32_t* stencils = nullptr;        // [n_owned][n]
    int64_t* gid = nullptr;             // [n_owned + n_halo]
    std::vector<int64_t> gid_host;
    int64_t recv_count[MAXP] = {}, recv_off[MAXP + 1] = {};      // by peer rank
    std::vector<int32_t> send_ids[MAXP];                         // local owned ids asked for by every peer
    bool send_set[MAXP] = {};
    int32_t* send_idx = nullptr;
    void* inbox = nullptr;              // IPC buffer: [halo values | flag block | reverse inbox]
    size_t flags_off = 0, rinbox_off = 0, inbox_bytes = 0;
    void* peer_base[MAXP] = {};
    ShardCtl* ctl = nullptr;
    double* text = nullptr;             // [n_owned + n_halo] scratch of the transposed product
    ShardDev dev{};
    bool finalized = false;
};

namespace {

int ensure_gid_host(rbffd_shard* s) {
    if (!s->gid_host.empty()) return RBFFD_OK;
    rbffd_context* ctx = s->ctx;
    s->gid_host.resize((size_t)(s->n_owned + s->n_halo));
    CUDA_TRY(ctx, cudaMemcpyAsync(s->gid_host.data(), s->gid, sizeof(int64_t) * s->gid_host.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RBFFD_OK;
}

template <int TPR, int VEC, int ITERS, int MODE>
int launch_shard(rbffd_shard* s, rbffd_operator* op, int nm, const SpmvMats& m, const double* x, const SpmvEpilogue& ep, double* y) {
    rbffd_context* ctx = s->ctx;
    const int rows_per_block = 256 / TPR;
    const int PB = MODE == 0 ? s->dev.chunk_off[s->dev.nsend] : 0;
    const int IB = ceil_div_i64(s->n_int, rows_per_block);
    const int BB = ceil_div_i64(s->n_owned - s->n_int, rows_per_block);
    const int grid = PB + IB + BB;
    if (grid == 0) return RBFFD_OK;
    cudaStream_t st = ctx->stream;
    switch (nm) {
        case 1: shard_spmv_kernel<TPR, 1, VEC, ITERS, MODE><<<grid, 256, 0, st>>>(s->dev, op->n, op->colind, m, x, ep, y, PB, IB); break;
        case 2: shard_spmv_kernel<TPR, 2, VEC, ITERS, MODE><<<grid, 256, 0, st>>>(s->dev, op->n, op->colind, m, x, ep, y, PB, IB); break;
        case 3: shard_spmv_kernel<TPR, 3, VEC, ITERS, MODE><<<grid, 256, 0, st>>>(s->dev, op->n, op->colind, m, x, ep, y, PB, IB); break;
        case 4: shard_spmv_kernel<TPR, 4, VEC, ITERS, MODE><<<grid, 256, 0, st>>>(s->dev, op->n, op->colind, m, x, ep, y, PB, IB); break;
        case 5: shard_spmv_kernel<TPR, 5, VEC, ITERS, MODE><<<grid, 256, 0, st>>>(s->dev, op->n, op->colind, m, x, ep, y, PB, IB); break;
        default: shard_spmv_kernel<TPR, 6, VEC, ITERS, MODE><<<grid, 256, 0, st>>>(s->dev, op->n, op->colind, m, x, ep, y, PB, IB); break;
    }
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

template <int MODE>
int dispatch_shard(rbffd_shard* s, rbffd_operator* op, int nterms, const int32_t* which, const double* coef, const double* x,
                   const SpmvEpilogue& ep, double* y) {
    rbffd_context* ctx = s->ctx;
    if (!op || !which || !coef || !x || !y) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_spmv: NULL argument");
    if (nterms < 1 || nterms > SPMV_MAXMAT) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_spmv: 1..%d terms per launch (combine the operators first: rbffd_operator_combine_device)", SPMV_MAXMAT);
    if (op->M != s->n_owned || op->N != s->n_owned + s->n_halo)
        RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_spmv: operator is %lld x %lld, the shard needs %lld x %lld", (long long)op->M, (long long)op->N,
                   (long long)s->n_owned, (long long)(s->n_owned + s->n_halo));
    if (MODE == 0 && !s->finalized) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_spmv: rbffd_shard_finalize / _connect have not run");
    if (ep.su && (y == x || y == ep.su)) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_spmv_stage: out must not alias x or u");
    SpmvMats m{};
    const size_t stride = (size_t)op->M * op->n;
    for (int i = 0; i < nterms; ++i) {
        if (which[i] < 0 || which[i] >= op->nmat) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_spmv: matrix index %d out of range", which[i]);
        m.v[i] = op->vals + stride * which[i];
        m.c[i] = coef[i];
    }
    const int n = op->n;
    bool even = (n % 2) == 0 && reinterpret_cast<uintptr_t>(op->colind) % 8 == 0;
    for (int i = 0; i < nterms; ++i) even = even && reinterpret_cast<uintptr_t>(m.v[i]) % 16 == 0;
    if (even) {
        if (n <= 16) return launch_shard<4, 2, 2, MODE>(s, op, nterms, m, x, ep, y);
        if (n <= 32) return launch_shard<8, 2, 2, MODE>(s, op, nterms, m, x, ep, y);
        if (n <= 48) return launch_shard<8, 2, 3, MODE>(s, op, nterms, m, x, ep, y);
        if (n <= 64) return launch_shard<8, 2, 4, MODE>(s, op, nterms, m, x, ep, y);
        if (n <= 128) return launch_shard<16, 2, 4, MODE>(s, op, nterms, m, x, ep, y);
    } else {
        if (n <= 32) return launch_shard<8, 1, 4, MODE>(s, op, nterms, m, x, ep, y);
        if (n <= 64) return launch_shard<8, 1, 8, MODE>(s, op, nterms, m, x, ep, y);
        if (n <= 128) return launch_shard<16, 1, 8, MODE>(s, op, nterms, m, x, ep, y);
    }
    RBFFD_FAIL(ctx, RBFFD_ERR_UNSUPPORTED, "shard_spmv: row length %d > 128", n);
}

// smallest cut surface over all factorisations of nparts into dim factors
void choose_blocks(const double* ext, int dim, int nparts, int* best) {
    double best_cost = std::numeric_limits<double>::infinity();
    int b[3] = {1, 1, 1};
    for (b[0] = 1; b[0] <= nparts; ++b[0]) {
        if (nparts % b[0]) continue;
        const int r0 = nparts / b[0];
        for (b[1] = 1; b[1] <= r0; ++b[1]) {
            if (r0 % b[1]) continue;
            b[2] = r0 / b[1];
            if (dim == 2 && b[2] != 1) continue;
            double cost = 0.0;
            for (int a = 0; a < dim; ++a) {
                double area = 1.0;
                for (int c = 0; c < dim; ++c) if (c != a) area *= std::max(ext[c], 1e-300);
                cost += (b[a] - 1) * area;
            }
            if (cost < best_cost) { best_cost = cost; best[0] = b[0]; best[1] = b[1]; best[2] = b[2]; }
        }
    }
}

// splits idx[lo, hi) into nb groups of equal size along `axis` (coordinate quantiles), then recurses over the next axis
void split_axis(const double* X, int dim, int64_t* idx, int64_t lo, int64_t hi, int axis, const int* blocks, int label, int32_t* part) {
    if (axis == dim) {
        for (int64_t i = lo; i < hi; ++i) part[idx[i]] = label;
        return;
    }
    const int nb = blocks[axis];
    const int64_t len = hi - lo;
    int64_t prev = lo;
    for (int k = 0; k < nb; ++k) {
        const int64_t cut = lo + (len * (k + 1)) / nb;
        if (k + 1 < nb && cut > prev && cut < hi)
            std::nth_element(idx + prev, idx + cut, idx + hi, [&](int64_t a, int64_t b) {
                const double xa = X[a * dim + axis], xb = X[b * dim + axis];
                return xa < xb || (xa == xb && a < b);
            });
        split_axis(X, dim, idx, prev, cut, axis + 1, blocks, label * nb + k, part);
        prev = cut;
    }
}

}  // namespace

extern "C" {

int rbffd_shard_plan_host(const double* X, int64_t N, int32_t dim, int32_t nparts, const int32_t* blocks, int32_t* part_out) {
    if (!X || !part_out || N < 1 || dim < 2 || dim > 3 || nparts < 1 || nparts > MAXP) return RBFFD_ERR_INVALID;
    int b[3] = {1, 1, 1};
    bool given = blocks != nullptr;
    if (given) {
        int prod = 1;
        for (int a = 0; a < dim; ++a) { b[a] = blocks[a]; given = given && b[a] > 0; prod *= std::max(b[a], 1); }
        if (given && prod != nparts) return RBFFD_ERR_INVALID;
    }
    if (!given) {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, ext[3] = {1, 1, 1};
        for (int64_t i = 0; i < N; ++i)
            for (int a = 0; a < dim; ++a) { lo[a] = std::min(lo[a], X[i * dim + a]); hi[a] = std::max(hi[a], X[i * dim + a]); }
        for (int a = 0; a < dim; ++a) ext[a] = hi[a] - lo[a];
        choose_blocks(ext, dim, nparts, b);
    }
    std::vector<int64_t> idx((size_t)N);
    std::iota(idx.begin(), idx.end(), (int64_t)0);
    split_axis(X, dim, idx.data(), 0, N, 0, b, 0, part_out);
    return RBFFD_OK;
}

int rbffd_shard_create_device(rbffd_context* ctx, int32_t dim, int32_t n, const double* Xc, const int64_t* gidc, const int32_t* owner,
                              int64_t nc, const double* box_lo, const double* box_hi, int32_t rank, int32_t nparts, rbffd_shard** out) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (!out || !Xc || !gidc || !owner || !box_lo || !box_hi || dim < 2 || dim > 3 || n < 1 || nc < 1 || nparts < 1 || nparts > MAXP || rank < 0 ||
        rank >= nparts)
        RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_create: bad arguments");
    if (nc > 0x7fffffff - 1024) RBFFD_FAIL(ctx, RBFFD_ERR_UNSUPPORTED, "shard_create: more than 2^31 candidate nodes on one rank");
    *out = nullptr;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int B = 256;
    DevBuf<int> flag, pos;
    CUDA_TRY(ctx, flag.alloc(nc + 1, st));
    CUDA_TRY(ctx, pos.alloc(nc + 1, st));
    CUDA_TRY(ctx, cudaMemsetAsync(flag.p + nc, 0, sizeof(int), st));
    flag_owned_kernel<<<ceil_div_i64(nc, B), B, 0, st>>>(owner, nc, rank, flag.p);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, exclusive_scan(ctx, flag.p, pos.p, nc + 1));
    int n_owned_i = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&n_owned_i, pos.p + nc, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    const int64_t nq = n_owned_i;
    if (nq < 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_create: rank %d owns no node", rank);
    if (n > nc) RBFFD_FAIL(ctx, RBFFD_ERR_K_TOO_LARGE, "shard_create: n=%d exceeds the %lld nodes held by rank %d", n, (long long)nc, rank);
    DevBuf<int32_t> own_c, st_c, new_of_q, map_c;
    DevBuf<double> Q, d2;
    CUDA_TRY(ctx, own_c.alloc(nq, st));
    compact_kernel<<<ceil_div_i64(nc, B), B, 0, st>>>(flag.p, pos.p, nc, own_c.p);
    CUDA_TRY(ctx, Q.alloc(nq * dim, st));
    gather_coords_kernel<<<ceil_div_i64(nq, B), B, 0, st>>>(Xc, own_c.p, nq, dim, Q.p);
    KLAUNCH(ctx); KLAUNCH(ctx);
    CUDA_TRY(ctx, st_c.alloc(nq * n, st));
    CUDA_TRY(ctx, d2.alloc(nq * n, st));
    // exact kNN of the owned nodes among all candidates; candidates are in ascending global id, so ties are broken as on one GPU
    RBFFD_TRY(rbffd_knn_impl(ctx, Xc, nc, dim, Q.p, nq, n, nullptr, nullptr, false, st_c.p, d2.p));
    DevBuf<int> ref, brow, nbrow, ipos, bpos, bad;
    CUDA_TRY(ctx, ref.alloc(nc + 1, st));
    CUDA_TRY(ctx, brow.alloc(nq + 1, st));
    CUDA_TRY(ctx, nbrow.alloc(nq + 1, st));
    CUDA_TRY(ctx, ipos.alloc(nq + 1, st));
    CUDA_TRY(ctx, bpos.alloc(nq + 1, st));
    CUDA_TRY(ctx, bad.alloc(1, st));
    CUDA_TRY(ctx, cudaMemsetAsync(ref.p, 0, sizeof(int) * (nc + 1), st));
    CUDA_TRY(ctx, cudaMemsetAsync(brow.p + nq, 0, sizeof(int), st));
    CUDA_TRY(ctx, cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    Box box;
    for (int a = 0; a < 3; ++a) { box.lo[a] = a < dim ? box_lo[a] : -INFINITY; box.hi[a] = a < dim ? box_hi[a] : INFINITY; }
    check_mark_kernel<<<ceil_div_i64(nq, B), B, 0, st>>>(st_c.p, d2.p, Q.p, nq, n, dim, box, owner, rank, ref.p, brow.p, bad.p);
    KLAUNCH(ctx);
    not_kernel<<<ceil_div_i64(nq + 1, B), B, 0, st>>>(brow.p, nq + 1, nbrow.p);
    CUDA_TRY(ctx, exclusive_scan(ctx, nbrow.p, ipos.p, nq + 1));
    CUDA_TRY(ctx, exclusive_scan(ctx, brow.p, bpos.p, nq + 1));
    CUDA_TRY(ctx, exclusive_scan(ctx, ref.p, pos.p, nc + 1));          // pos now numbers the referenced foreign candidates
    int h_bad = 0, n_bnd = 0, n_halo_i = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(&n_bnd, bpos.p + nq, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(&n_halo_i, pos.p + nc, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (h_bad) RBFFD_FAIL(ctx, RBFFD_ERR_HALO, "shard_create: the candidate halo of rank %d is too thin for exact stencils (n=%d): widen the margin", rank, n);
    const int64_t nh = n_halo_i, n_int = nq - n_bnd;
    // halo = referenced foreign candidates, grouped by owner (stable: ascending global id inside every group)
    DevBuf<int32_t> halo_c, halo_sorted, own_h, own_h_sorted;
    CUDA_TRY(ctx, halo_c.alloc(nh, st));
    CUDA_TRY(ctx, halo_sorted.alloc(nh, st));
    CUDA_TRY(ctx, own_h.alloc(nh, st));
    CUDA_TRY(ctx, own_h_sorted.alloc(nh, st));
    std::vector<int32_t> h_owner((size_t)nh);
    if (nh > 0) {
        compact_kernel<<<ceil_div_i64(nc, B), B, 0, st>>>(ref.p, pos.p, nc, halo_c.p);
        gather_i32_kernel<<<ceil_div_i64(nh, B), B, 0, st>>>(owner, halo_c.p, nh, own_h.p);
        KLAUNCH(ctx); KLAUNCH(ctx);
        size_t bytes = 0;
        CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, bytes, own_h.p, own_h_sorted.p, halo_c.p, halo_sorted.p, (int)nh, 0, 8, st));
        DevBuf<unsigned char> tmp;
        CUDA_TRY(ctx, tmp.alloc(bytes, st));
        CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, bytes, own_h.p, own_h_sorted.p, halo_c.p, halo_sorted.p, (int)nh, 0, 8, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(h_owner.data(), own_h_sorted.p, sizeof(int32_t) * nh, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(ctx, map_c.alloc(nc, st));
    CUDA_TRY(ctx, new_of_q.alloc(nq, st));
    new_row_kernel<<<ceil_div_i64(nq, B), B, 0, st>>>(brow.p, ipos.p, bpos.p, nq, (int)n_int, own_c.p, new_of_q.p, map_c.p);
    KLAUNCH(ctx);
    if (nh > 0) { halo_map_kernel<<<ceil_div_i64(nh, B), B, 0, st>>>(halo_sorted.p, nh, (int32_t)nq, map_c.p); KLAUNCH(ctx); }
    rbffd_shard* s = new (std::nothrow) rbffd_shard();
    if (!s) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_create: out of host memory");
    s->ctx = ctx; s->dim = dim; s->n = n; s->rank = rank; s->nparts = nparts;
    s->n_owned = nq; s->n_int = n_int; s->n_halo = nh;
    auto fail = [&](cudaError_t e) { rbffd_shard_destroy(s); ctx->err = std::string("CUDA error in shard_create: ") + cudaGetErrorString(e); return RBFFD_ERR_CUDA; };
    cudaError_t e;
    if ((e = cudaMalloc((void**)&s->X_local, sizeof(double) * (size_t)(nq + nh) * dim)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc((void**)&s->gid, sizeof(int64_t) * (size_t)(nq + nh))) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc((void**)&s->stencils, sizeof(int32_t) * (size_t)nq * n)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc((void**)&s->ctl, sizeof(ShardCtl))) != cudaSuccess) return fail(e);
    if ((e = cudaMemsetAsync(s->ctl, 0, sizeof(ShardCtl), st)) != cudaSuccess) return fail(e);
    local_nodes_kernel<<<ceil_div_i64(nq + nh, B), B, 0, st>>>(Xc, gidc, dim, own_c.p, new_of_q.p, nq, halo_sorted.p, nh, s->X_local, s->gid);
    remap_stencils_kernel<<<ceil_div_i64(nq * n, B), B, 0, st>>>(st_c.p, new_of_q.p, map_c.p, nq, n, s->stencils);
    KLAUNCH(ctx); KLAUNCH(ctx);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(e);
    for (int64_t i = 0; i < nh; ++i) {
        const int o = h_owner[(size_t)i];
        if (o < 0 || o >= nparts || o == rank) { rbffd_shard_destroy(s); RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_create: owner %d of a halo node is out of range", o); }
        s->recv_count[o]++;
    }
    for (int p = 0; p < nparts; ++p) s->recv_off[p + 1] = s->recv_off[p] + s->recv_count[p];
    *out = s;
    return RBFFD_OK;
}

int rbffd_shard_create_host(rbffd_context* ctx, const double* X, int64_t N, int32_t dim, const int32_t* part, int32_t nparts,
                            int32_t rank, int32_t n, rbffd_shard** out) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (!X || !part || !out || N < 1 || dim < 2 || dim > 3 || nparts < 1 || nparts > MAXP || rank < 0 || rank >= nparts || n < 1)
        RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_create_host: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    double glo[3] = {1e300, 1e300, 1e300}, ghi[3] = {-1e300, -1e300, -1e300}, olo[3] = {1e300, 1e300, 1e300}, ohi[3] = {-1e300, -1e300, -1e300};
    int64_t n_owned = 0;
    for (int64_t i = 0; i < N; ++i) {
        const bool mine = part[i] == rank;
        n_owned += mine;
        for (int a = 0; a < dim; ++a) {
            const double x = X[i * dim + a];
            if (!(x == x) || std::isinf(x)) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_create_host: non-finite coordinate at node %lld", (long long)i);
            glo[a] = std::min(glo[a], x); ghi[a] = std::max(ghi[a], x);
            if (mine) { olo[a] = std::min(olo[a], x); ohi[a] = std::max(ohi[a], x); }
        }
    }
    if (n_owned < 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_create_host: rank %d owns no node", rank);
    if (n > N) RBFFD_FAIL(ctx, RBFFD_ERR_K_TOO_LARGE, "shard_create_host: n=%d exceeds the number of points %lld", n, (long long)N);
    // first guess of the margin: radius of a ball holding n nodes at the owned block's mean density, times 1.5
    double vol = 1.0;
    for (int a = 0; a < dim; ++a) vol *= std::max(ohi[a] - olo[a], 1e-12 * std::max(1.0, std::fabs(ohi[a])));
    const double cd = dim == 2 ? M_PI : 4.0 * M_PI / 3.0;
    double margin = 1.5 * std::pow((double)n * vol / ((double)n_owned * cd), 1.0 / dim);
    double gext = 0.0;
    for (int a = 0; a < dim; ++a) gext = std::max(gext, ghi[a] - glo[a]);
    std::vector<double> hx;
    std::vector<int64_t> hg;
    std::vector<int32_t> ho;
    for (int attempt = 0; attempt < 24; ++attempt) {
        double blo[3], bhi[3];
        bool covers_all = true;
        for (int a = 0; a < dim; ++a) {
            blo[a] = olo[a] - margin; bhi[a] = ohi[a] + margin;
            covers_all = covers_all && blo[a] < glo[a] && bhi[a] > ghi[a];
        }
        hx.clear(); hg.clear(); ho.clear();
        for (int64_t i = 0; i < N; ++i) {
            bool in = true;
            for (int a = 0; a < dim; ++a) { const double x = X[i * dim + a]; in = in && x >= blo[a] && x <= bhi[a]; }
            if (!in) continue;
            for (int a = 0; a < dim; ++a) hx.push_back(X[i * dim + a]);
            hg.push_back(i);
            ho.push_back(part[i]);
        }
        const int64_t nc = (int64_t)hg.size();
        // faces beyond which no node exists at all are open
        for (int a = 0; a < dim; ++a) {
            if (blo[a] < glo[a]) blo[a] = -INFINITY;
            if (bhi[a] > ghi[a]) bhi[a] = INFINITY;
        }
        int rc = RBFFD_ERR_HALO;
        if (nc >= n) {
            DevBuf<double> dX;
            DevBuf<int64_t> dG;
            DevBuf<int32_t> dO;
            CUDA_TRY(ctx, dX.alloc(nc * dim, ctx->stream));
            CUDA_TRY(ctx, dG.alloc(nc, ctx->stream));
            CUDA_TRY(ctx, dO.alloc(nc, ctx->stream));
            CUDA_TRY(ctx, cudaMemcpyAsync(dX.p, hx.data(), sizeof(double) * nc * dim, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(ctx, cudaMemcpyAsync(dG.p, hg.data(), sizeof(int64_t) * nc, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(ctx, cudaMemcpyAsync(dO.p, ho.data(), sizeof(int32_t) * nc, cudaMemcpyHostToDevice, ctx->stream));
            rc = rbffd_shard_create_device(ctx, dim, n, dX.p, dG.p, dO.p, nc, blo, bhi, rank, nparts, out);
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        }
        if (rc != RBFFD_ERR_HALO) return rc;
        if (covers_all) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_create_host: exactness proof failed although every node is a candidate");
        margin = std::min(1.6 * margin, 2.0 * gext + 1.0);
    }
    RBFFD_FAIL(ctx, RBFFD_ERR_HALO, "shard_create_host: no sufficient halo margin found");
}

int rbffd_shard_destroy(rbffd_shard* s) {
    if (!s) return RBFFD_OK;
    rbffd_context* ctx = s->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int p = 0; p < MAXP; ++p)
        if (s->peer_base[p]) cudaIpcCloseMemHandle(s->peer_base[p]);
    if (s->inbox) cudaFree(s->inbox);
    if (s->X_local) cudaFree(s->X_local);
    if (s->gid) cudaFree(s->gid);
    if (s->stencils) cudaFree(s->stencils);
    if (s->send_idx) cudaFree(s->send_idx);
    if (s->ctl) cudaFree(s->ctl);
    if (s->text) cudaFree(s->text);
    delete s;
    return RBFFD_OK;
}

int rbffd_shard_status(rbffd_shard* s) {
    if (!s) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = s->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    unsigned flag = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&flag, &s->ctl->timed_out, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag) RBFFD_FAIL(ctx, RBFFD_ERR_HALO, "shard %d: a halo exchange timed out waiting for a peer (crashed rank or mismatched sequence of rbffd_shard_spmv* calls); results since then are invalid", s->rank);
    return RBFFD_OK;
}

int rbffd_shard_info(const rbffd_shard* s, int64_t* n_owned, int64_t* n_interior, int64_t* n_halo) {
    if (!s) return RBFFD_ERR_INVALID;
    if (n_owned) *n_owned = s->n_owned;
    if (n_interior) *n_interior = s->n_int;
    if (n_halo) *n_halo = s->n_halo;
    return RBFFD_OK;
}

int rbffd_shard_global_ids_host(rbffd_shard* s, int32_t index_base, int64_t* gid_out) {
    if (!s) return RBFFD_ERR_INVALID;
    if (!gid_out) RBFFD_FAIL(s->ctx, RBFFD_ERR_INVALID, "shard_global_ids: NULL output");
    RBFFD_TRY(ensure_gid_host(s));
    for (size_t i = 0; i < s->gid_host.size(); ++i) gid_out[i] = s->gid_host[i] + index_base;
    return RBFFD_OK;
}

int rbffd_shard_device_arrays(const rbffd_shard* s, const double** X_local, const int32_t** stencils) {
    if (!s) return RBFFD_ERR_INVALID;
    if (X_local) *X_local = s->X_local;
    if (stencils) *stencils = s->stencils;
    return RBFFD_OK;
}

int rbffd_shard_recv_count(const rbffd_shard* s, int32_t peer, int64_t* count) {
    if (!s || !count || peer < 0 || peer >= s->nparts) return RBFFD_ERR_INVALID;
    *count = s->recv_count[peer];
    return RBFFD_OK;
}

int rbffd_shard_recv_ids_host(rbffd_shard* s, int32_t peer, int32_t index_base, int64_t* ids_out) {
    if (!s) return RBFFD_ERR_INVALID;
    if (peer < 0 || peer >= s->nparts || (!ids_out && s->recv_count[peer] > 0)) RBFFD_FAIL(s->ctx, RBFFD_ERR_INVALID, "shard_recv_ids: bad arguments");
    RBFFD_TRY(ensure_gid_host(s));
    for (int64_t i = 0; i < s->recv_count[peer]; ++i) ids_out[i] = s->gid_host[(size_t)(s->n_owned + s->recv_off[peer] + i)] + index_base;
    return RBFFD_OK;
}

int rbffd_shard_set_send_ids_host(rbffd_shard* s, int32_t peer, int32_t index_base, const int64_t* ids, int64_t count) {
    if (!s) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = s->ctx;
    if (peer < 0 || peer >= s->nparts || peer == s->rank || count < 0 || (count > 0 && !ids)) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_set_send_ids: bad arguments");
    if (s->finalized) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_set_send_ids: the shard is already finalized");
    RBFFD_TRY(ensure_gid_host(s));
    // owned nodes are numbered [interior | boundary], each run in ascending global id: two binary searches, no sort
    const int64_t* g0 = s->gid_host.data();
    const int64_t* runs[3] = {g0, g0 + s->n_int, g0 + s->n_owned};
    std::vector<int32_t>& out = s->send_ids[peer];
    out.resize((size_t)count);
    for (int64_t i = 0; i < count; ++i) {
        const int64_t g = ids[i] - index_base;
        int32_t found = -1;
        for (int r = 0; r < 2 && found < 0; ++r) {
            const int64_t* it = std::lower_bound(runs[r], runs[r + 1], g);
            if (it != runs[r + 1] && *it == g) found = (int32_t)(it - g0);
        }
        if (found < 0)
            RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_set_send_ids: rank %d asked rank %d for node %lld, which it does not own", peer, s->rank, (long long)g);
        out[(size_t)i] = found;
    }
    s->send_set[peer] = true;
    return RBFFD_OK;
}

int rbffd_shard_finalize(rbffd_shard* s, unsigned char* handle64, int64_t* offsets) {
    if (!s) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = s->ctx;
    if (!handle64 || !offsets) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_finalize: NULL output");
    if (s->finalized) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_finalize: called twice");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ShardDev& d = s->dev;
    d = ShardDev{};
    d.me = s->rank;
    std::vector<int32_t> flat;
    for (int p = 0; p < s->nparts; ++p) {
        if (p == s->rank || s->send_ids[p].empty()) continue;
        const int k = d.nsend++;
        d.send_peer[k] = p;
        d.send_off[k + 1] = d.send_off[k] + (long long)s->send_ids[p].size();
        d.chunk_off[k + 1] = d.chunk_off[k] + (int)((s->send_ids[p].size() + PUSH_CHUNK - 1) / PUSH_CHUNK);
        flat.insert(flat.end(), s->send_ids[p].begin(), s->send_ids[p].end());
    }
    for (int p = 0; p < s->nparts; ++p) {
        if (s->recv_count[p] == 0) continue;
        const int r = d.nrecv++;
        d.recv_peer[r] = p;
        d.recv_off[r] = s->recv_off[p];
        d.recv_off[r + 1] = s->recv_off[p] + s->recv_count[p];
        d.rchunk_off[r + 1] = d.rchunk_off[r] + (int)((s->recv_count[p] + PUSH_CHUNK - 1) / PUSH_CHUNK);
    }
    const size_t total_send = flat.size();
    s->flags_off = (((size_t)s->n_halo * 8 + 255) / 256) * 256;
    s->rinbox_off = s->flags_off + 256 * ((F_WORDS * 4 + 255) / 256);
    s->inbox_bytes = s->rinbox_off + std::max<size_t>(total_send, 1) * 8;
    CUDA_TRY(ctx, cudaMalloc(&s->inbox, s->inbox_bytes));
    CUDA_TRY(ctx, cudaMemset(s->inbox, 0, s->inbox_bytes));
    CUDA_TRY(ctx, cudaMalloc((void**)&s->send_idx, sizeof(int32_t) * std::max<size_t>(total_send, 1)));
    if (total_send) CUDA_TRY(ctx, cudaMemcpy(s->send_idx, flat.data(), sizeof(int32_t) * total_send, cudaMemcpyHostToDevice));
    cudaIpcMemHandle_t h;
    CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, s->inbox));
    memcpy(handle64, &h, 64);
    d.send_idx = s->send_idx;
    d.inbox = reinterpret_cast<double*>(s->inbox);
    d.flags = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(s->inbox) + s->flags_off);
    d.rinbox = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(s->inbox) + s->rinbox_off);
    d.ctl = s->ctl;
    d.n_owned = s->n_owned;
    d.n_int = s->n_int;
    // where every peer's data lands in MY buffers (the peers need these to address their stores)
    for (int p = 0; p < s->nparts; ++p) { offsets[2 * p] = s->recv_off[p]; offsets[2 * p + 1] = -1; }
    for (int k = 0; k < d.nsend; ++k) offsets[2 * d.send_peer[k] + 1] = (int64_t)(s->rinbox_off / 8) + d.send_off[k];
    offsets[2 * s->nparts] = (int64_t)s->flags_off;
    s->finalized = true;
    return RBFFD_OK;
}

// fwd_offset: where MY values start in the peer's inbox (doubles) = the peer's offsets[2 * me]; rev_offset: where MY
// transposed-product contributions start in the peer's buffer (doubles) = the peer's offsets[2 * me + 1]; flags_offset: the
// peer's flag block (bytes) = the peer's offsets[2 * nparts]
int rbffd_shard_connect(rbffd_shard* s, int32_t peer, const unsigned char* handle64, int64_t fwd_offset, int64_t rev_offset, int64_t flags_offset) {
    if (!s) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = s->ctx;
    if (!s->finalized || !handle64 || peer < 0 || peer >= s->nparts || peer == s->rank || flags_offset < 0)
        RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_connect: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!s->peer_base[peer]) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handle64, 64);
        CUDA_TRY(ctx, cudaIpcOpenMemHandle(&s->peer_base[peer], h, cudaIpcMemLazyEnablePeerAccess));
    }
    unsigned char* base = reinterpret_cast<unsigned char*>(s->peer_base[peer]);
    unsigned* f = reinterpret_cast<unsigned*>(base + flags_offset);
    ShardDev& d = s->dev;
    for (int k = 0; k < d.nsend; ++k)
        if (d.send_peer[k] == peer) {
            if (fwd_offset < 0) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_connect: peer %d reserved no room for the values of rank %d", peer, s->rank);
            d.send_dst[k] = reinterpret_cast<double*>(base) + fwd_offset;
            d.send_flags[k] = f;
        }
    for (int r = 0; r < d.nrecv; ++r)
        if (d.recv_peer[r] == peer) {
            if (rev_offset < 0) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_connect: peer %d reserved no reverse room for rank %d", peer, s->rank);
            d.recv_rdst[r] = reinterpret_cast<double*>(base) + rev_offset;
            d.recv_flags[r] = f;
        }
    return RBFFD_OK;
}

int rbffd_shard_spmv_device(rbffd_shard* s, rbffd_operator* op, int32_t nterms, const int32_t* which, const double* coef, const double* x, double* y) {
    if (!s) return RBFFD_ERR_INVALID;
    CUDA_TRY(s->ctx, cudaSetDevice(s->ctx->device));
    const ShardDev& d = s->dev;
    for (int k = 0; k < d.nsend; ++k)
        if (!d.send_dst[k] || !d.send_flags[k]) RBFFD_FAIL(s->ctx, RBFFD_ERR_INVALID, "shard_spmv: peer %d is not connected", d.send_peer[k]);
    for (int r = 0; r < d.nrecv; ++r)
        if (!d.recv_flags[r]) RBFFD_FAIL(s->ctx, RBFFD_ERR_INVALID, "shard_spmv: peer %d is not connected", d.recv_peer[r]);
    return dispatch_shard<0>(s, op, nterms, which, coef, x, SpmvEpilogue{0.0, nullptr, 0.0, 0.0, 0.0}, y);
}

// out[0:n_owned] = a * u + b * (x + dt * sum_i coef[i] * D[which[i]] * [x ; halo of x]): one SSP-RK stage of the sharded
// semidiscretisation as ONE launch (halo exchange, product and stage update)
int rbffd_shard_spmv_stage_device(rbffd_shard* s, rbffd_operator* op, int32_t nterms, const int32_t* which, const double* coef,
                                  const double* x, double a, const double* u, double b, double dt, double* out) {
    if (!s) return RBFFD_ERR_INVALID;
    if (!u) RBFFD_FAIL(s->ctx, RBFFD_ERR_INVALID, "shard_spmv_stage: NULL argument");
    CUDA_TRY(s->ctx, cudaSetDevice(s->ctx->device));
    const ShardDev& d = s->dev;
    for (int k = 0; k < d.nsend; ++k)
        if (!d.send_dst[k] || !d.send_flags[k]) RBFFD_FAIL(s->ctx, RBFFD_ERR_INVALID, "shard_spmv_stage: peer %d is not connected", d.send_peer[k]);
    for (int r = 0; r < d.nrecv; ++r)
        if (!d.recv_flags[r]) RBFFD_FAIL(s->ctx, RBFFD_ERR_INVALID, "shard_spmv_stage: peer %d is not connected", d.recv_peer[r]);
    return dispatch_shard<0>(s, op, nterms, which, coef, x, SpmvEpilogue{0.0, u, a, b, dt}, out);
}

int rbffd_shard_spmv_local_device(rbffd_shard* s, rbffd_operator* op, int32_t nterms, const int32_t* which, const double* coef, const double* x, double* y) {
    if (!s) return RBFFD_ERR_INVALID;
    CUDA_TRY(s->ctx, cudaSetDevice(s->ctx->device));
    if (!s->inbox) {                     // single-process use without finalize: a plain inbox
        s->flags_off = (((size_t)s->n_halo * 8 + 255) / 256) * 256;
        s->inbox_bytes = s->flags_off + 256;
        CUDA_TRY(s->ctx, cudaMalloc(&s->inbox, s->inbox_bytes));
        CUDA_TRY(s->ctx, cudaMemset(s->inbox, 0, s->inbox_bytes));
        s->dev.inbox = reinterpret_cast<double*>(s->inbox);
        s->dev.n_owned = s->n_owned;
        s->dev.n_int = s->n_int;
        s->dev.me = s->rank;
    }
    return dispatch_shard<1>(s, op, nterms, which, coef, x, SpmvEpilogue{0.0, nullptr, 0.0, 0.0, 0.0}, y);
}

int rbffd_shard_pack_device(rbffd_shard* s, int32_t peer, const double* x, double* sendbuf) {
    if (!s) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = s->ctx;
    if (!s->finalized || peer < 0 || peer >= s->nparts || !x) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_pack: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const ShardDev& d = s->dev;
    for (int k = 0; k < d.nsend; ++k) {
        if (d.send_peer[k] != peer) continue;
        const long long cnt = d.send_off[k + 1] - d.send_off[k];
        if (cnt > 0 && !sendbuf) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_pack: NULL buffer");
        shard_pack_kernel<<<ceil_div_i64(cnt, 256), 256, 0, ctx->stream>>>(x, s->send_idx + d.send_off[k], cnt, sendbuf);
        KLAUNCH(ctx);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return RBFFD_OK;
}

int rbffd_shard_unpack_device(rbffd_shard* s, int32_t peer, const double* recvbuf) {
    if (!s) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = s->ctx;
    if (peer < 0 || peer >= s->nparts || (s->recv_count[peer] > 0 && (!recvbuf || !s->inbox))) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_unpack: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (s->recv_count[peer] > 0)
        CUDA_TRY(ctx, cudaMemcpyAsync(reinterpret_cast<double*>(s->inbox) + s->recv_off[peer], recvbuf, sizeof(double) * s->recv_count[peer],
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    return RBFFD_OK;
}

int rbffd_shard_spmv_t_device(rbffd_shard* s, rbffd_operator* op, int32_t which, double alpha, const double* v, double beta, double* y) {
    if (!s) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = s->ctx;
    if (!op || !v || !y) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_spmv_t: NULL argument");
    if (!s->finalized) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_spmv_t: rbffd_shard_finalize / _connect have not run");
    if (op->M != s->n_owned || op->N != s->n_owned + s->n_halo) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_spmv_t: operator shape does not match the shard");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const ShardDev& d = s->dev;
    for (int r = 0; r < d.nrecv; ++r)
        if (!d.recv_rdst[r] || !d.recv_flags[r]) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_spmv_t: peer %d is not connected", d.recv_peer[r]);
    for (int k = 0; k < d.nsend; ++k)
        if (!d.send_flags[k]) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_spmv_t: peer %d is not connected", d.send_peer[k]);
    if (!s->text) CUDA_TRY(ctx, cudaMalloc((void**)&s->text, sizeof(double) * (size_t)(s->n_owned + s->n_halo)));
    cudaStream_t st = ctx->stream;
    RBFFD_TRY(rbffd_spmv_t_impl(op, which, alpha, v, 0.0, s->text));             // all local columns: owned and halo
    if (d.rchunk_off[d.nrecv] > 0) { shard_tpush_kernel<<<d.rchunk_off[d.nrecv], 256, 0, st>>>(d, s->text); KLAUNCH(ctx); }
    shard_tinit_kernel<<<ceil_div_i64(s->n_owned, 256), 256, 0, st>>>(s->text, s->n_owned, beta, y);
    KLAUNCH(ctx);
    for (int k = 0; k < d.nsend; ++k) {                                            // rank order: deterministic sums
        const long long cnt = d.send_off[k + 1] - d.send_off[k];
        shard_tcombine_kernel<<<ceil_div_i64(cnt, 256), 256, 0, st>>>(d, k, y);
        KLAUNCH(ctx);
    }
    shard_tack_kernel<<<1, 32, 0, st>>>(d);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

int rbffd_shard_spmv_t_local_device(rbffd_shard* s, rbffd_operator* op, int32_t which, double alpha, const double* v, double beta, double* y) {
    if (!s) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = s->ctx;
    if (!op || !v || !y) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_spmv_t_local: NULL argument");
    if (op->M != s->n_owned || op->N != s->n_owned + s->n_halo) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_spmv_t_local: operator shape does not match the shard");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!s->text) CUDA_TRY(ctx, cudaMalloc((void**)&s->text, sizeof(double) * (size_t)(s->n_owned + s->n_halo)));
    RBFFD_TRY(rbffd_spmv_t_impl(op, which, alpha, v, 0.0, s->text));
    shard_tinit_kernel<<<ceil_div_i64(s->n_owned, 256), 256, 0, ctx->stream>>>(s->text, s->n_owned, beta, y);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

int rbffd_shard_tpack_device(rbffd_shard* s, int32_t peer, double* sendbuf) {
    if (!s) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = s->ctx;
    if (peer < 0 || peer >= s->nparts || !s->text || (s->recv_count[peer] > 0 && !sendbuf)) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_tpack: bad arguments (run rbffd_shard_spmv_t_local_device first)");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (s->recv_count[peer] > 0)
        CUDA_TRY(ctx, cudaMemcpyAsync(sendbuf, s->text + s->n_owned + s->recv_off[peer], sizeof(double) * s->recv_count[peer], cudaMemcpyDeviceToDevice, ctx->stream));
    return RBFFD_OK;
}

int rbffd_shard_tunpack_add_device(rbffd_shard* s, int32_t peer, const double* recvbuf, double* y) {
    if (!s) return RBFFD_ERR_INVALID;
    rbffd_context* ctx = s->ctx;
    if (!s->finalized || peer < 0 || peer >= s->nparts || !y) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_tunpack_add: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const ShardDev& d = s->dev;
    for (int k = 0; k < d.nsend; ++k) {
        if (d.send_peer[k] != peer) continue;
        const long long cnt = d.send_off[k + 1] - d.send_off[k];
        if (cnt > 0 && !recvbuf) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "shard_tunpack_add: NULL buffer");
        shard_scatter_add_kernel<<<ceil_div_i64(cnt, 256), 256, 0, ctx->stream>>>(recvbuf, s->send_idx + d.send_off[k], cnt, y);
        KLAUNCH(ctx);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return RBFFD_OK;
}

int rbffd_shard_send_count(const rbffd_shard* s, int32_t peer, int64_t* count) {
    if (!s || !count || peer < 0 || peer >= s->nparts) return RBFFD_ERR_INVALID;
    *count = (int64_t)s->send_ids[peer].size();
    return RBFFD_OK;
}

int rbffd_jittered_lattice_box_device(rbffd_context* ctx, int32_t dim, int64_t g, uint64_t seed, const int64_t* lo, const int64_t* hi,
                                      const int32_t* blocks, double* X_out, int64_t* gid_out, int32_t* owner_out) {
    if (!ctx) return RBFFD_ERR_INVALID;
    if (dim < 2 || dim > 3 || g < 1 || !lo || !hi || !X_out) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "jittered_lattice_box: bad arguments");
    LatBox bx{};
    long long count = 1;
    for (int a = 0; a < dim; ++a) {
        if (lo[a] < 0 || hi[a] > g || hi[a] < lo[a]) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "jittered_lattice_box: box outside the lattice");
        bx.lo[a] = lo[a]; bx.ext[a] = hi[a] - lo[a];
        count *= bx.ext[a];
        bx.blocks[a] = blocks ? blocks[a] : 1;
        if (bx.blocks[a] < 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "jittered_lattice_box: blocks must be >= 1");
    }
    bx.have_blocks = blocks != nullptr;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return RBFFD_OK;
    lattice_box_kernel<<<ceil_div_i64(count, 256), 256, 0, ctx->stream>>>(dim, g, seed, bx, count, X_out, gid_out, owner_out);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

}  // extern "C"
