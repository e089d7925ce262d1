/*
 * rbffd.h -- C ABI of librbffd.so: the B200 (sm_100a) implementation of the operator-generation and
 * operator-application hot path of RadialBasisFiniteDifferences.jl.
 *
 * The reference has no FFI layer; its boundary is the exported Julia API
 * (src/RadialBasisFiniteDifferences.jl:25-75).  Every entry point below names the reference call it
 * replaces.  Julia reaches these with `ccall` (INTEGRATION.md shows the stubs); the Python mirror in
 * radialbasisfinitedifferences.jl_b200/api.py binds exactly the same symbols with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.
 *   - "_host" entry points take HOST buffers (the caller's Julia/NumPy arrays), do the H2D/D2H copies
 *     themselves and block until the result is in the caller's buffers.
 *   - "_device" entry points take DEVICE pointers on the context's device and are asynchronous on the
 *     context's stream (rbffd_set_stream / rbffd_synchronize).
 *   - node coordinates are interleaved FP64 (x0 y0 [z0] x1 y1 ...), i.e. pointer(X) of a
 *     Vector{SVector{d,Float64}} (test/poisson_test.jl:21-22, src/extractcoordinates.jl:11).
 *   - index arrays crossing the host boundary are int64 and use opts->index_base (Julia: 1).
 *     Device index arrays are int32, 0-based.
 *   - every operator row holds exactly n entries (zeros kept; src/generate_operator.jl:171-182), so the
 *     CSR row pointer is implicit: rowptr[k] = k*n.  All operators of one call share `colind`.
 *   - return value: 0 = ok, otherwise an rbffd_status; rbffd_last_error() gives the message.
 *     There is no CPU fallback: without a CUDA device every compute call returns RBFFD_ERR_CUDA.
 */
#ifndef RBFFD_H
#define RBFFD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBFFD_MAX_OPS 12
#define RBFFD_MAX_DIM 3

typedef enum {
    RBFFD_OK = 0,
    RBFFD_ERR_INVALID = 1,    /* bad argument (DimensionMismatch / ArgumentError in the reference)       */
    RBFFD_ERR_K_TOO_LARGE = 2,/* n > number of visible points (ArgumentError from knn)                   */
    RBFFD_ERR_SINGULAR = 3,   /* zero pivot (SingularException from inv, interpolationmatrix.jl:8)        */
    RBFFD_ERR_CUDA = 4,       /* CUDA runtime failure, including "no device"                            */
    RBFFD_ERR_UNSUPPORTED = 5 /* parameter outside the compiled range (e.g. even PHS power)              */
} rbffd_status;

/* operator kinds (ops[i][0]) */
#define RBFFD_OP_DERIV 0    /* d^alpha with alpha = (ops[i][1], ops[i][2], ops[i][3]); weights are multiplied by
                               s^alpha exactly as generate_operator.jl:161-166 / hyperviscosity_operator.jl:159-160 */
#define RBFFD_OP_LAPLACE 1  /* Dxx+Dyy[+Dzz] from ONE right-hand side; extension, not in the reference */

typedef struct {
    int32_t dim;        /* 2 (reference) or 3 (extension)                                               */
    int32_t p;          /* PHS power r^p, odd                         rbfbasis.jl:9                      */
    int32_t polydeg;    /* augmented polynomial degree                polynomialbasis.jl:8               */
    int32_t n;          /* stencil size                               generate_operator.jl:45            */
    int32_t nops;       /* number of operators (right-hand sides) per row, <= RBFFD_MAX_OPS              */
    int32_t ops[RBFFD_MAX_OPS][4];
    int32_t index_base; /* 0 or 1: base of int64 index arrays crossing the host boundary                */
    int32_t sort_columns; /* != 0: entries of every row sorted by column (CSR form of the CSC the reference builds) */
    int32_t kernel;     /* 0 = auto, 1 = generic shared-memory LU kernel, 2 = register/DMMA Gauss-Jordan kernels,
                           3 = null-space kernel (2 and 3: error if not applicable), 4 = as 3, but rows with Y != X are
                           solved one by one (no sharing of the elimination between the rows of a centre) */
    int32_t variant;    /* 0 = scaled two-set methods (generate_operator.jl:29,192; hyperviscosity_operator.jl:26,177)
                           1 = legacy collocated methods generate_operator(X, p, n, polydeg) (generate_operator.jl:354) and
                               hyperviscosity_operator(K, X, p, n, polydeg) (hyperviscosity_operator.jl:314): no scaling,
                               centre node moved to (eps, eps), RBF right-hand sides evaluated at X_j - x_c.  Y must be X. */
    int32_t index_width; /* 0 or 64: host index arrays are int64 (SparseMatrixCSC{Float64,Int64}); 32: rbffd_generate_operator_host
                            writes int32 indices into colind_out (a SparseMatrixCSC{Float64,Int32}: half the PCIe bytes, no widening pass) */
    int32_t reserved[3];
} rbffd_options;

typedef struct rbffd_context rbffd_context;     /* one per (device, stream); not thread-safe, thread-compatible */
typedef struct rbffd_operator rbffd_operator;   /* device-resident operator set sharing one sparsity pattern   */

/* ---- context ------------------------------------------------------------------------------------------ */
int rbffd_create(int device, rbffd_context** ctx);
int rbffd_destroy(rbffd_context* ctx);
const char* rbffd_last_error(const rbffd_context* ctx);    /* valid until the next call on ctx; ctx may be NULL */
int rbffd_set_stream(rbffd_context* ctx, void* cuda_stream); /* cudaStream_t owned by the caller (e.g. torch)   */
int rbffd_get_stream(rbffd_context* ctx, void** cuda_stream);/* the stream the context launches on right now     */
int rbffd_reset_stream(rbffd_context* ctx);                  /* back to the context's own (non-blocking) stream   */
int rbffd_synchronize(rbffd_context* ctx);
int rbffd_version(void);
/* milliseconds (CUDA events) of the phases of the last generate call: [0] binning [1] knn [2] nearest
 * [3] weights [4] column sort/copies ; n <= 8 */
int rbffd_timings(rbffd_context* ctx, double* ms, int n);
/* number of hand-written kernels launched through ctx so far (CUB sort passes are not counted) */
long long rbffd_launch_count(const rbffd_context* ctx);
/* FP64 roofline denominator measured live: dependent-chain DFMA and DMMA (mma.sync.m8n8k4.f64) kernels, TFLOP/s */
int rbffd_measure_fp64_peak(rbffd_context* ctx, double* dfma_tflops, double* dmma_tflops);

/* ---- neighbour search --------------------------------------------------------------------------------- */
/* Replaces KDTree(X); knn(tree, Q, k, true)  (generate_operator.jl:43-47) and, with groups, the masked
 * search of calculateneighbors (calculateneighbors.jl:16-42,83-87).  group code per node: 0 interior,
 * 1+2b boundary b, 2+2b ghost set b.  Results ascending by (squared distance, index).
 * xgroup/qgroup/d2_out may be NULL. */
int rbffd_knn_device(rbffd_context* ctx, const double* X, int64_t N, int32_t dim,
                     const double* Q, int64_t NQ, int32_t k,
                     const int32_t* xgroup, const int32_t* qgroup,
                     int32_t* idx_out /* NQ*k */, double* d2_out /* NQ*k or NULL */);

/* calculateneighbors(X, Y, n, X_idx_in, X_idx_bc, X_idx_bc_g, ...)  (calculateneighbors.jl:1-97).
 * xgroup NULL = unmasked (the inline search of generate_operator.jl:43-47).
 * dist_* receive Euclidean distances (like NearestNeighbors); any output may be NULL. */
int rbffd_calculateneighbors_host(rbffd_context* ctx, const double* X, int64_t N, const double* Y, int64_t M,
                                  int32_t dim, int32_t n, const int32_t* xgroup, int32_t index_base,
                                  int64_t* idxs_x /* N*n */, int64_t* idxs_y_x /* M */,
                                  double* dists_x /* N*n */, double* dists_y_x /* M */);

/* ---- operator generation ------------------------------------------------------------------------------ */
/* generate_operator(X, Y, p, n, polydeg[, index sets])  (generate_operator.jl:29, :192) and
 * hyperviscosity_operator(K, X, Y, p, n, polydeg[, index sets]) (hyperviscosity_operator.jl:26, :177):
 * one neighbour search and ONE factorisation per X node serve every operator in opts->ops.
 * Y == NULL means Y = X.  colind_out[M*n] and vals_out[nops*M*n] are caller-allocated host buffers. */
/* colind_out: int64 [M*n], or int32 [M*n] passed through the same pointer when opts->index_width == 32 */
int rbffd_generate_operator_host(rbffd_context* ctx, const rbffd_options* opts,
                                 const double* X, int64_t N, const double* Y, int64_t M,
                                 const int32_t* xgroup, int64_t* colind_out, double* vals_out);

/* same, everything device resident: stencils [NS*n] int32 from rbffd_knn_device (or the caller; entries index
 * X, entry 0 of a stencil is its centre node; NS <= 0 means NS = N), center[M] int32 = stencil used by row k
 * (NULL: row k uses stencil k, needs M == NS -- the sharded case: NS owned nodes, N = owned + halo).
 * colind_out int32 [M*n], vals_out [nops*M*n]. */
int rbffd_weights_device(rbffd_context* ctx, const rbffd_options* opts,
                         const double* X, int64_t N, const double* Y, int64_t M,
                         const int32_t* stencils, int64_t NS, const int32_t* center,
                         int32_t* colind_out, double* vals_out);

/* stencils of every X node (n nearest, masked by xgroup) and the nearest X node of every Y row in one binning
 * pass -- the two searches of generate_operator.jl:45-47, device resident.  Either output may be NULL. */
int rbffd_stencils_device(rbffd_context* ctx, const double* X, int64_t N, int32_t dim, const double* Y, int64_t M,
                          int32_t n, const int32_t* xgroup, int32_t* stencils_out /* N*n */, int32_t* center_out /* M */);

/* kNN + nearest + weights, device resident; returns an operator handle (matrices never leave HBM). */
int rbffd_operator_generate(rbffd_context* ctx, const rbffd_options* opts,
                            const double* X_dev, int64_t N, const double* Y_dev, int64_t M,
                            const int32_t* xgroup_dev, rbffd_operator** op);
/* the same from HOST coordinates (a host language without its own device arrays, e.g. Julia without CUDA.jl): uploads X / Y /
 * the group codes, generates on the device and returns the handle; nothing of size nnz crosses PCIe */
int rbffd_operator_generate_host(rbffd_context* ctx, const rbffd_options* opts, const double* X, int64_t N, const double* Y, int64_t M,
                                 const int32_t* xgroup, rbffd_operator** op);
/* plain device buffers for such callers (field vectors of the time loop): cudaMalloc / cudaFree / synchronous copies on the
 * context's stream */
int rbffd_device_malloc(rbffd_context* ctx, int64_t bytes, void** ptr);
int rbffd_device_free(rbffd_context* ctx, void* ptr);
int rbffd_device_upload(rbffd_context* ctx, void* dst_device, const void* src_host, int64_t bytes);
int rbffd_device_download(rbffd_context* ctx, void* dst_host, const void* src_device, int64_t bytes);
/* wrap host CSR data (fixed row length) */
int rbffd_operator_from_host(rbffd_context* ctx, int64_t M, int64_t N, int32_t n, int32_t nmat,
                             const int64_t* colind, int32_t index_base, const double* vals, rbffd_operator** op);
/* non-owning view of caller-managed device arrays (colind int32 [M*n] 0-based, vals [nmat][M*n]) */
int rbffd_operator_from_device(rbffd_context* ctx, int64_t M, int64_t N, int32_t n, int32_t nmat,
                               const int32_t* colind, const double* vals, rbffd_operator** op);
int rbffd_operator_destroy(rbffd_operator* op);
int rbffd_operator_info(const rbffd_operator* op, int64_t* M, int64_t* N, int32_t* n, int32_t* nmat);
/* device pointers of the shared pattern and of matrix `which` (for callers that manage their own kernels) */
int rbffd_operator_pointers(const rbffd_operator* op, int32_t which, const int32_t** colind, const double** vals);
int rbffd_operator_to_host(rbffd_operator* op, int32_t index_base, int64_t* colind_out, double* vals_out);

/* ---- operator application ----------------------------------------------------------------------------- */
/* y = alpha * D[which] * x + beta * y       (D*u, examples/adv_diff_test.jl:151-152); device pointers */
int rbffd_spmv_device(rbffd_operator* op, int32_t which, double alpha, const double* x, double beta, double* y);
/* y = alpha * D[which]' * v + beta * y      (E' * v, adv_diff_test.jl:151); y has N entries */
int rbffd_spmv_t_device(rbffd_operator* op, int32_t which, double alpha, const double* v, double beta, double* y);
/* y = sum_i coef[i] * D[which[i]] * x  in ONE pass over the shared pattern (fused multi-operator SpMV) */
int rbffd_spmv_multi_device(rbffd_operator* op, int32_t nterms, const int32_t* which, const double* coef,
                            const double* x, double* y);
/* vals_out[e] = sum_i coef[i] * D[which[i]].vals[e]  over the M*n entries of the shared pattern: the scalar-times-sparse
 * and sparse-plus-sparse products of `alpha*D_xx + alpha*D_yy - u_x*D_x - u_y*D_y` (adv_diff_test.jl:151-152) done ONCE
 * when the coefficients are constant, so that every right-hand-side evaluation is a single-matrix SpMV (12n + 16 B/row).
 * vals_out: caller-owned device array of M*n doubles; wrap it with rbffd_operator_from_device(…, colind of op, vals_out). */
int rbffd_operator_combine_device(rbffd_operator* op, int32_t nterms, const int32_t* which, const double* coef, double* vals_out);
int rbffd_spmv_host(rbffd_operator* op, int32_t which, double alpha, const double* x, double beta, double* y);
int rbffd_spmv_t_host(rbffd_operator* op, int32_t which, double alpha, const double* v, double beta, double* y);

/* du = E' * (alpha*Dxx*u + alpha*Dyy*u - ux*Dx*u - uy*Dy*u) - gamma * (Dxk + Dyk) * u
 * = the interior line of cons_sys (adv_diff_test.jl:151-152).  iE..iDyk index the operator's matrices;
 * iDxk/iDyk < 0 drops the hyperviscosity term.  u, du: device, length N (= M). */
typedef struct {
    int32_t iE, iDx, iDy, iDxx, iDyy, iDxk, iDyk;
    int32_t flags;      /* RBFFD_ADVDIFF_* */
    double alpha, ux, uy, gamma;
} rbffd_advdiff_params;
/* rows are collocated with the nodes (Y == X, as in adv_diff_test.jl): E is then the identity, and what the weight solve returns
 * for it deviates from I by rounding noise only (eps * cond(A_i)).  When max|E - I| <= 1e-8 (checked once per operator, outside
 * CUDA-graph capture) E' * w is taken as w and the whole line becomes ONE pass over the shared pattern with up to six value
 * planes; the result differs from the three-product form by that noise (|E - I| relative), and is the more accurate of the two. */
#define RBFFD_ADVDIFF_COLLOCATED 1
int rbffd_rhs_advdiff_device(rbffd_operator* op, const rbffd_advdiff_params* prm, const double* u, double* du);
/* out = a*u + b*(x + dt * cons_sys(x)): one Shu-Osher stage of an SSP-RK scheme (the reference hands cons_sys to
 * OrdinaryDiffEq's SSPRK43, adv_diff_test.jl:196-199) with the right-hand side of rbffd_rhs_advdiff_device; ONE launch on the
 * collocated path.  out must not alias x or u. */
int rbffd_rhs_advdiff_stage_device(rbffd_operator* op, const rbffd_advdiff_params* prm, const double* x, double a, const double* u,
                                   double b, double dt, double* out);
/* out[i] = a*u[i] + b*(x[i] + dt*du[i]), i < N: the stage combination alone, for right-hand sides that mutate their argument
 * between the product and the combination (cons_sys updates the ghost nodes of u AFTER du is formed, adv_diff_test.jl:151-176,
 * and the integrator then combines the updated u with du).  Element-wise: out may alias u or x. */
int rbffd_stage_update_device(rbffd_context* ctx, int64_t N, double a, const double* u, double b, const double* x, double dt,
                              const double* du, double* out);
/* out = a*u + b*(x + dt * sum_i coef[i] * D[which[i]] * x), nterms <= 6, rows = nodes (M == N): the same stage for any
 * linear semidiscretisation over the shared pattern, ONE launch */
int rbffd_spmv_stage_device(rbffd_operator* op, int32_t nterms, const int32_t* which, const double* coef, const double* x,
                            double a, const double* u, double b, double dt, double* out);
int rbffd_rhs_advdiff_host(rbffd_operator* op, const rbffd_advdiff_params* prm, const double* u, double* du);

/* ---- ghost-node boundary updates of cons_sys (examples/adv_diff_test.jl:118-141 set-up, :162-176 per call) ---- */
/* nb boundaries, applied in the given order.  kind[b] = 1: u[ghost_b] = -inv(D[bc_b, ghost_b]) * (D[bc_b, non-ghost_b] * u)
 * with D = matrix which[b] of `op` (Dx for left/right, Dy for top/bottom in the example); kind[b] = 0:
 * u[bc_b] = u[ghost_b] = value[b] (:166-167).  ptr[nb+1] delimits boundary b's entries in bc_idx / ghost_idx
 * (|ghost_b| == |bc_b|, src/processmesh.jl:174).  Host int64 indices with index_base; u is a device / host vector. */
typedef struct rbffd_bc rbffd_bc;
int rbffd_bc_create(rbffd_operator* op, int32_t nb, const int32_t* kind, const int32_t* which, const double* value,
                    const int64_t* ptr, const int64_t* bc_idx, const int64_t* ghost_idx, int32_t index_base, rbffd_bc** bc);
int rbffd_bc_apply_device(rbffd_bc* bc, double* u);
int rbffd_bc_apply_host(rbffd_bc* bc, double* u);
int rbffd_bc_destroy(rbffd_bc* bc);

/* rows [row0, row1) only, reading x through `xmap` is not needed: sharded operators store LOCAL column
 * ids into [owned | halo]; pack/unpack kernels move halo values (multi-GPU SpMV). */
int rbffd_gather_device(rbffd_context* ctx, const double* src, const int32_t* index, int64_t count, double* dst);
int rbffd_scatter_add_device(rbffd_context* ctx, const double* src, const int32_t* index, int64_t count, double* dst);

/* ---- NVLink peer-memory halo exchange (one process per GPU) ------------------------------------ */
/* cudaMalloc + CUDA IPC export / import: the field u = [halo_lo | owned | halo_hi] and its 4 flag words live in such a
 * buffer so that the neighbouring ranks can store into it directly.  handle64: 64 bytes (cudaIpcMemHandle_t). */
int rbffd_ipc_alloc(rbffd_context* ctx, int64_t bytes, void** ptr, unsigned char* handle64);
int rbffd_ipc_free(rbffd_context* ctx, void* ptr);
int rbffd_ipc_open(rbffd_context* ctx, const unsigned char* handle64, void** peer_ptr);
int rbffd_ipc_close(rbffd_context* ctx, void* peer_ptr);
typedef struct {
    double* u;                 /* my field: n_lo + n_owned + n_hi doubles                                          */
    int64_t n_lo, n_owned, n_hi;
    void* flags;               /* my 4 x uint32 flag block (in the same IPC buffer, zero-initialised)               */
    double* peer_lo_u;         /* lower neighbour's field (mapped peer memory) or NULL                              */
    void* peer_lo_flags;
    int64_t peer_lo_offset;    /* where my first rows go in ITS field (= its n_lo + its n_owned)                    */
    int64_t count_to_lo;       /* = its n_hi                                                                        */
    double* peer_hi_u;
    void* peer_hi_flags;
    int64_t peer_hi_offset;    /* = 0 (its halo_lo)                                                                 */
    int64_t count_to_hi;       /* = its n_lo                                                                        */
} rbffd_halo;
/* epoch = 1, 2, 3, ... (one per exchange, same on every rank).  Call order per application of the operator:
 * push -> interior rows -> wait -> boundary rows -> ack.  All three are asynchronous kernels on the context's stream. */
int rbffd_halo_push_device(rbffd_context* ctx, const rbffd_halo* h, uint32_t epoch);
int rbffd_halo_wait_device(rbffd_context* ctx, const rbffd_halo* h, uint32_t epoch);
int rbffd_halo_ack_device(rbffd_context* ctx, const rbffd_halo* h, uint32_t epoch);

/* ---- spatial-block sharding of arbitrary node sets (BASELINE.json north_star: "nodes are partitioned by spatial blocks
 *      across the 8 GPUs of one box").  The reference has no distributed code (src/domains/domains.jl:7-8).
 *      One process per GPU.  Generation needs no communication: every rank searches its own nodes among
 *      [owned | candidate halo] coordinates.  Application needs ONE halo exchange per product, fused into the SpMV launch:
 *      peers store their values straight into this rank's inbox over NVLink (CUDA IPC), boundary rows wait for the flags.
 *      Local numbering of a shard: [interior owned rows | boundary owned rows | halo]; "boundary" = references a halo column.
 *      All vectors handed to rbffd_shard_spmv* hold the n_owned owned values in that local order. ---- */
#define RBFFD_ERR_HALO 6          /* candidate halo too thin: a stencil could reach past the nodes held locally */
#define RBFFD_SHARD_MAX_PEERS 16
typedef struct rbffd_shard rbffd_shard;
/* Partition N nodes into nparts spatial blocks: a b0 x b1 [x b2] grid whose cut planes are coordinate quantiles (every block of
 * one slab holds the same number of nodes +-1).  blocks = NULL or zeros: the factorisation with the smallest cut surface.
 * part_out[i] in [0, nparts).  Host only (no device needed). */
int rbffd_shard_plan_host(const double* X, int64_t N, int32_t dim, int32_t nparts, const int32_t* blocks, int32_t* part_out);
/* Build the shard of `rank` from the full node set and the partition (host arrays): owned = part == rank; candidate halo =
 * foreign nodes inside the owned bounding box inflated by a margin (grown until the exactness proof below holds). */
int rbffd_shard_create_host(rbffd_context* ctx, const double* X, int64_t N, int32_t dim, const int32_t* part, int32_t nparts,
                            int32_t rank, int32_t n, rbffd_shard** shard);
/* Same from device-resident candidates: Xc [nc][dim], gid [nc] (ASCENDING global ids: ties of the neighbour search are then
 * broken by global id, as on one GPU), owner [nc].  box_lo/hi[dim] delimit the region all of whose nodes are among the
 * candidates (+-infinity where nothing lies beyond).  Exactness proof: every owned stencil's radius is smaller than the
 * distance of its centre to the faces of that box; otherwise RBFFD_ERR_HALO. */
int rbffd_shard_create_device(rbffd_context* ctx, int32_t dim, int32_t n, const double* Xc, const int64_t* gid, const int32_t* owner,
                              int64_t nc, const double* box_lo, const double* box_hi, int32_t rank, int32_t nparts, rbffd_shard** shard);
int rbffd_shard_destroy(rbffd_shard* shard);
int rbffd_shard_info(const rbffd_shard* shard, int64_t* n_owned, int64_t* n_interior, int64_t* n_halo);
/* synchronises the context's stream and returns RBFFD_ERR_HALO when a fused exchange gave up waiting for a peer (the waits inside
 * the kernels time out after 20 s instead of hanging the GPU: crashed rank, mismatched call sequence), RBFFD_OK otherwise */
int rbffd_shard_status(rbffd_shard* shard);
/* local id -> global id, n_owned + n_halo entries */
int rbffd_shard_global_ids_host(rbffd_shard* shard, int32_t index_base, int64_t* gid_out);
/* device arrays of the shard: coordinates [n_owned + n_halo][dim] and stencils [n_owned][n] (local ids, entry 0 = the row's node):
 * feed them to rbffd_weights_device(ctx, opts, X_local, n_owned + n_halo, X_local, n_owned, stencils, n_owned, NULL, ...). */
int rbffd_shard_device_arrays(const rbffd_shard* shard, const double** X_local, const int32_t** stencils);
/* halo nodes this rank needs from `peer` (count, then global ids in halo order); the caller ships the lists to the peers
 * (torch.distributed / MPI all-to-all) and hands every rank the lists addressed to it: */
int rbffd_shard_recv_count(const rbffd_shard* shard, int32_t peer, int64_t* count);
int rbffd_shard_recv_ids_host(rbffd_shard* shard, int32_t peer, int32_t index_base, int64_t* ids_out);
int rbffd_shard_set_send_ids_host(rbffd_shard* shard, int32_t peer, int32_t index_base, const int64_t* ids, int64_t count);
/* after all send lists are set: allocates the IPC inbox [halo values | flags | reverse inbox]; handle64 = cudaIpcMemHandle_t.
 * offsets[2 * nparts + 1]: for every peer p, offsets[2p] = where ITS values land in my inbox (doubles), offsets[2p + 1] = where
 * its transposed-product contributions land in my buffer (doubles; -1: none); offsets[2 * nparts] = byte offset of my flag
 * block.  The caller all-gathers handle + offsets (torch.distributed / MPI) and connects every peer it exchanges with: */
int rbffd_shard_finalize(rbffd_shard* shard, unsigned char* handle64, int64_t* offsets);
/* fwd_offset = the peer's offsets[2 * me], rev_offset = the peer's offsets[2 * me + 1], flags_offset = the peer's offsets[2 * nparts] */
int rbffd_shard_connect(rbffd_shard* shard, int32_t peer, const unsigned char* handle64, int64_t fwd_offset, int64_t rev_offset,
                        int64_t flags_offset);
/* y[0:n_owned] = sum_i coef[i] * D[which[i]] * [x ; halo of x]   (nterms <= 6), `op` = operators over the shard's local pattern
 * (rows n_owned, columns n_owned + n_halo).  ONE kernel launch: the first CTAs store x's boundary values into the peers'
 * inboxes and publish the epoch, interior rows run meanwhile, the CTAs of the boundary rows wait for the peers' flags, the
 * last of them acknowledges.  The epoch lives in device memory: the call can be captured into a CUDA graph.
 * Every rank must make the same sequence of rbffd_shard_spmv* calls. */
int rbffd_shard_spmv_device(rbffd_shard* shard, rbffd_operator* op, int32_t nterms, const int32_t* which, const double* coef,
                            const double* x, double* y);
/* y[0:n_owned] = alpha * D[which]' * v + beta * y : the transposed product E' * v of adv_diff_test.jl:151 on a sharded operator.
 * Contributions to halo columns travel back to their owners (reverse exchange) and are added in rank order (deterministic). */
int rbffd_shard_spmv_t_device(rbffd_shard* shard, rbffd_operator* op, int32_t which, double alpha, const double* v, double beta, double* y);
/* out[0:n_owned] = a*u + b*(x + dt * sum_i coef[i] * D[which[i]] * [x ; halo of x]): one SSP-RK stage of the sharded
 * semidiscretisation as ONE launch (halo exchange + product + stage update); out must not alias x or u */
int rbffd_shard_spmv_stage_device(rbffd_shard* shard, rbffd_operator* op, int32_t nterms, const int32_t* which, const double* coef,
                                  const double* x, double a, const double* u, double b, double dt, double* out);
/* transport-agnostic form of the two exchanges (single-process tests, NCCL / gloo / MPI transports): pack = the values of x
 * this rank owes `peer` (count = what the peer asked for), unpack = fill the halo segment owned by `peer` from its packed values;
 * rbffd_shard_spmv_local_device is the same product as rbffd_shard_spmv_device WITHOUT any exchange (inbox filled by unpack). */
int rbffd_shard_pack_device(rbffd_shard* shard, int32_t peer, const double* x, double* sendbuf);
int rbffd_shard_unpack_device(rbffd_shard* shard, int32_t peer, const double* recvbuf);
int rbffd_shard_spmv_local_device(rbffd_shard* shard, rbffd_operator* op, int32_t nterms, const int32_t* which, const double* coef,
                                  const double* x, double* y);
/* transposed product in three steps: local part (y = alpha D' v restricted to owned columns + beta y; the halo-column sums stay
 * inside the shard), tpack = the sums owed to `peer` (recv_count(peer) doubles, contiguous), tunpack_add = add `peer`'s sums
 * to the nodes it had asked for (send order).  Call tunpack_add in ascending peer order to reproduce rbffd_shard_spmv_t_device. */
int rbffd_shard_spmv_t_local_device(rbffd_shard* shard, rbffd_operator* op, int32_t which, double alpha, const double* v, double beta, double* y);
int rbffd_shard_tpack_device(rbffd_shard* shard, int32_t peer, double* sendbuf);
int rbffd_shard_tunpack_add_device(rbffd_shard* shard, int32_t peer, const double* recvbuf, double* y);
/* number of values this rank sends to `peer` per product (= what the peer's recv list asked for) */
int rbffd_shard_send_count(const rbffd_shard* shard, int32_t peer, int64_t* count);

/* ---- synthetic node sets: jittered lattice in [0,1]^d, counter-based RNG, on device ------ */
/* node (i,j[,l]) of a g^d lattice, linear ids [first, first+count): ((i,j,l)+0.5+0.5*(U-0.5))/g */
int rbffd_jittered_lattice_device(rbffd_context* ctx, int32_t dim, int64_t g, uint64_t seed,
                                  int64_t first, int64_t count, double* X_out);
/* the nodes of the lattice box [lo, hi) (lattice coordinates per axis), enumerated in ascending global linear id; also writes
 * the ids and, when blocks != NULL, the owner of every node under the regular blocks[dim] block partition of the lattice
 * (block index along axis a = (c_a * blocks[a]) / g, owner = (b0 * blocks[1] + b1) * blocks[2] + b2 with x slowest) */
int rbffd_jittered_lattice_box_device(rbffd_context* ctx, int32_t dim, int64_t g, uint64_t seed, const int64_t* lo, const int64_t* hi,
                                      const int32_t* blocks, double* X_out, int64_t* gid_out, int32_t* owner_out);

#ifdef __cplusplus
}
#endif
#endif /* RBFFD_H */
