// weights.cu -- K4 of SURVEY.md §2: the fused RBF-FD weight kernel.
//
// One launch replaces, for every X node i and every Y row k whose nearest X node is i:
//   scalestencil            (src/scalestencil.jl:10-20)
//   rbfblock/polynomialblock/interpolationmatrix   (src/rbfblock.jl:14-20, src/polynomialblock.jl:26-31,
//                                                   src/interpolationmatrix.jl:5-8)
//   the RHS assembly, M_inv*RHS and the rescaling  (src/generate_operator.jl:89-167,
//                                                   src/hyperviscosity_operator.jl:97-161,
//                                                   src/polylinearoperator.jl:36-44,63-67)
//   the COO->sparse assembly                       (src/generate_operator.jl:171-182): rows are written
//                                                   straight into fixed-row-length CSR.
// The reference materialises inv(A) for every node (2 m^2 doubles per node); here A is factorised
// (partial-pivot LU, FP64) and solved for all right-hand sides, nothing of size m^2 reaches HBM.
//
// This file holds the host driver, the operator tables and the GENERIC kernel (any m that fits shared
// memory, any number of Y rows per centre).  The register/DMMA fast path lives in weights_fast.cu.
#include <cub/device/device_radix_sort.cuh>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "tables.cuh"

// ---------------------------------------------------------------------------------------------------
// host: operator tables
// ---------------------------------------------------------------------------------------------------
namespace {

struct HTerm { double coef; int e[3]; int rpow; };

int derivative_terms(int p, int d, const int* alpha, HTerm* out, int cap) {
    HTerm cur[64], nxt[64];
    int nc = 1;
    cur[0] = HTerm{1.0, {0, 0, 0}, p};
    for (int a = 0; a < d; ++a)
        for (int rep = 0; rep < alpha[a]; ++rep) {
            int nn = 0;
            for (int t = 0; t < nc; ++t) {
                HTerm cand[2];
                int ncand = 0;
                if (cur[t].e[a] > 0) { cand[ncand] = cur[t]; cand[ncand].coef *= cur[t].e[a]; cand[ncand].e[a] -= 1; ncand++; }
                if (cur[t].rpow != 0) { cand[ncand] = cur[t]; cand[ncand].coef *= cur[t].rpow; cand[ncand].e[a] += 1; cand[ncand].rpow -= 2; ncand++; }
                for (int c = 0; c < ncand; ++c) {
                    int found = -1;
                    for (int u = 0; u < nn; ++u)
                        if (nxt[u].e[0] == cand[c].e[0] && nxt[u].e[1] == cand[c].e[1] && nxt[u].e[2] == cand[c].e[2] &&
                            nxt[u].rpow == cand[c].rpow) { found = u; break; }
                    if (found >= 0) nxt[found].coef += cand[c].coef;
                    else if (nn < 64) nxt[nn++] = cand[c];
                }
            }
            nc = 0;
            for (int u = 0; u < nn; ++u) if (nxt[u].coef != 0.0) cur[nc++] = nxt[u];
        }
    if (nc > cap) return -1;
    for (int t = 0; t < nc; ++t) out[t] = cur[t];
    return nc;
}

}  // namespace

int build_op_tables(const rbffd_options* o, OpTables* T, char* err, int errlen) {
    memset(T, 0, sizeof(*T));
    const int d = o->dim;
    if (d < 1 || d > 3) { snprintf(err, errlen, "dim must be 1..3 (got %d)", d); return RBFFD_ERR_INVALID; }
    if (o->p < 1 || (o->p & 1) == 0) { snprintf(err, errlen, "PHS power p must be a positive odd integer (got %d)", o->p); return RBFFD_ERR_UNSUPPORTED; }
    if (o->polydeg < 0) { snprintf(err, errlen, "polydeg must be >= 0"); return RBFFD_ERR_INVALID; }
    if (o->nops < 1 || o->nops > RBFFD_MAX_OPS) { snprintf(err, errlen, "nops must be 1..%d", RBFFD_MAX_OPS); return RBFFD_ERR_INVALID; }
    int64_t q = 1;
    for (int t = 1; t <= d; ++t) q = q * (o->polydeg + t) / t;
    if (q > TAB_MAX_MONO) { snprintf(err, errlen, "polydeg %d gives %lld monomials (max %d)", o->polydeg, (long long)q, TAB_MAX_MONO); return RBFFD_ERR_UNSUPPORTED; }
    if (o->n < 1) { snprintf(err, errlen, "stencil size n must be >= 1"); return RBFFD_ERR_INVALID; }
    if (o->n < q) { snprintf(err, errlen, "stencil size n=%d is smaller than the %lld polynomial terms: A is singular", o->n, (long long)q); return RBFFD_ERR_SINGULAR; }
    T->dim = d; T->p = o->p; T->n = o->n; T->q = (int)q; T->m = o->n + (int)q; T->nops = o->nops;
    // graded exponent table (column order of P does not affect the first n solution entries)
    int qi = 0;
    for (int g = 0; g <= o->polydeg; ++g)
        for (int a = g; a >= 0; --a) {
            if (d == 1) { if (a == g) { T->mono[qi][0] = (int8_t)a; qi++; } continue; }
            for (int b = g - a; b >= 0; --b) {
                int c = g - a - b;
                if (d == 2 && c != 0) continue;
                T->mono[qi][0] = (int8_t)a; T->mono[qi][1] = (int8_t)b; T->mono[qi][2] = (int8_t)c; qi++;
            }
        }
    for (int tq = 1; tq < qi; ++tq) {
        int ax = d - 1;
        while (ax > 0 && T->mono[tq][ax] == 0) --ax;
        int8_t pe[3] = {T->mono[tq][0], T->mono[tq][1], T->mono[tq][2]};
        pe[ax] -= 1;
        int par = 0;
        for (int u = 0; u < tq; ++u)
            if (T->mono[u][0] == pe[0] && T->mono[u][1] == pe[1] && T->mono[u][2] == pe[2]) { par = u; break; }
        T->mpar[tq] = (int8_t)par;
        T->maxis[tq] = (int8_t)ax;
    }
    int nt = 0;
    for (int i = 0; i < o->nops; ++i) {
        T->kind[i] = o->ops[i][0];
        HTerm tmp[64];
        if (o->ops[i][0] == RBFFD_OP_DERIV) {
            int al[3] = {o->ops[i][1], o->ops[i][2], o->ops[i][3]};
            int tot = 0;
            for (int a = 0; a < 3; ++a) {
                if (al[a] < 0 || (a >= d && al[a] != 0)) { snprintf(err, errlen, "operator %d: bad derivative multi-index", i); return RBFFD_ERR_INVALID; }
                tot += al[a];
                T->alpha[i][a] = (int8_t)al[a];
            }
            // derivatives of r^p of total order >= p are singular at r = 0 (finite only if K < p)
            if (tot >= o->p + 1) { snprintf(err, errlen, "operator %d: derivative order %d too high for r^%d", i, tot, o->p); return RBFFD_ERR_UNSUPPORTED; }
            if (tot > 100) return RBFFD_ERR_INVALID;
            int c = derivative_terms(o->p, d, al, tmp, 64);
            if (c < 0 || nt + c > TAB_MAX_TERMS) { snprintf(err, errlen, "operator table overflow"); return RBFFD_ERR_UNSUPPORTED; }
            T->tb[3 * i] = (int16_t)nt;
            for (int t = 0; t < c; ++t) {
                T->coef[nt] = tmp[t].coef;
                T->te[nt][0] = (int8_t)tmp[t].e[0]; T->te[nt][1] = (int8_t)tmp[t].e[1]; T->te[nt][2] = (int8_t)tmp[t].e[2];
                T->te[nt][3] = (int8_t)tmp[t].rpow;
                nt++;
            }
            T->tb[3 * i + 1] = T->tb[3 * i + 2] = (int16_t)nt;
        } else if (o->ops[i][0] == RBFFD_OP_LAPLACE) {
            for (int a = 0; a < 3; ++a) {
                T->tb[3 * i + a] = (int16_t)nt;
                if (a >= d) continue;
                int al[3] = {0, 0, 0};
                al[a] = 2;
                int c = derivative_terms(o->p, d, al, tmp, 64);
                if (c < 0 || nt + c > TAB_MAX_TERMS) { snprintf(err, errlen, "operator table overflow"); return RBFFD_ERR_UNSUPPORTED; }
                for (int t = 0; t < c; ++t) {
                    T->coef[nt] = tmp[t].coef;
                    T->te[nt][0] = (int8_t)tmp[t].e[0]; T->te[nt][1] = (int8_t)tmp[t].e[1]; T->te[nt][2] = (int8_t)tmp[t].e[2];
                    T->te[nt][3] = (int8_t)tmp[t].rpow;
                    nt++;
                }
            }
        } else {
            snprintf(err, errlen, "operator %d: unknown kind %d", i, o->ops[i][0]);
            return RBFFD_ERR_INVALID;
        }
        T->tb[3 * i + 3] = (int16_t)nt;
    }
    return RBFFD_OK;
}

int rbffd_validate_options(rbffd_context* ctx, const rbffd_options* o) {
    if (!o) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "options pointer is NULL");
    OpTables T;
    char msg[256];
    int rc = build_op_tables(o, &T, msg, sizeof(msg));
    if (rc != RBFFD_OK) RBFFD_FAIL(ctx, rc, "%s", msg);
    if (o->index_base != 0 && o->index_base != 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "index_base must be 0 or 1");
    if (o->variant != 0 && o->variant != 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "variant must be 0 (two-set methods) or 1 (legacy collocated methods)");
    if (o->index_width != 0 && o->index_width != 32 && o->index_width != 64) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "index_width must be 0, 32 or 64");
    return RBFFD_OK;
}

// ---------------------------------------------------------------------------------------------------
// device: generic shared-memory kernel (one warp per X node)
// ---------------------------------------------------------------------------------------------------
namespace {

struct WArgs {
    const double* X;
    const double* Y;
    const int32_t* stencils;   // [N][n]
    const int32_t* rows;       // [M] row ids grouped by centre
    const int32_t* seg;        // [N+1] segment starts into rows
    int64_t N, M;
    int32_t* colind;           // [M][n]
    double* vals;              // [nops][M][n]
    int* fail;                 // first singular node + 1 (atomicMin over 0x7fffffff)
    int variant;               // 1: legacy collocated methods (generate_operator.jl:354-491, hyperviscosity_operator.jl:314-440)
    OpTables T;
};

__global__ void seg_start_kernel(const int* __restrict__ key_sorted, int64_t n, int nseg, int* __restrict__ start) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i > n) return;
    int prev = i == 0 ? -1 : key_sorted[i - 1];
    int cur = i == n ? nseg : key_sorted[i];
    for (int c = prev + 1; c <= cur; ++c) start[c] = (int)i;
}

__global__ void iota_kernel(int* p, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int)i;
}

// All index checks of one call in ONE launch: stencil ids in [0, nx) (16-byte loads when the array is aligned), centre ids in
// [0, nctr), and whether centre is the identity.  flags[1]: centre != identity, flags[2]: an index out of range.
__global__ void validate_indices_kernel(const int* __restrict__ st, int64_t ns, int nx, const int* __restrict__ center, int64_t m, int nctr,
                                        int want_identity, int* flags) {
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    bool bad = false, moved = false;
    if (st) {
        const bool vec = (reinterpret_cast<uintptr_t>(st) & 15) == 0;
        const int64_t n4 = vec ? ns >> 2 : 0;
        const int4* st4 = reinterpret_cast<const int4*>(st);
        for (int64_t i = tid; i < n4; i += nth) {
            const int4 v = __ldcs(st4 + i);
            bad = bad || (unsigned)v.x >= (unsigned)nx || (unsigned)v.y >= (unsigned)nx || (unsigned)v.z >= (unsigned)nx || (unsigned)v.w >= (unsigned)nx;
        }
        for (int64_t i = 4 * n4 + tid; i < ns; i += nth) bad = bad || (unsigned)st[i] >= (unsigned)nx;
    }
    if (center) {
        for (int64_t i = tid; i < m; i += nth) {
            const int c = center[i];
            bad = bad || (unsigned)c >= (unsigned)nctr;
            moved = moved || c != (int)i;
        }
    }
    if (bad) flags[2] = 1;
    if (want_identity && moved) flags[1] = 1;
}

template <int D>
__global__ void __launch_bounds__(128) weights_generic_kernel(WArgs a, int warps_per_block, int smem_per_warp) {
    extern __shared__ __align__(16) unsigned char wsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const OpTables& T = a.T;
    const int n = T.n, q = T.q, m = T.m, nops = T.nops;
    const int lda = m | 1;
    unsigned char* base = wsm + (size_t)warp * smem_per_warp;
    double* A = reinterpret_cast<double*>(base);           // m * lda
    double* rhs = A + (size_t)m * lda;                     // nops * m
    double* S = rhs + (size_t)nops * m;                    // n * D
    int* piv = reinterpret_cast<int*>(S + (size_t)n * D);  // m
    const double EPS = 2.220446049250313e-16;
    const unsigned FULL = 0xffffffffu;

    for (int64_t i = blockIdx.x * (int64_t)warps_per_block + warp; i < a.N; i += (int64_t)gridDim.x * warps_per_block) {
        const int r0 = a.seg[i], r1 = a.seg[i + 1];
        if (r0 == r1) continue;   // no Y row maps to this centre: nothing to compute
        const int32_t* st = a.stencils + i * n;
        const int c0 = st[0];
        double xc[D], s[D];
#pragma unroll
        for (int c = 0; c < D; ++c) xc[c] = a.X[(int64_t)c0 * D + c];
        // ---- scalestencil.jl:10-20 ----
        double mx[D];
#pragma unroll
        for (int c = 0; c < D; ++c) mx[c] = 0.0;
        for (int j = lane; j < n; j += 32) {
            const int id = st[j];
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double v = a.X[(int64_t)id * D + c] - xc[c];
                S[j * D + c] = v;
                mx[c] = fmax(mx[c], fabs(v));
            }
        }
#pragma unroll
        for (int c = 0; c < D; ++c) {
            for (int o = 16; o > 0; o >>= 1) mx[c] = fmax(mx[c], __shfl_xor_sync(FULL, mx[c], o));
            s[c] = a.variant ? 1.0 : 1.0 / mx[c];                 // legacy method: stencils are not scaled (:403-406)
        }
        __syncwarp();
        for (int j = lane; j < n; j += 32) {
#pragma unroll
            for (int c = 0; c < D; ++c) S[j * D + c] = (a.variant && j == 0) ? EPS : S[j * D + c] * s[c];   // legacy: centre := (eps, eps) (:410)
        }
        __syncwarp();
        // ---- interpolationmatrix.jl:5 : A = [Phi P; P' 0] ----
        for (int e = lane; e < n * n; e += 32) {
            const int ia = e / n, ib = e - ia * n;
            if (ia > ib) continue;
            double v = 0.0;
            if (ia < ib) {
                double r2 = 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) { double t = S[ia * D + c] - S[ib * D + c]; r2 += t * t; }
                v = rpow_i(sqrt(r2), r2, T.p);
            }
            A[ia * lda + ib] = v;
            A[ib * lda + ia] = v;
        }
        for (int e = lane; e < n * q; e += 32) {
            const int j = e / q, t = e - j * q;
            double v = 1.0;
#pragma unroll
            for (int c = 0; c < D; ++c) v *= ipow_u(S[j * D + c], T.mono[t][c]);
            A[j * lda + n + t] = v;
            A[(n + t) * lda + j] = v;
        }
        for (int e = lane; e < q * q; e += 32) A[(n + e / q) * lda + n + (e % q)] = 0.0;
        __syncwarp();
        // ---- partial-pivot LU in shared memory ----
        bool singular = false;
        for (int k = 0; k < m; ++k) {
            double best = -1.0;
            int bi = k;
            for (int r = k + lane; r < m; r += 32) {
                double v = fabs(A[r * lda + k]);
                if (v > best) { best = v; bi = r; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(FULL, best, o);
                int oi = __shfl_xor_sync(FULL, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (!(best > 0.0)) { singular = true; break; }
            if (lane == 0) piv[k] = bi;
            if (bi != k) {
                for (int j = lane; j < m; j += 32) {
                    double t = A[k * lda + j];
                    A[k * lda + j] = A[bi * lda + j];
                    A[bi * lda + j] = t;
                }
            }
            __syncwarp();
            const double rinv = 1.0 / A[k * lda + k];
            for (int r = k + 1 + lane; r < m; r += 32) A[r * lda + k] *= rinv;
            __syncwarp();
            for (int j0 = k + 1; j0 < m; j0 += 32) {
                const int j = j0 + lane;
                const double u = j < m ? A[k * lda + j] : 0.0;
                if (j < m)
                    for (int r = k + 1; r < m; ++r) A[r * lda + j] -= A[r * lda + k] * u;
            }
            __syncwarp();
        }
        if (singular) {
            if (lane == 0) atomicMin(a.fail, (int)i + 1);
            // leave NaNs so a caller that ignores the status cannot mistake the rows for weights
            for (int rr = r0; rr < r1; ++rr) {
                const int64_t row = a.rows[rr];
                for (int j = lane; j < n; j += 32) {
                    a.colind[row * n + j] = st[j];
                    for (int o = 0; o < nops; ++o) a.vals[((int64_t)o * a.M + row) * n + j] = nan("");
                }
            }
            continue;
        }
        // ---- one RHS set per Y row of this centre (generate_operator.jl:89-167) ----
        for (int rr = r0; rr < r1; ++rr) {
            const int64_t row = a.rows[rr];
            double eta[D];
#pragma unroll
            for (int c = 0; c < D; ++c) eta[c] = a.variant ? EPS : (a.Y[row * D + c] - xc[c]) * s[c];   // legacy: polynomial rows at X_shift[1] (:429)
            for (int j = lane; j < n; j += 32) {
                double del[D];
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    double t = eta[c] - S[j * D + c];
                    del[c] = t == 0.0 ? EPS : t;     // generate_operator.jl:127-132
                    if (a.variant) del[c] = S[j * D + c];          // legacy: RBF rows at X_shift[j] = X_j - x_c (:433)
                }
                for (int o = 0; o < nops; ++o) rhs[o * m + j] = rhs_rbf_entry<D>(T, o, del, s);
            }
            for (int t = lane; t < q; t += 32)
                for (int o = 0; o < nops; ++o) rhs[o * m + n + t] = rhs_poly_entry<D>(T, o, t, eta, s);
            __syncwarp();
            // row interchanges (lanes over right-hand sides)
            for (int o = lane; o < nops; o += 32)
                for (int k = 0; k < m; ++k) {
                    const int pk = piv[k];
                    if (pk != k) { double t = rhs[o * m + k]; rhs[o * m + k] = rhs[o * m + pk]; rhs[o * m + pk] = t; }
                }
            __syncwarp();
            // forward substitution, unit lower
            for (int k = 0; k < m - 1; ++k) {
                for (int o = 0; o < nops; ++o) {
                    const double bk = rhs[o * m + k];
                    for (int r = k + 1 + lane; r < m; r += 32) rhs[o * m + r] -= A[r * lda + k] * bk;
                }
                __syncwarp();
            }
            // back substitution
            for (int k = m - 1; k >= 0; --k) {
                const double dinv = 1.0 / A[k * lda + k];
                for (int o = 0; o < nops; ++o) {
                    const double xk = rhs[o * m + k] * dinv;
                    for (int r = lane; r < k; r += 32) rhs[o * m + r] -= A[r * lda + k] * xk;
                    __syncwarp();
                    if (lane == 0) rhs[o * m + k] = xk;
                }
                __syncwarp();
            }
            // ---- rescale + CSR row write (generate_operator.jl:161-182) ----
            for (int o = 0; o < nops; ++o) {
                const double f = op_post_factor<D>(T, o, s);
                for (int j = lane; j < n; j += 32) a.vals[((int64_t)o * a.M + row) * n + j] = f * rhs[o * m + j];
            }
            for (int j = lane; j < n; j += 32) a.colind[row * n + j] = st[j];
            __syncwarp();
        }
    }
}

// one warp per row: sort the row's entries by column index (columns within a row are distinct)
__global__ void sort_rows_kernel(int64_t M, int n, int nops, int32_t* __restrict__ colind, double* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    extern __shared__ int srt_sm[];
    int* cols = srt_sm + (threadIdx.x >> 5) * n;
    int32_t* crow = colind + row * n;
    for (int j = lane; j < n; j += 32) cols[j] = crow[j];
    __syncwarp();
    constexpr int MAXPER = 8;   // n <= 256
    int rank[MAXPER];
    int cnt = 0;
    for (int j = lane; j < n; j += 32, ++cnt) {
        const int c = cols[j];
        int r = 0;
        for (int t = 0; t < n; ++t) r += (cols[t] < c) || (cols[t] == c && t < j);
        rank[cnt] = r;
    }
    __syncwarp();
    cnt = 0;
    for (int j = lane; j < n; j += 32, ++cnt) crow[rank[cnt]] = cols[j];
    for (int o = 0; o < nops; ++o) {
        double* vrow = vals + ((int64_t)o * M + row) * n;
        double v[MAXPER];
        cnt = 0;
        for (int j = lane; j < n; j += 32, ++cnt) v[cnt] = vrow[j];
        __syncwarp();
        cnt = 0;
        for (int j = lane; j < n; j += 32, ++cnt) vrow[rank[cnt]] = v[cnt];
    }
}

}  // namespace

// fast path (weights_fast.cu); returns RBFFD_ERR_UNSUPPORTED when the configuration has no fast kernel
int rbffd_weights_fast(rbffd_context* ctx, const OpTables& T, const double* X, int64_t N, const double* Y, int64_t M,
                       const int32_t* stencils, int32_t* colind_out, double* vals_out, int* fail_flag);
// null-space path (weights_ns.cu): n <= 32, q <= 12; UNSUPPORTED also when a stencil fails its definiteness check
int rbffd_weights_ns(rbffd_context* ctx, const OpTables& T, const double* X, int64_t N, const double* Y, int64_t M,
                     const int32_t* stencils, const int32_t* center, int32_t* colind_out, double* vals_out, int* fail_flag, int64_t NX = 0);
// multi-warp null-space path (weights_nsw.cu): n <= 64, n - q <= 48; UNSUPPORTED also when a stencil fails its checks
int rbffd_weights_nsw(rbffd_context* ctx, const OpTables& T, const double* X, int64_t N, const double* Y, int64_t M,
                      const int32_t* stencils, const int32_t* center, int32_t* colind_out, double* vals_out);
// two-stage null-space path (weights_ns2.cu): same scope as weights_nsw.cu, P reduction in its own one-warp-per-stencil kernel
int rbffd_weights_ns2(rbffd_context* ctx, const OpTables& T, const double* X, int64_t N, const double* Y, int64_t M,
                      const int32_t* stencils, const int32_t* center, int32_t* colind_out, double* vals_out);
// multi-warp register/DMMA path for 48 < m <= 96 (weights_mw.cu)
int rbffd_weights_mw(rbffd_context* ctx, const OpTables& T, const double* X, int64_t N, const double* Y, int64_t M,
                     const int32_t* stencils, int32_t* colind_out, double* vals_out, int* fail_flag);

int rbffd_weights_impl(rbffd_context* ctx, const rbffd_options* opts, const double* X, int64_t NX,
                       const double* Y, int64_t M, const int32_t* stencils, int64_t N, const int32_t* center,
                       int32_t* colind_out, double* vals_out) {
    RBFFD_TRY(rbffd_validate_options(ctx, opts));
    WArgs a;
    char msg[256];
    build_op_tables(opts, &a.T, msg, sizeof(msg));
    const OpTables& T = a.T;
    if (M == 0) return RBFFD_OK;
    if (T.n > NX) RBFFD_FAIL(ctx, RBFFD_ERR_K_TOO_LARGE, "n=%d exceeds the number of points %lld", T.n, (long long)NX);
    if (T.n > 256) RBFFD_FAIL(ctx, RBFFD_ERR_UNSUPPORTED, "stencil size n=%d > 256", T.n);
    cudaStream_t st = ctx->stream;
    if (!Y) Y = X;

    DevBuf<int> flags;   // [0] singular node+1, [1] centre != identity, [2] bad index
    int h_flags[4] = {0x7fffffff, 0, 0, 0};
    const bool deferred = ctx->deferred_flags != nullptr && ctx->trusted_stencils && !center;
    if (deferred) flags.p = ctx->deferred_flags + 8 * ctx->deferred_slot;     // initialised by the caller, never fetched here
    else {
        CUDA_TRY(ctx, flags.alloc(4, st));
        rbffd_set_flags_kernel<<<1, 32, 0, st>>>(flags.p, h_flags[0], h_flags[1], h_flags[2], h_flags[3]);
    }
    struct Unhook { DevBuf<int>& f; bool on; ~Unhook() { if (on) f.p = nullptr; } } unhook{flags, deferred};
    if (!ctx->trusted_stencils || center) {
        const int64_t work = std::max<int64_t>(ctx->trusted_stencils ? 0 : (N * T.n) / 4, center ? M : 0);
        const int vb = (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)ctx->sm_count * 16));
        validate_indices_kernel<<<vb, 256, 0, st>>>(ctx->trusted_stencils ? nullptr : stencils, N * T.n, (int)NX, center, M, (int)N,
                                                    center && M == N ? 1 : 0, flags.p);
        KLAUNCH(ctx);
    }
    bool identity = false;
    if (center) {
        CUDA_TRY(ctx, rbffd_fetch_flags(ctx, flags.p, 4, h_flags));
        identity = (M == N) && h_flags[1] == 0;
    } else {
        if (M != N) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "center == NULL requires M == N");
        if (!ctx->trusted_stencils) CUDA_TRY(ctx, rbffd_fetch_flags(ctx, flags.p, 4, h_flags));
        identity = true;
    }
    if (h_flags[2]) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "stencil or centre index out of range");

    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[4], st));
    int rc = RBFFD_ERR_UNSUPPORTED;
    a.variant = opts->variant;
    if (opts->variant != 0 && opts->kernel > 1) RBFFD_FAIL(ctx, RBFFD_ERR_UNSUPPORTED, "the legacy collocated variant runs on the generic kernel only");
    // Y != X (generate_operator.jl:89-167, several rows per centre): the single-warp null-space kernel shares one elimination
    // between up to three rows of a centre (segmented mode, weights_ns.cu); kernel = 4 and the larger stencils run one row per
    // work item, reading the stencil of the row's centre (the elimination is repeated per row: still ~8x the generic kernel's rate)
    const bool want_ns = opts->kernel == 3 || opts->kernel == 4;
    const bool rowwise = !identity && opts->kernel != 1 && opts->kernel != 2 && opts->variant == 0;
    if (rowwise) {
        rc = rbffd_weights_ns(ctx, T, X, M, Y, M, stencils, center, colind_out, vals_out, flags.p, opts->kernel == 4 ? 0 : N);
        if (rc == RBFFD_ERR_UNSUPPORTED && T.n > 32) rc = rbffd_weights_ns2(ctx, T, X, M, Y, M, stencils, center, colind_out, vals_out);
        if (rc == RBFFD_ERR_UNSUPPORTED && T.n > 32) rc = rbffd_weights_nsw(ctx, T, X, M, Y, M, stencils, center, colind_out, vals_out);
        if (rc != RBFFD_OK && rc != RBFFD_ERR_UNSUPPORTED) return rc;
        if (want_ns && rc != RBFFD_OK) RBFFD_FAIL(ctx, RBFFD_ERR_UNSUPPORTED, "null-space kernel not applicable (n=%d, q=%d) or a stencil failed its definiteness check", T.n, T.q);
    }
    if (identity && opts->kernel != 1 && opts->variant == 0) {
        if (opts->kernel != 2) {
            rc = rbffd_weights_ns(ctx, T, X, N, Y, M, stencils, nullptr, colind_out, vals_out, flags.p);
            if (rc == RBFFD_ERR_UNSUPPORTED && T.n > 32) rc = rbffd_weights_ns2(ctx, T, X, N, Y, M, stencils, nullptr, colind_out, vals_out);
            if (rc == RBFFD_ERR_UNSUPPORTED && T.n > 32) rc = rbffd_weights_nsw(ctx, T, X, N, Y, M, stencils, nullptr, colind_out, vals_out);
            if (rc != RBFFD_OK && rc != RBFFD_ERR_UNSUPPORTED) return rc;
            if (want_ns && rc != RBFFD_OK) RBFFD_FAIL(ctx, RBFFD_ERR_UNSUPPORTED, "null-space kernel not applicable (n=%d, q=%d) or a stencil failed its definiteness check", T.n, T.q);
        }
        if (rc == RBFFD_ERR_UNSUPPORTED) rc = rbffd_weights_fast(ctx, T, X, N, Y, M, stencils, colind_out, vals_out, flags.p);
        if (rc == RBFFD_ERR_UNSUPPORTED) rc = rbffd_weights_mw(ctx, T, X, N, Y, M, stencils, colind_out, vals_out, flags.p);
        if (rc != RBFFD_OK && rc != RBFFD_ERR_UNSUPPORTED) return rc;
    }
    if (rc == RBFFD_ERR_UNSUPPORTED) {
        if (opts->kernel == 2) RBFFD_FAIL(ctx, RBFFD_ERR_UNSUPPORTED, "no register/DMMA kernel for this configuration (m=%d, identity=%d)", T.m, (int)identity);
        // rows grouped by centre
        DevBuf<int> rows, seg, keys_sorted, ident;
        CUDA_TRY(ctx, rows.alloc(M, st));
        CUDA_TRY(ctx, seg.alloc(N + 1, st));
        if (identity) {
            iota_kernel<<<ceil_div_i64(M, 256), 256, 0, st>>>(rows.p, M);
            iota_kernel<<<ceil_div_i64(N + 1, 256), 256, 0, st>>>(seg.p, N + 1);
            KLAUNCH(ctx); KLAUNCH(ctx);
        } else {
            CUDA_TRY(ctx, keys_sorted.alloc(M, st));
            CUDA_TRY(ctx, ident.alloc(M, st));
            iota_kernel<<<ceil_div_i64(M, 256), 256, 0, st>>>(ident.p, M);
            KLAUNCH(ctx);
            int bits = 1;
            while ((1ll << bits) < N) ++bits;
            size_t tmp_bytes = 0;
            CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, center, keys_sorted.p, ident.p, rows.p, (int)M, 0, bits, st));
            DevBuf<unsigned char> tmp;
            CUDA_TRY(ctx, tmp.alloc(tmp_bytes, st));
            CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, center, keys_sorted.p, ident.p, rows.p, (int)M, 0, bits, st));
            seg_start_kernel<<<ceil_div_i64(M + 1, 256), 256, 0, st>>>(keys_sorted.p, M, (int)N, seg.p);
            KLAUNCH(ctx);
        }
        a.X = X; a.Y = Y; a.stencils = stencils; a.rows = rows.p; a.seg = seg.p;
        a.N = N; a.M = M; a.colind = colind_out; a.vals = vals_out; a.fail = flags.p;
        const int lda = T.m | 1;
        size_t per_warp = sizeof(double) * ((size_t)T.m * lda + (size_t)T.nops * T.m + (size_t)T.n * T.dim) + sizeof(int) * T.m;
        per_warp = (per_warp + 15) & ~(size_t)15;
        int wpb = (int)std::min<size_t>(4, (size_t)ctx->max_smem_optin / per_warp);
        if (wpb < 1) RBFFD_FAIL(ctx, RBFFD_ERR_UNSUPPORTED, "m=%d needs %zu B of shared memory per stencil (max %d)", T.m, per_warp, ctx->max_smem_optin);
        const size_t smem = per_warp * wpb;
        const int blocks_per_sm = std::max<int>(1, (int)(std::min<size_t>(ctx->max_smem_optin, 227 * 1024) / smem));
        int grid = (int)std::min<int64_t>((N + wpb - 1) / wpb, (int64_t)ctx->sm_count * blocks_per_sm * 4);
        auto launch = [&](auto kern) -> cudaError_t {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            kern<<<grid, wpb * 32, smem, st>>>(a, wpb, (int)per_warp);
            KLAUNCH(ctx);
            return cudaGetLastError();
        };
        if (T.dim == 1) CUDA_TRY(ctx, launch(weights_generic_kernel<1>));
        else if (T.dim == 2) CUDA_TRY(ctx, launch(weights_generic_kernel<2>));
        else CUDA_TRY(ctx, launch(weights_generic_kernel<3>));
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[5], st));
    if (opts->sort_columns) {
        const int wpb = 8;
        sort_rows_kernel<<<ceil_div_i64(M, wpb), wpb * 32, wpb * T.n * sizeof(int), st>>>(M, T.n, T.nops, colind_out, vals_out);
        KLAUNCH(ctx);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[6], st));
    if (deferred) return RBFFD_OK;
    CUDA_TRY(ctx, rbffd_fetch_flags(ctx, flags.p, 4, h_flags));
    float ms;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    ctx->timings[3] = ms;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]));
    ctx->timings[4] = ms;
    if (h_flags[0] != 0x7fffffff)
        RBFFD_FAIL(ctx, RBFFD_ERR_SINGULAR, "singular interpolation matrix at X node %d (0-based)", h_flags[0] - 1);
    return RBFFD_OK;
}
