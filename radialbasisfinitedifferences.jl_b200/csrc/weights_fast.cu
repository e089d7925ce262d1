// weights_fast.cu -- register-resident FP64 tensor-core (DMMA) fast path of the fused weight kernel (K4).
//
// Collocated rows (row k uses stencil k: one right-hand-side set per factorisation), m_pad = 8*MT <= 48,
// at most 8 operators.  One warp owns one stencil and never spills the matrix to shared memory:
//
//   * The scaled saddle-point matrix [Phi P; P' 0] and its right-hand sides are assembled once in a
//     shared-memory staging tile (Phi by symmetric pairs: every r^p is evaluated once), then loaded into
//     mma.m8n8k4.f64 ACCUMULATOR fragments: MT x (MT+1) tiles of 8x8, 2 doubles per lane per tile; the
//     last tile column holds the right-hand sides.
//   * Blocked Gauss-Jordan elimination with partial pivoting, 4 columns per block step:
//       1. the 4 panel columns are transposed through shared memory into a row-per-lane layout,
//       2. 4 pivoted elimination steps run on the panel in registers (pivot = CREDUX.MAX over packed
//          |a| keys, pivot row broadcast with SHFL); every row accumulates its row W[i, 0:4] of the
//          rank-4 transform  T = I + W * E_pivots'  (rows never move: implicit permutation),
//       3. W becomes the A fragments, the 4 RAW pivot rows (dumped by their owner lanes) the B fragments,
//          and ONE pass of DMMAs applies  X += W * X[pivots, :]  to every remaining tile and to the RHS.
//     Gauss-Jordan needs no back substitution and no stored factors: the solution is RHS_row / pivot_row.
//   * weights are rescaled (generate_operator.jl:161-166) and scattered straight into the CSR rows.
//
// Replaces the same reference lines as weights.cu (scalestencil.jl:10-20, interpolationmatrix.jl:5-8,
// generate_operator.jl:89-167, hyperviscosity_operator.jl:97-161).  Flop accounting in bench.py uses the
// LU convention (2/3 m^3 + 2 m^2 r) although this kernel executes ~m^3 (all row tiles every step).
#include "common.cuh"
#include "tables.cuh"

namespace {

struct FArgs {
    const double* X;
    const double* Y;
    const int32_t* stencils;   // [NS][n]
    int64_t NS, M;
    int32_t* colind;           // [M][n]
    double* vals;              // [nops][M][n]
    int* fail;
    OpTables T;
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int pad4mod16(int x) {   // smallest y >= x with y % 16 == 4  (conflict-free fragment strides)
    int y = x;
    while (y % 16 != 4) ++y;
    return y;
}

#ifndef FAST_MIN_BLOCKS
#define FAST_MIN_BLOCKS 3
#endif

template <int D, int MT>
struct FastCfg {
    static constexpr int MP = 8 * MT;            // padded system size
    static constexpr int NT = MT + 1;            // tile columns incl. the RHS tile
    static constexpr int NC = 8 * NT;
    static constexpr int LDG = (MP + 8) | 1;     // staging row stride (odd): columns [0,MP) matrix (identity padded), [MP,MP+nops) RHS
    static constexpr int PS = pad4mod16(MP);     // panel / multiplier buffer stride
    static constexpr int US = pad4mod16(NC);     // pivot-row buffer stride
    static constexpr int NSMAX = 48;             // max stencil size n
    static constexpr int STAGE = MP * LDG;       // doubles
    // Pbuf/Lbuf/Ubuf/rinv/pivcol alias the staging tile once the fragments are loaded
    static constexpr int SMALL = 4 * PS + 4 * PS + 4 * US + MP + (MP + 1) / 2;
    static_assert(SMALL <= STAGE, "staging tile too small for the exchange buffers");
    static constexpr int DOUBLES = STAGE + NSMAX * D;
    static constexpr int BYTES_PER_WARP = ((DOUBLES * 8) + 15) & ~15;
    static constexpr int WARPS = 4;
};

template <int D, int MT>
__global__ void __launch_bounds__(128, (MT <= 5 ? FAST_MIN_BLOCKS : 2)) weights_dmma_kernel(FArgs a) {
    using C = FastCfg<D, MT>;
    constexpr int MP = C::MP, NT = C::NT, LDG = C::LDG, PS = C::PS, US = C::US;
    extern __shared__ __align__(16) unsigned char fsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const OpTables& T = a.T;
    const int n = T.n, q = T.q, m = T.m, nops = T.nops;
    unsigned char* base = fsm + (size_t)warp * C::BYTES_PER_WARP;
    double* G = reinterpret_cast<double*>(base);       // staging [MP][LDG]
    double* S = G + C::STAGE;                          // [n][D] scaled offsets
    double* Pbuf = G;                                  // [4][PS] panel columns (column-major)   (aliases G)
    double* Lbuf = Pbuf + 4 * PS;                      // [4][PS] transform rows W
    double* Ubuf = Lbuf + 4 * PS;                      // [4][US] raw pivot rows
    double* rinv_s = Ubuf + 4 * US;                    // [MP]
    int* pivcol_s = reinterpret_cast<int*>(rinv_s + MP);   // [MP]
    // per-lane exchange addresses (independent of the stencil)
    double* const pb_w = Pbuf + (2 * (t & 1)) * PS + g;    // extraction: + e*PS + 8*I
    const double* const pb_r = Pbuf + lane;                // row-per-lane: + cc*PS + 32*z
    double* const lb_w = Lbuf + lane;
    const double* const lb_r = Lbuf + t * PS + g;          // A fragment: + 8*I
    const double* const ub_r = Ubuf + t * US + g;          // B fragment: + 8*J
    const double EPS = 2.220446049250313e-16;
    const unsigned FULL = 0xffffffffu;

    for (int64_t i = blockIdx.x * (int64_t)C::WARPS + warp; i < a.NS; i += (int64_t)gridDim.x * C::WARPS) {
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
            s[c] = 1.0 / mx[c];
        }
        __syncwarp();
        for (int j = lane; j < n; j += 32) {
#pragma unroll
            for (int c = 0; c < D; ++c) S[j * D + c] = S[j * D + c] * s[c];
        }
        __syncwarp();
        // ---- assemble [Phi P; P' 0 | RHS] in the staging tile; Phi by symmetric pairs ----
        {
            const int half = (n + 1) >> 1;
            const int hp = (T.p - 1) >> 1;
            const int hp0 = hp;
            for (int tt = 0; tt < half; ++tt) {
                const int i2 = n - 1 - tt, n1 = n - 1 - tt;
                for (int cidx = lane; cidx < n - 1; cidx += 32) {
                    int ia, ib;
                    if (cidx < n1) { ia = tt; ib = tt + 1 + cidx; }
                    else { if (i2 == tt) continue; ia = i2; ib = i2 + 1 + (cidx - n1); }
                    double r2 = 0.0;
#pragma unroll
                    for (int c = 0; c < D; ++c) { double dd = S[ia * D + c] - S[ib * D + c]; r2 += dd * dd; }
                    double v = fast_sqrt(r2);
                    for (int e = 0; e < hp; ++e) v *= r2;
                    G[ia * LDG + ib] = v;
                    G[ib * LDG + ia] = v;
                }
            }
            for (int j = lane; j < n; j += 32) {
                double* grow = G + j * LDG;
                grow[j] = 0.0;
                grow[n] = 1.0;
                G[n * LDG + j] = 1.0;
                for (int tq = 1; tq < q; ++tq) {
                    const double v = grow[n + T.mpar[tq]] * S[j * D + T.maxis[tq]];
                    grow[n + tq] = v;
                    G[(n + tq) * LDG + j] = v;
                }
            }
            {
                double* zrow = G + n * LDG + n;
                for (int qa = 0; qa < q; ++qa, zrow += LDG)
                    for (int qb = lane; qb < q; qb += 32) zrow[qb] = 0.0;
            }
            if (m < MP) {   // identity padding rows / columns
                for (int r_ = m; r_ < MP; ++r_)
                    for (int cq = lane; cq < MP; cq += 32) { G[r_ * LDG + cq] = r_ == cq ? 1.0 : 0.0; if (cq < m) G[cq * LDG + r_] = 0.0; }
            }
            // right-hand sides of row k = i (generate_operator.jl:110-157)
            double eta[D];
#pragma unroll
            for (int c = 0; c < D; ++c) eta[c] = (a.Y[i * D + c] - xc[c]) * s[c];
            bool eta_zero = true;
#pragma unroll
            for (int c = 0; c < D; ++c) eta_zero = eta_zero && (eta[c] == 0.0);
            for (int j = lane; j < n; j += 32) {
                double del[D];
                double r2 = 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    double dd = eta[c] - S[j * D + c];
                    del[c] = dd == 0.0 ? EPS : dd;
                    r2 += del[c] * del[c];
                }
                const double r = fast_sqrt(r2);
                double rp2 = T.p >= 3 ? r : fast_rcp(r);
                for (int e = 1; e < hp0; ++e) rp2 *= r2;
                const double rp = rp2 * r2, rp4 = rp2 * fast_rcp(r2);
                for (int o = 0; o < nops; ++o) G[j * LDG + MP + o] = rhs_rbf_entry_fast<D>(T, o, del, s, r, r2, rp, rp2, rp4);
            }
            if (eta_zero) {
                for (int tq = lane; tq < q; tq += 32)
                    for (int o = 0; o < nops; ++o) G[(n + tq) * LDG + MP + o] = rhs_poly_entry_at_zero<D>(T, o, tq, s);
            } else {
                for (int tq = lane; tq < q; tq += 32)
                    for (int o = 0; o < nops; ++o) G[(n + tq) * LDG + MP + o] = rhs_poly_entry<D>(T, o, tq, eta, s);
            }
        }
        __syncwarp();
        // ---- accumulator fragments: lane (g,t) holds rows 8I+g, columns 8J+2t, 8J+2t+1 ----
        double c[MT][NT][2];
        {
            const double* gl = G + g * LDG + 2 * t;
#pragma unroll
            for (int I = 0; I < MT; ++I) {
#pragma unroll
                for (int J = 0; J < MT; ++J) {
                    c[I][J][0] = gl[8 * I * LDG + 8 * J];
                    c[I][J][1] = gl[8 * I * LDG + 8 * J + 1];
                }
                const double r0 = gl[8 * I * LDG + MP], r1 = gl[8 * I * LDG + MP + 1];
                c[I][MT][0] = (8 * I + g < m && 2 * t < nops) ? r0 : 0.0;
                c[I][MT][1] = (8 * I + g < m && 2 * t + 1 < nops) ? r1 : 0.0;
            }
        }
        __syncwarp();     // staging tile is dead from here on: Pbuf/Lbuf/Ubuf alias it

        // ---- blocked Gauss-Jordan, row-per-lane bookkeeping: lane owns rows lane (slot 0) and lane+32 (slot 1) ----
        constexpr int NZ = MP > 32 ? 2 : 1;
        bool done[NZ];
#pragma unroll
        for (int z = 0; z < NZ; ++z) done[z] = (lane + 32 * z) >= MP;
        bool ok = true;
        const bool has1 = lane + 32 < MP;

#pragma unroll
        for (int kb = 0; kb < 2 * MT; ++kb) {
            const int Jp = kb >> 1, h = kb & 1;
            // 1. panel columns -> Pbuf (column-major)
            if ((t >> 1) == h) {
#pragma unroll
                for (int I = 0; I < MT; ++I) {
                    pb_w[8 * I] = c[I][Jp][0];
                    pb_w[PS + 8 * I] = c[I][Jp][1];
                }
            }
            __syncwarp();
            // 2. row-per-lane load
            double av[NZ][4], w[NZ][4];
#pragma unroll
            for (int z = 0; z < NZ; ++z)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    av[z][cc] = (z == 0 || has1) ? pb_r[cc * PS + 32 * z] : 0.0;
                    w[z][cc] = 0.0;
                }
            // 3. four pivoted Gauss-Jordan steps on the panel
            unsigned prows = 0;      // 4 pivot rows, 8 bits each (warp-uniform)
#pragma unroll
            for (int sidx = 0; sidx < 4; ++sidx) {
                unsigned key = 0;
#pragma unroll
                for (int z = 0; z < NZ; ++z) {
                    unsigned hi = (unsigned)__double2hiint(av[z][sidx]) & 0x7fffffc0u;
                    unsigned kz = done[z] ? 0u : (hi | (unsigned)(lane + 32 * z));
                    key = max(key, kz);
                }
                const unsigned kmax = __reduce_max_sync(FULL, key);
                if (kmax < 64u) ok = false;          // zero (or denormal) pivot column: singular
                const int pr = kmax & 63;            // pivot row = lane + 32 * slot
                const int pl = pr & 31;
                prows |= (unsigned)pr << (8 * sidx);
                double pv[4], wp[4];
#pragma unroll
                for (int cc = sidx; cc < 4; ++cc) {
                    double src = av[0][cc];
                    if (NZ > 1) src = (pr & 32) ? av[NZ - 1][cc] : src;
                    pv[cc] = __shfl_sync(FULL, src, pl);
                }
#pragma unroll
                for (int cc = 0; cc < sidx; ++cc) {
                    double src = w[0][cc];
                    if (NZ > 1) src = (pr & 32) ? w[NZ - 1][cc] : src;
                    wp[cc] = __shfl_sync(FULL, src, pl);
                }
                const double rinv = fast_rcp(pv[sidx]);
                rinv_s[pr] = rinv;                 // same value from every lane: benign, branch-free
                pivcol_s[pr] = 4 * kb + sidx;
#pragma unroll
                for (int z = 0; z < NZ; ++z) {
                    const bool ispiv = (lane + 32 * z) == pr;
                    const double nl = ispiv ? 0.0 : av[z][sidx] * (-rinv);
#pragma unroll
                    for (int cc = sidx + 1; cc < 4; ++cc) av[z][cc] = fma(nl, pv[cc], av[z][cc]);
#pragma unroll
                    for (int cc = 0; cc < sidx; ++cc) w[z][cc] = fma(nl, wp[cc], w[z][cc]);
                    w[z][sidx] = nl;
                    done[z] = done[z] || ispiv;
                }
            }
            // 4. transform rows W -> Lbuf -> A fragments
#pragma unroll
            for (int z = 0; z < NZ; ++z) {
                if (z == 0 || has1) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) lb_w[cc * PS + 32 * z] = w[z][cc];
                }
            }
            // 5. raw pivot rows -> Ubuf: warp-uniform switch on the pivot's tile row, owner lanes (g == row & 7) store
            const int jlo = h == 0 ? Jp : Jp + 1;       // first tile column still alive (tile Jp's right half when h == 0)
#pragma unroll
            for (int sidx = 0; sidx < 4; ++sidx) {
                const int pr = (prows >> (8 * sidx)) & 63;
                const bool mine = g == (pr & 7);
                double2* dst = reinterpret_cast<double2*>(Ubuf + sidx * US + 2 * t);
                switch (pr >> 3) {
#define RBFFD_DUMP_CASE(II)                                                                           \
                    case II:                                                                             \
                        if (II < MT && mine) {                                                           \
                            _Pragma("unroll") for (int J = 0; J < NT; ++J)                               \
                                if (J >= jlo) dst[4 * J] = make_double2(c[II < MT ? II : 0][J][0], c[II < MT ? II : 0][J][1]); \
                        }                                                                                \
                        break;
                    RBFFD_DUMP_CASE(0) RBFFD_DUMP_CASE(1) RBFFD_DUMP_CASE(2) RBFFD_DUMP_CASE(3) RBFFD_DUMP_CASE(4) RBFFD_DUMP_CASE(5)
#undef RBFFD_DUMP_CASE
                    default: break;
                }
            }
            __syncwarp();
            double af[MT];
#pragma unroll
            for (int I = 0; I < MT; ++I) af[I] = lb_r[8 * I];
            // 6. X += W * X[pivots, :] on every live tile (DMMA)
#pragma unroll
            for (int J = 0; J < NT; ++J) {
                if (J >= jlo) {
                    const double bf = ub_r[8 * J];
#pragma unroll
                    for (int I = 0; I < MT; ++I) dmma884(c[I][J][0], c[I][J][1], af[I], bf);
                }
            }
            __syncwarp();
        }

        // ---- solution = RHS_row / pivot_row, rescale, scatter into the CSR row ----
        const int64_t krow = i;
        double f[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) f[e] = (2 * t + e) < nops ? op_post_factor<D>(T, 2 * t + e, s) : 0.0;
#pragma unroll
        for (int I = 0; I < MT; ++I) {
            const int row = 8 * I + g;
            const int pc = pivcol_s[row];
            const double ri = rinv_s[row];
            if (pc < n) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int o = 2 * t + e;
                    if (o < nops) {
                        const double v = ok ? f[e] * (c[I][MT][e] * ri) : nan("");
                        a.vals[((int64_t)o * a.M + krow) * n + pc] = v;
                    }
                }
            }
        }
        for (int j = lane; j < n; j += 32) a.colind[krow * n + j] = st[j];
        if (!ok && lane == 0) atomicMin(a.fail, (int)i + 1);
        __syncwarp();
    }
}

template <int D, int MT>
int launch_fast(rbffd_context* ctx, const FArgs& a) {
    using C = FastCfg<D, MT>;
    const size_t smem = (size_t)C::BYTES_PER_WARP * C::WARPS;
    if ((int64_t)smem > ctx->max_smem_optin) return RBFFD_ERR_UNSUPPORTED;
    auto kern = weights_dmma_kernel<D, MT>;
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks_needed = (a.NS + C::WARPS - 1) / C::WARPS;
    const int grid = (int)std::min<int64_t>(blocks_needed, (int64_t)ctx->sm_count * 3 * 8);
    kern<<<grid, C::WARPS * 32, smem, ctx->stream>>>(a);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

template <int D>
int dispatch_mt(rbffd_context* ctx, const FArgs& a, int mt) {
    switch (mt) {
        case 3: return launch_fast<D, 3>(ctx, a);
        case 4: return launch_fast<D, 4>(ctx, a);
        case 5: return launch_fast<D, 5>(ctx, a);
        case 6: return launch_fast<D, 6>(ctx, a);
        default: return RBFFD_ERR_UNSUPPORTED;
    }
}

}  // namespace

// Collocated fast path; RBFFD_ERR_UNSUPPORTED tells the caller to use the generic kernel.
int rbffd_weights_fast(rbffd_context* ctx, const OpTables& T, const double* X, int64_t NS, const double* Y, int64_t M,
                       const int32_t* stencils, int32_t* colind_out, double* vals_out, int* fail_flag) {
    if (T.nops > 8 || T.n > 48 || T.m > 48 || T.dim < 2 || NS != M) return RBFFD_ERR_UNSUPPORTED;
    const int mt = std::max(3, (T.m + 7) / 8);
    FArgs a;
    a.X = X; a.Y = Y; a.stencils = stencils; a.NS = NS; a.M = M;
    a.colind = colind_out; a.vals = vals_out; a.fail = fail_flag; a.T = T;
    if (T.dim == 2) return dispatch_mt<2>(ctx, a, mt);
    return dispatch_mt<3>(ctx, a, mt);
}
