// weights_mw.cu -- multi-warp variant of the register-resident DMMA Gauss-Jordan weight kernel (K4) for
// 48 < m <= 96 (BASELINE configs 3-5: m = 65 and m = 80).  One CTA of NW warps owns one stencil:
//   * tile COLUMNS of the accumulator-fragment matrix are dealt cyclically to the warps (tile column J lives in
//     warp J % NW), so every warp keeps MT x ceil((MT+1)/NW) tiles in registers and the DMMA work stays balanced
//     while columns die from the left;
//   * the pivoted panel steps run with ONE ROW PER THREAD (m_pad <= 32*NW): warp-level CREDUX.MAX, the warp
//     winners publish their candidate row through shared memory, one __syncthreads per column step;
//   * transform rows W and the raw pivot rows are exchanged through shared memory exactly as in weights_fast.cu,
//     each warp dumps / updates only its own tile columns.
// Same algorithm, same reference lines as weights_fast.cu; see the header of that file.
#include "common.cuh"
#include "tables.cuh"

namespace {

struct MArgs {
    const double* X;
    const double* Y;
    const int32_t* stencils;   // [NS][n]
    int64_t NS, M;
    int32_t* colind;           // [M][n]
    double* vals;              // [nops][M][n]
    int* fail;
    OpTables T;
};

__device__ __forceinline__ void dmma884m(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int pad4mod16m(int x) {
    int y = x;
    while (y % 16 != 4) ++y;
    return y;
}

template <int D, int MT, int NW>
struct MwCfg {
    static constexpr int MP = 8 * MT;
    static constexpr int NT = MT + 1;
    static constexpr int NC = 8 * NT;
    static constexpr int JW = (NT + NW - 1) / NW;      // tile columns per warp
    static constexpr int LDG = (MP + 8) | 1;
    static constexpr int PS = pad4mod16m(MP);
    static constexpr int US = pad4mod16m(NC);
    static constexpr int NSMAX = 96;
    static constexpr int STAGE = MP * LDG;
    static constexpr int SMALL = 4 * PS + 4 * PS + 4 * US + MP + (MP + 1) / 2;
    static_assert(SMALL <= STAGE, "staging tile too small for the exchange buffers");
    static_assert(MP <= 32 * NW, "one row per thread");
    static constexpr int CAND = 2 * NW * 4 + 2 * NW;    // candidate rows (doubles) + keys (stored as doubles' worth)
    static constexpr int DOUBLES = STAGE + NSMAX * D + CAND;
    static constexpr int BYTES = ((DOUBLES * 8) + 15) & ~15;
    static constexpr int THREADS = 32 * NW;
};

template <int D, int MT, int NW>
__global__ void __launch_bounds__(32 * NW, (NW == 2 ? 4 : (MT <= 10 ? 3 : 2)))
weights_dmma_mw_kernel(MArgs a) {
    using C = MwCfg<D, MT, NW>;
    constexpr int MP = C::MP, NT = C::NT, JW = C::JW, LDG = C::LDG, PS = C::PS, US = C::US, NTH = C::THREADS;
    extern __shared__ __align__(16) unsigned char msm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const OpTables& T = a.T;
    const int n = T.n, q = T.q, m = T.m, nops = T.nops;
    double* G = reinterpret_cast<double*>(msm);
    double* S = G + C::STAGE;
    double* cand_val = S + C::NSMAX * D;                       // [2][NW][4]
    unsigned* cand_key = reinterpret_cast<unsigned*>(cand_val + 2 * NW * 4);   // [2][NW]
    double* Pbuf = G;
    double* Lbuf = Pbuf + 4 * PS;
    double* Ubuf = Lbuf + 4 * PS;
    double* rinv_s = Ubuf + 4 * US;
    int* pivcol_s = reinterpret_cast<int*>(rinv_s + MP);
    double* const pb_w = Pbuf + (2 * (t & 1)) * PS + g;
    const double* const lb_r = Lbuf + t * PS + g;
    const double* const ub_r = Ubuf + t * US + g;
    const double EPS = 2.220446049250313e-16;
    const unsigned FULL = 0xffffffffu;

    for (int64_t i = blockIdx.x; i < a.NS; i += gridDim.x) {
        const int32_t* st = a.stencils + i * n;
        const int c0 = st[0];
        double xc[D], s[D];
#pragma unroll
        for (int c = 0; c < D; ++c) xc[c] = a.X[(int64_t)c0 * D + c];
        // ---- scalestencil.jl:10-20 (every warp reduces the whole stencil: identical s in all warps) ----
        double mx[D];
#pragma unroll
        for (int c = 0; c < D; ++c) mx[c] = 0.0;
        for (int j = lane; j < n; j += 32) {
            const int id = st[j];
#pragma unroll
            for (int c = 0; c < D; ++c) mx[c] = fmax(mx[c], fabs(a.X[(int64_t)id * D + c] - xc[c]));
        }
#pragma unroll
        for (int c = 0; c < D; ++c) {
            for (int o = 16; o > 0; o >>= 1) mx[c] = fmax(mx[c], __shfl_xor_sync(FULL, mx[c], o));
            s[c] = 1.0 / mx[c];
        }
        for (int j = tid; j < n; j += NTH) {
            const int id = st[j];
#pragma unroll
            for (int c = 0; c < D; ++c) S[j * D + c] = (a.X[(int64_t)id * D + c] - xc[c]) * s[c];
        }
        __syncthreads();
        // ---- assemble [Phi P; P' 0 | RHS] in the staging tile (all warps) ----
        {
            const int half = (n + 1) >> 1;
            const int hp = (T.p - 1) >> 1;
            for (int tt = warp; tt < half; tt += NW) {
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
            for (int j = tid; j < n; j += NTH) {
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
            for (int qa = warp; qa < q; qa += NW)
                for (int qb = lane; qb < q; qb += 32) G[(n + qa) * LDG + n + qb] = 0.0;
            if (m < MP) {
                for (int r_ = m + warp; r_ < MP; r_ += NW)
                    for (int cq = lane; cq < MP; cq += 32) { G[r_ * LDG + cq] = r_ == cq ? 1.0 : 0.0; if (cq < m) G[cq * LDG + r_] = 0.0; }
            }
            double eta[D];
            bool eta_zero = true;
#pragma unroll
            for (int c = 0; c < D; ++c) { eta[c] = (a.Y[i * D + c] - xc[c]) * s[c]; eta_zero = eta_zero && (eta[c] == 0.0); }
            for (int j = tid; j < n; j += NTH) {
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
                for (int e = 1; e < hp; ++e) rp2 *= r2;
                const double rp = rp2 * r2, rp4 = rp2 * fast_rcp(r2);
                for (int o = 0; o < nops; ++o) G[j * LDG + MP + o] = rhs_rbf_entry_fast<D>(T, o, del, s, r, r2, rp, rp2, rp4);
            }
            for (int tq = tid; tq < q; tq += NTH)
                for (int o = 0; o < nops; ++o)
                    G[(n + tq) * LDG + MP + o] = eta_zero ? rhs_poly_entry_at_zero<D>(T, o, tq, s) : rhs_poly_entry<D>(T, o, tq, eta, s);
        }
        __syncthreads();
        // ---- accumulator fragments: warp w owns tile columns J = w + NW*jl ----
        double c[MT][JW][2];
        {
            const double* gl = G + g * LDG + 2 * t;
#pragma unroll
            for (int jl = 0; jl < JW; ++jl) {
                const int J = warp + NW * jl;
#pragma unroll
                for (int I = 0; I < MT; ++I) {
                    if (J < MT) {
                        c[I][jl][0] = gl[8 * I * LDG + 8 * J];
                        c[I][jl][1] = gl[8 * I * LDG + 8 * J + 1];
                    } else if (J == MT) {
                        const double r0 = gl[8 * I * LDG + MP], r1 = gl[8 * I * LDG + MP + 1];
                        c[I][jl][0] = (8 * I + g < m && 2 * t < nops) ? r0 : 0.0;
                        c[I][jl][1] = (8 * I + g < m && 2 * t + 1 < nops) ? r1 : 0.0;
                    } else {
                        c[I][jl][0] = 0.0;
                        c[I][jl][1] = 0.0;
                    }
                }
            }
        }
        __syncthreads();     // staging tile is dead: Pbuf/Lbuf/Ubuf alias it

        bool done = tid >= MP;
        bool ok = true;
#pragma unroll
        for (int kb = 0; kb < 2 * MT; ++kb) {
            const int Jp = kb >> 1, h = kb & 1;
            // 1. owner warp writes the panel columns
            if (warp == Jp % NW && (t >> 1) == h) {
#pragma unroll
                for (int I = 0; I < MT; ++I) {
                    pb_w[8 * I] = c[I][Jp / NW][0];
                    pb_w[PS + 8 * I] = c[I][Jp / NW][1];
                }
            }
            __syncthreads();
            // 2. one row per thread
            double av[4], w[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) { av[cc] = tid < MP ? Pbuf[cc * PS + tid] : 0.0; w[cc] = 0.0; }
            // 3. four pivoted Gauss-Jordan steps; one __syncthreads each
            unsigned prows = 0;
#pragma unroll
            for (int sidx = 0; sidx < 4; ++sidx) {
                const int par = sidx & 1;
                const unsigned hi = (unsigned)__double2hiint(av[sidx]) & 0x7fffff80u;
                const unsigned key = done ? 0u : (hi | (unsigned)tid);
                const unsigned wmax = __reduce_max_sync(FULL, key);
                if (key == wmax && wmax != 0u) {
                    double* cv = cand_val + (par * NW + warp) * 4;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) cv[cc] = cc >= sidx ? av[cc] : w[cc];
                }
                if (lane == 0) cand_key[par * NW + warp] = wmax;
                __syncthreads();
                unsigned kmax = cand_key[par * NW];
                int ww = 0;
#pragma unroll
                for (int x = 1; x < NW; ++x) {
                    const unsigned kx = cand_key[par * NW + x];
                    if (kx > kmax) { kmax = kx; ww = x; }
                }
                if (kmax < 128u) ok = false;
                const int pr = kmax & 127;
                prows |= (unsigned)pr << (8 * sidx);
                double pv[4];
                const double* cv = cand_val + (par * NW + ww) * 4;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) pv[cc] = cv[cc];
                const double rinv = fast_rcp(pv[sidx]);
                if (warp == 0) { rinv_s[pr] = rinv; pivcol_s[pr] = 4 * kb + sidx; }   // same value from every lane
                const bool ispiv = tid == pr;
                const double nl = ispiv ? 0.0 : av[sidx] * (-rinv);
#pragma unroll
                for (int cc = sidx + 1; cc < 4; ++cc) av[cc] = fma(nl, pv[cc], av[cc]);
#pragma unroll
                for (int cc = 0; cc < sidx; ++cc) w[cc] = fma(nl, pv[cc], w[cc]);
                w[sidx] = nl;
                done = done || ispiv;
            }
            // 4. transform rows
            if (tid < MP) {
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) Lbuf[cc * PS + tid] = w[cc];
            }
            // 5. raw pivot rows of this warp's tile columns
            const int jlo = h == 0 ? Jp : Jp + 1;
#pragma unroll
            for (int sidx = 0; sidx < 4; ++sidx) {
                const int pr = (prows >> (8 * sidx)) & 127;
                const bool mine = g == (pr & 7);
                double2* dst = reinterpret_cast<double2*>(Ubuf + sidx * US + 2 * t);
                switch (pr >> 3) {
#define RBFFD_DUMP_CASE(II)                                                                             \
                    case II:                                                                               \
                        if (II < MT && mine) {                                                             \
                            _Pragma("unroll") for (int jl = 0; jl < JW; ++jl) {                            \
                                const int J = warp + NW * jl;                                              \
                                if (J >= jlo && J < NT)                                                    \
                                    dst[4 * J] = make_double2(c[II < MT ? II : 0][jl][0], c[II < MT ? II : 0][jl][1]); \
                            }                                                                              \
                        }                                                                                  \
                        break;
                    RBFFD_DUMP_CASE(0) RBFFD_DUMP_CASE(1) RBFFD_DUMP_CASE(2) RBFFD_DUMP_CASE(3) RBFFD_DUMP_CASE(4) RBFFD_DUMP_CASE(5)
                    RBFFD_DUMP_CASE(6) RBFFD_DUMP_CASE(7) RBFFD_DUMP_CASE(8) RBFFD_DUMP_CASE(9) RBFFD_DUMP_CASE(10) RBFFD_DUMP_CASE(11)
#undef RBFFD_DUMP_CASE
                    default: break;
                }
            }
            __syncthreads();
            // 6. X += W * X[pivots, :] on this warp's live tiles
            double bf[JW];
            bool live[JW];
#pragma unroll
            for (int jl = 0; jl < JW; ++jl) {
                const int J = warp + NW * jl;
                live[jl] = J >= jlo && J < NT;
                bf[jl] = live[jl] ? ub_r[8 * J] : 0.0;
            }
#pragma unroll
            for (int I = 0; I < MT; ++I) {
                const double af = lb_r[8 * I];
#pragma unroll
                for (int jl = 0; jl < JW; ++jl)
                    if (live[jl]) dmma884m(c[I][jl][0], c[I][jl][1], af, bf[jl]);
            }
            // the next block step's first __syncthreads (after the panel write) orders these reads before the
            // next overwrite of Lbuf/Ubuf; Pbuf is not read after step 2
        }
        __syncthreads();
        // ---- solution = RHS_row / pivot_row: the warp that owns the RHS tile scatters the CSR row ----
        if (warp == MT % NW) {
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
                        if (o < nops) a.vals[((int64_t)o * a.M + i) * n + pc] = ok ? f[e] * (c[I][MT / NW][e] * ri) : nan("");
                    }
                }
            }
        }
        for (int j = tid; j < n; j += NTH) a.colind[i * n + j] = st[j];
        if (!ok && tid == 0) atomicMin(a.fail, (int)i + 1);
        __syncthreads();     // exchange buffers alias the next stencil's staging tile
    }
}

template <int D, int MT, int NW>
int launch_mw(rbffd_context* ctx, const MArgs& a) {
    using C = MwCfg<D, MT, NW>;
    const size_t smem = (size_t)C::BYTES;
    if ((int64_t)smem > ctx->max_smem_optin) return RBFFD_ERR_UNSUPPORTED;
    auto kern = weights_dmma_mw_kernel<D, MT, NW>;
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>(a.NS, (int64_t)ctx->sm_count * 3 * 8);
    kern<<<grid, C::THREADS, smem, ctx->stream>>>(a);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

template <int D>
int dispatch_mw(rbffd_context* ctx, const MArgs& a, int mt) {
    switch (mt) {
        case 7: return launch_mw<D, 7, 2>(ctx, a);
        case 8: return launch_mw<D, 8, 2>(ctx, a);
        case 9: return launch_mw<D, 9, 4>(ctx, a);
        case 10: return launch_mw<D, 10, 4>(ctx, a);
        case 11: return launch_mw<D, 11, 4>(ctx, a);
        case 12: return launch_mw<D, 12, 4>(ctx, a);
        default: return RBFFD_ERR_UNSUPPORTED;
    }
}

}  // namespace

// Collocated rows, 48 < m <= 96, <= 8 operators; RBFFD_ERR_UNSUPPORTED otherwise.
int rbffd_weights_mw(rbffd_context* ctx, const OpTables& T, const double* X, int64_t NS, const double* Y, int64_t M,
                     const int32_t* stencils, int32_t* colind_out, double* vals_out, int* fail_flag) {
    if (T.nops > 8 || T.n > 96 || T.m > 96 || T.m <= 48 || T.dim < 2 || NS != M) return RBFFD_ERR_UNSUPPORTED;
    const int mt = (T.m + 7) / 8;
    MArgs a;
    a.X = X; a.Y = Y; a.stencils = stencils; a.NS = NS; a.M = M;
    a.colind = colind_out; a.vals = vals_out; a.fail = fail_flag; a.T = T;
    if (T.dim == 2) return dispatch_mw<2>(ctx, a, mt);
    return dispatch_mw<3>(ctx, a, mt);
}
