// weights_ns.cu -- null-space (range-space-free) fast path of the fused weight kernel (K4): pivoting only where the
// mathematics needs it.
//
// The saddle-point system  [Phi P; P' 0] [w; lam] = [b; g]  (src/interpolationmatrix.jl:5, generate_operator.jl:157)
// is solved for w without ever factorising the indefinite (n+q) x (n+q) matrix:
//   1. column reduction of P (n x q, the only pivoted part: q steps instead of n+q) picks q "basic" stencil nodes and
//      gives, row by row,  P C = [I; W'] ; the same column operations applied to g' give a particular solution
//      w_p (supported on the basic nodes) of  P' w_p = g ;
//   2. Z = [-W; I] spans null(P').  PHS r^p is conditionally definite of order (p+1)/2 <= polydeg+1, so
//      S = Z' Phi Z  is definite ((-1)^((p+1)/2) S is SPD): it is formed with FP64 tensor-core DMMAs
//      (Y = Phi[:,N] - Phi[:,B] W,  S = Y[N,:] - W' Y[B,:]; the right-hand sides ride along as a fourth tile column)
//      and eliminated WITHOUT pivoting;
//   3. w[N] = S^-1 Z'(b - Phi w_p),  w[B] = w_p - W w[N].
// Measured against extended precision on the BASELINE stencil families the error is <= 0.06 eps cond(A), the same as
// pivoted LU; a non-definite S or a rank-deficient P raises a flag and the batch is redone by the
// Gauss-Jordan kernel (weights_fast.cu).  Work: ~(n-q)^2 (n+q) flops instead of (n+q)^3, 10 pivot searches instead
// of 40 at config 2, and half the registers, i.e. more resident warps.
// Scope: collocated rows, n <= 32, q <= 12, n - q <= 24, <= 8 operators, polydeg >= (p-1)/2.
// Replaces the same reference lines as weights.cu.
#include "common.cuh"
#include "tables.cuh"

namespace {

struct NArgs {
    const double* X;
    const double* Y;
    const int32_t* stencils;   // [NS][n]
    int64_t NS, M;
    int32_t* colind;           // [M][n]
    double* vals;              // [nops][M][n]
    int* fail;                 // flags[0]: singular node + 1
    int* redo;                 // set to 1 when any stencil needs the pivoted fallback
    OpTables T;
};

__device__ __forceinline__ void dmma884n(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int NS_QP = 12;      // padded polynomial count (3 k-steps of 4)
constexpr int NS_NB = 24;      // padded null-space dimension (3 tiles)
constexpr int NS_LD = 33;      // row stride of the Phi tile
constexpr int NS_US = 34;      // row stride of the Y / [S|t] exchange tile (even: 16-byte rows)
constexpr int NS_WARPS = 4;

template <int D>
struct NsCfg {
    // doubles per warp
    static constexpr int G = 32 * NS_US;                 // Phi~ (permuted, stride NS_LD); later the Y tile, then the [S|t] tile
    static constexpr int YB = 0;                         // (Y aliases G)
    static constexpr int WT = NS_NB * NS_QP;             // W' rows of the non-basic nodes
    static constexpr int WP = NS_QP * 8;                 // particular solutions, [c][op]
    static constexpr int BT = 32 * 8;                    // RBF right-hand sides, [pos][op]
    static constexpr int SC = 32 * D;                    // permuted scaled coordinates
    static constexpr int PR = NS_QP + 4;                 // pivot-row broadcast buffer
    static constexpr int YS = 8 * NS_NB;                 // solution y, [op][a]
    static constexpr int DOUBLES = G + YB + WT + WP + BT + SC + PR + YS;
    static constexpr int BYTES_PER_WARP = ((DOUBLES * 8 + 32 * 4) + 15) & ~15;   // + perm[32]
};

template <int D>
__global__ void __launch_bounds__(128, 3) weights_ns_kernel(NArgs a) {
    using C = NsCfg<D>;
    constexpr int LD = NS_LD, US = NS_US, QP = NS_QP;
    extern __shared__ __align__(16) unsigned char nsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const OpTables& T = a.T;
    const int n = T.n, q = T.q, nops = T.nops, nb = n - q;
    double* G = reinterpret_cast<double*>(nsm + (size_t)warp * C::BYTES_PER_WARP);
    double* Yb = G;                                   // aliases G: written only after every read of Phi~ is done
    double* Wt = G + C::G;
    double* WpT = Wt + C::WT;
    double* Bt = WpT + C::WP;
    double* Sc = Bt + C::BT;
    double* Pr = Sc + C::SC;
    double* Ys = Pr + C::PR;
    int* perm = reinterpret_cast<int*>(Ys + C::YS);
    const double EPS = 2.220446049250313e-16;
    const unsigned FULL = 0xffffffffu;
    const double sgn = (((T.p + 1) >> 1) & 1) ? -1.0 : 1.0;       // (-1)^((p+1)/2) S is positive definite
    const int hp = (T.p - 1) >> 1;

    for (int64_t i = blockIdx.x * (int64_t)NS_WARPS + warp; i < a.NS; i += (int64_t)gridDim.x * NS_WARPS) {
        const int32_t* st = a.stencils + i * n;
        const int c0 = st[0];
        double xc[D], s[D], sx[D];
#pragma unroll
        for (int c = 0; c < D; ++c) xc[c] = a.X[(int64_t)c0 * D + c];
        // ---- scalestencil.jl:10-20: lane l owns stencil node l ----
        {
            const int id = lane < n ? st[lane] : c0;
            double mx[D];
#pragma unroll
            for (int c = 0; c < D; ++c) { sx[c] = a.X[(int64_t)id * D + c] - xc[c]; mx[c] = fabs(sx[c]); }
#pragma unroll
            for (int c = 0; c < D; ++c) {
                for (int o = 16; o > 0; o >>= 1) mx[c] = fmax(mx[c], __shfl_xor_sync(FULL, mx[c], o));
                s[c] = 1.0 / mx[c];
                sx[c] = sx[c] * s[c];
            }
        }
        double eta[D];
        bool eta_zero = true;
#pragma unroll
        for (int c = 0; c < D; ++c) { eta[c] = (a.Y[i * D + c] - xc[c]) * s[c]; eta_zero = eta_zero && (eta[c] == 0.0); }

        // ---- 1. column reduction of [P; g']: lane l < n holds row l of P, lanes o < nops also hold g_o' ----
        double prow[QP], grow[QP];
        const bool gl = n + nops <= 32;                           // warp-uniform: g rows fit into spare lanes
        const int go = gl ? lane - n : lane;                      // operator whose g row this lane owns
        const bool gown = go >= 0 && go < nops;
        {
            prow[0] = lane < n ? 1.0 : 0.0;
#pragma unroll
            for (int c = 1; c < QP; ++c) {
                double v = 0.0;
                if (c < q && lane < n) {
                    // mono[c] = mono[parent] * x[axis]; parents precede children, so a select chain over the
                    // already computed entries resolves the (warp-uniform) parent index
                    const int par = T.mpar[c], ax = T.maxis[c];
                    double pv = prow[0];
#pragma unroll
                    for (int u = 1; u < QP; ++u) if (u < c && u == par) pv = prow[u];
                    double xa = sx[0];
#pragma unroll
                    for (int d2 = 1; d2 < D; ++d2) if (d2 == ax) xa = sx[d2];
                    v = pv * xa;
                }
                prow[c] = v;
            }
            // g rows: lane c evaluates the polynomial right-hand side of monomial c for every operator (the table
            // walk is per monomial), staged through shared memory so that the row of operator o lands in ONE lane.
            // With room in the warp (n + nops <= 32) that lane is n + o and g' rides along as an extra row of P.
            for (int o = 0; o < nops; ++o) {
                double v = 0.0;
                if (lane < q) v = eta_zero ? rhs_poly_entry_at_zero<D>(T, o, lane, s) : rhs_poly_entry<D>(T, o, lane, eta, s);
                if (lane < QP) WpT[o * QP + lane] = v;           // scratch use of the w_p tile: [op][c]
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < QP; c += 2) {
                const double2 v = gown ? *reinterpret_cast<const double2*>(WpT + go * QP + c) : make_double2(0.0, 0.0);
                if (gl) { if (gown) { prow[c] = v.x; prow[c + 1] = v.y; } grow[c] = 0.0; grow[c + 1] = 0.0; }
                else { grow[c] = v.x; grow[c + 1] = v.y; }
            }
            __syncwarp();
        }
        bool ok = true;
        bool basic = false;
        int mybasic = 0;
#pragma unroll
        for (int j = 0; j < QP; ++j) {
            if (j < q) {
                const unsigned hi = (unsigned)__double2hiint(prow[j]) & 0x7fffffe0u;
                const unsigned key = (lane < n && !basic) ? (hi | (unsigned)lane) : 0u;
                const unsigned kmax = __reduce_max_sync(FULL, key);
                if (kmax < 32u) ok = false;                         // P is rank deficient on this stencil
                const int pl = kmax & 31;
                if (lane == pl) {
#pragma unroll
                    for (int c = 0; c < QP; c += 2) *reinterpret_cast<double2*>(Pr + c) = make_double2(prow[c], prow[c + 1]);
                    basic = true;
                    mybasic = j;
                }
                __syncwarp();
                double pr[QP];
#pragma unroll
                for (int c = 0; c < QP; c += 2) {
                    const double2 v = *reinterpret_cast<const double2*>(Pr + c);
                    pr[c] = v.x;
                    pr[c + 1] = v.y;
                }
                __syncwarp();
                const double rinv = fast_rcp(pr[j]);
                const double tl = prow[j] * rinv;
#pragma unroll
                for (int c = 0; c < QP; ++c)
                    if (c != j) prow[c] = fma(-tl, pr[c], prow[c]);
                prow[j] = tl;
                if (!gl) {                                        // g rows kept in the second register set
                    const double tg = grow[j] * rinv;
#pragma unroll
                    for (int c = 0; c < QP; ++c)
                        if (c != j) grow[c] = fma(-tg, pr[c], grow[c]);
                    grow[j] = tg;
                }
            }
        }
        // positions: non-basic nodes first (0..nb-1, in stencil order), then the basic ones in pivot order
        const unsigned nbmask = __ballot_sync(FULL, lane < n && !basic);
        const int pos = lane < n ? (basic ? nb + mybasic : __popc(nbmask & ((1u << lane) - 1u))) : 31;
        // ---- 2. stage W', w_p, permuted coordinates; zero the tiles that are read with padding ----
        for (int e = lane; e < C::G; e += 32) G[e] = 0.0;
        for (int e = lane; e < C::WT; e += 32) Wt[e] = 0.0;
        for (int e = lane; e < C::WP + C::BT; e += 32) WpT[e] = 0.0;      // WpT and Bt are contiguous
        __syncwarp();
        if (lane < n) {
            perm[pos] = lane;
#pragma unroll
            for (int c = 0; c < D; ++c) Sc[pos * D + c] = sx[c];
            if (!basic) {
#pragma unroll
                for (int c = 0; c < QP; ++c) Wt[pos * QP + c] = prow[c];
            }
            // RBF part of the right-hand sides at this node (generate_operator.jl:123-154)
            double del[D];
            double r2 = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double dd = eta[c] - sx[c];
                del[c] = dd == 0.0 ? EPS : dd;
                r2 += del[c] * del[c];
            }
            const double r = fast_sqrt(r2);
            double rp2 = T.p >= 3 ? r : fast_rcp(r);
            for (int e = 1; e < hp; ++e) rp2 *= r2;
            const double rp = rp2 * r2, rp4 = rp2 * fast_rcp(r2);
            for (int o = 0; o < nops; ++o) Bt[pos * 8 + o] = rhs_rbf_entry_fast<D>(T, o, del, s, r, r2, rp, rp2, rp4);
        }
        if (gown) {
#pragma unroll
            for (int c = 0; c < QP; ++c) WpT[c * 8 + go] = gl ? prow[c] : grow[c];
        }
        __syncwarp();
        // ---- 3. Phi~ in permuted order, by symmetric pairs ----
        // iteration tt pairs row tt with columns tt+1.. (lanes < n-1-tt) and row n-1-tt with columns lane+1 (the
        // other lanes): every iteration fills n-1 entries, all addresses advance by constants
        {
            const int half = (n + 1) >> 1;
            const double* __restrict__ Sr = Sc;
            double* __restrict__ Gw = G;
            for (int tt = 0; tt < half; ++tt) {
                const int i2 = n - 1 - tt;
                const bool first = lane < i2;
                const int ia = first ? tt : i2;
                const int ib = first ? tt + 1 + lane : lane + 1;
                if (lane < n - 1 && (first || i2 != tt)) {
                    double r2 = 0.0;
#pragma unroll
                    for (int c = 0; c < D; ++c) { const double dd = Sr[ia * D + c] - Sr[ib * D + c]; r2 = fma(dd, dd, r2); }
                    double v = fast_sqrt(r2);
                    for (int e = 0; e < hp; ++e) v *= r2;
                    Gw[ia * LD + ib] = v;
                    Gw[ib * LD + ia] = v;
                }
            }
        }
        __syncwarp();
        // ---- 4. Y = Phi~[:, N] - Phi~[:, B] W   (4 x 4 tiles; tile column 3 = right-hand sides with w_p) ----
        double c[4][4][2];
        {
            const double* gl = G + g * LD + 2 * t;
            const double* bl = Bt + g * 8 + 2 * t;
#pragma unroll
            for (int I = 0; I < 4; ++I) {
#pragma unroll
                for (int J = 0; J < 3; ++J) {
                    c[I][J][0] = gl[8 * I * LD + 8 * J];
                    c[I][J][1] = gl[8 * I * LD + 8 * J + 1];
                }
                c[I][3][0] = bl[8 * I * 8];
                c[I][3][1] = bl[8 * I * 8 + 1];
            }
            // columns >= nb of the first three tile columns belong to basic nodes: they are not part of Phi~[:, N]
#pragma unroll
            for (int I = 0; I < 4; ++I)
#pragma unroll
                for (int J = 0; J < 3; ++J)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (8 * J + 2 * t + e >= nb) c[I][J][e] = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double af[4], bf[4];
#pragma unroll
                for (int I = 0; I < 4; ++I) {
                    const int col = nb + 4 * k + t;
                    af[I] = (4 * k + t < q) ? -G[(8 * I + g) * LD + col] : 0.0;
                }
#pragma unroll
                for (int J = 0; J < 3; ++J) bf[J] = Wt[(8 * J + g) * QP + 4 * k + t];
                bf[3] = WpT[(4 * k + t) * 8 + g];
#pragma unroll
                for (int I = 0; I < 4; ++I)
#pragma unroll
                    for (int J = 0; J < 4; ++J) dmma884n(c[I][J][0], c[I][J][1], af[I], bf[J]);
            }
        }
        // ---- 5. [S | t] = Y[N, :] - W' Y[B, :] ----
        __syncwarp();                                     // Phi~ is dead: the Y tile reuses its storage
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J < 4; ++J)
                *reinterpret_cast<double2*>(Yb + (8 * I + g) * US + 8 * J + 2 * t) = make_double2(c[I][J][0], c[I][J][1]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double af[3], bf[4];
#pragma unroll
            for (int I = 0; I < 3; ++I) af[I] = -Wt[(8 * I + g) * QP + 4 * k + t];
#pragma unroll
            for (int J = 0; J < 4; ++J) bf[J] = (4 * k + t < q) ? Yb[(nb + 4 * k + t) * US + 8 * J + g] : 0.0;
#pragma unroll
            for (int I = 0; I < 3; ++I)
#pragma unroll
                for (int J = 0; J < 4; ++J) dmma884n(c[I][J][0], c[I][J][1], af[I], bf[J]);
        }
        __syncwarp();
        // identity padding outside the nb x nb block
#pragma unroll
        for (int I = 0; I < 3; ++I) {
            const int row = 8 * I + g;
#pragma unroll
            for (int J = 0; J < 3; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 8 * J + 2 * t + e;
                    if (row >= nb || col >= nb) c[I][J][e] = row == col ? sgn : 0.0;
                }
            if (row >= nb) { c[I][3][0] = 0.0; c[I][3][1] = 0.0; }
        }
        // ---- 6. blocked Gauss-Jordan WITHOUT pivoting on the definite S: the block step of weights_fast.cu with
        //         static pivot rows (row 4kb+s), so no pivot search, no row selects and a static pivot-row dump ----
        constexpr int PS6 = 36, US6 = 36;                 // strides == 4 (mod 16): conflict-free fragment loads
        double* Pbuf = G;                                 // [4][PS6]   (the Y tile is dead)
        double* Lbuf = Pbuf + 4 * PS6;                    // [4][PS6]
        double* Ubuf = Lbuf + 4 * PS6;                    // [4][US6]
        double* rinv_s = Ubuf + 4 * US6;                  // [24]
        double* const pb_w = Pbuf + (2 * (t & 1)) * PS6 + g;
        const double* const lb_r = Lbuf + t * PS6 + g;
        const double* const ub_r = Ubuf + t * US6 + g;
#pragma unroll
        for (int kb = 0; kb < 6; ++kb) {
            if (4 * kb < nb) {                            // warp-uniform: identity-padded block steps are skipped
                const int Jp = kb >> 1, h = kb & 1;
                if ((t >> 1) == h) {
#pragma unroll
                    for (int I = 0; I < 3; ++I) {
                        pb_w[8 * I] = c[I][Jp][0];
                        pb_w[PS6 + 8 * I] = c[I][Jp][1];
                    }
                }
                __syncwarp();
                double av[4], w[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) { av[cc] = lane < NS_NB ? Pbuf[cc * PS6 + lane] : 0.0; w[cc] = 0.0; }
#pragma unroll
                for (int sidx = 0; sidx < 4; ++sidx) {
                    const int pr = 4 * kb + sidx;         // static pivot row == lane pr
                    double pv[4], wp[4];
#pragma unroll
                    for (int cc = sidx; cc < 4; ++cc) pv[cc] = __shfl_sync(FULL, av[cc], pr);
#pragma unroll
                    for (int cc = 0; cc < sidx; ++cc) wp[cc] = __shfl_sync(FULL, w[cc], pr);
                    if (!(sgn * pv[sidx] > 0.0)) ok = false;      // S not definite: the pivoted kernel must take over
                    const double rinv = fast_rcp(pv[sidx]);
                    if (lane == 0) rinv_s[pr] = rinv;
                    const double nl = lane == pr ? 0.0 : av[sidx] * (-rinv);
#pragma unroll
                    for (int cc = sidx + 1; cc < 4; ++cc) av[cc] = fma(nl, pv[cc], av[cc]);
#pragma unroll
                    for (int cc = 0; cc < sidx; ++cc) w[cc] = fma(nl, wp[cc], w[cc]);
                    w[sidx] = nl;
                }
                if (lane < NS_NB) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) Lbuf[cc * PS6 + lane] = w[cc];
                }
                // raw pivot rows 4kb .. 4kb+3: tile row kb>>1, lanes with g>>2 == kb&1
                const int jlo = h == 0 ? Jp : Jp + 1;
                if ((g >> 2) == h) {
                    double2* dst = reinterpret_cast<double2*>(Ubuf + (g & 3) * US6 + 2 * t);
#pragma unroll
                    for (int J = 0; J < 4; ++J)
                        if (J >= jlo) dst[4 * J] = make_double2(c[Jp][J][0], c[Jp][J][1]);
                }
                __syncwarp();
                double af[3];
#pragma unroll
                for (int I = 0; I < 3; ++I) af[I] = lb_r[8 * I];
#pragma unroll
                for (int J = 0; J < 4; ++J) {
                    if (J >= jlo) {
                        const double bf = ub_r[8 * J];
#pragma unroll
                        for (int I = 0; I < 3; ++I) dmma884n(c[I][J][0], c[I][J][1], af[I], bf);
                    }
                }
                __syncwarp();
            }
        }
        // y = RHS_row / pivot_row
        __syncwarp();
#pragma unroll
        for (int I = 0; I < 3; ++I) {
            const int row = 8 * I + g;
            if (row < nb) {
                const double ri = rinv_s[row];
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (2 * t + e < nops) Ys[(2 * t + e) * NS_NB + row] = c[I][3][e] * ri;
            }
        }
        __syncwarp();
        // ---- 7. w[N] = y, w[B] = w_p - W y; rescale and scatter into the CSR row (generate_operator.jl:161-182) ----
        for (int o = 0; o < nops; ++o) {
            const double f = op_post_factor<D>(T, o, s);
            double* vrow = a.vals + ((int64_t)o * a.M + i) * n;
            double wv = 0.0;
            if (lane < nb) wv = Ys[o * NS_NB + lane];
            else if (lane < n) {
                const int cc = lane - nb;
                double acc = WpT[cc * 8 + o];
                for (int aa = 0; aa < nb; ++aa) acc = fma(-Wt[aa * QP + cc], Ys[o * NS_NB + aa], acc);
                wv = acc;
            }
            if (lane < n) vrow[perm[lane]] = ok ? f * wv : nan("");
        }
        for (int j = lane; j < n; j += 32) a.colind[i * n + j] = st[j];
        if (!ok && lane == 0) *a.redo = 1;
        __syncwarp();
    }
}

template <int D>
int launch_ns(rbffd_context* ctx, const NArgs& a) {
    using C = NsCfg<D>;
    const size_t smem = (size_t)C::BYTES_PER_WARP * NS_WARPS;
    if ((int64_t)smem > ctx->max_smem_optin) return RBFFD_ERR_UNSUPPORTED;
    auto kern = weights_ns_kernel<D>;
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks_needed = (a.NS + NS_WARPS - 1) / NS_WARPS;
    const int grid = (int)std::min<int64_t>(blocks_needed, (int64_t)ctx->sm_count * 3 * 8);
    kern<<<grid, NS_WARPS * 32, smem, ctx->stream>>>(a);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

}  // namespace

// Null-space fast path.  Returns RBFFD_ERR_UNSUPPORTED when the configuration is outside its scope, or when any stencil
// failed its definiteness / rank check (the caller then runs the pivoted Gauss-Jordan kernel over the batch).
int rbffd_weights_ns(rbffd_context* ctx, const OpTables& T, const double* X, int64_t NS, const double* Y, int64_t M,
                     const int32_t* stencils, int32_t* colind_out, double* vals_out, int* fail_flag) {
    const int nb = T.n - T.q;
    if (T.nops > 8 || T.n > 32 || T.q > NS_QP || nb < 1 || nb > NS_NB || T.dim < 2 || NS != M) return RBFFD_ERR_UNSUPPORTED;
    // conditional definiteness needs polynomial degree >= (p-1)/2: q >= C((p-1)/2 + d, d)
    {
        int need = 1;
        const int deg = (T.p - 1) / 2;
        for (int tq = 1; tq <= T.dim; ++tq) need = need * (deg + tq) / tq;
        if (T.q < need) return RBFFD_ERR_UNSUPPORTED;
    }
    NArgs a;
    a.X = X; a.Y = Y; a.stencils = stencils; a.NS = NS; a.M = M;
    a.colind = colind_out; a.vals = vals_out; a.fail = fail_flag; a.T = T;
    DevBuf<int> redo;
    CUDA_TRY(ctx, redo.alloc(1, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(redo.p, 0, sizeof(int), ctx->stream));
    a.redo = redo.p;
    int rc = T.dim == 2 ? launch_ns<2>(ctx, a) : launch_ns<3>(ctx, a);
    if (rc != RBFFD_OK) return rc;
    int h_redo = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&h_redo, redo.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return h_redo ? RBFFD_ERR_UNSUPPORTED : RBFFD_OK;
}
