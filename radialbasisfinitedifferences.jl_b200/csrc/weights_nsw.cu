// weights_nsw.cu -- multi-warp null-space weight kernel (K4) for the larger stencils of BASELINE configs 1, 3, 4, 5
// (32 < n <= 64: n = 42/50/60, q = 21/15/20).  Same mathematics as weights_ns.cu (see its header): column reduction of
// P with pivoting (q steps), S = Z' Phi Z formed by FP64 DMMAs, unpivoted blocked Gauss-Jordan on the definite S,
// w[N] = y, w[B] = w_p - W y.  One CTA of 4 warps owns one stencil at a time:
//   phase A  warps 0-1: one row of [P; g'] per lane (64 rows), q pivoted reduction steps; the two warp winners publish
//            their candidate rows through shared memory, ONE named barrier (64 threads) per step;
//            warps 2-3 meanwhile assemble Phi (original stencil order, circulant pairing) and the RBF right-hand sides;
//   phase B  Y = Phi~[:,N] - Phi~[:,B] W : 8 row tiles, 2 per warp; the permutation to (non-basic | basic) order is
//            applied while gathering the fragments (rows / columns through perm[]);
//   phase C  [S | t] = Y[N,:] - W' Y[B,:] : tile COLUMNS dealt cyclically to the warps (column J in warp J % 4), the
//            layout the elimination wants;
//   phase D  blocked Gauss-Jordan without pivoting, 4 columns per step: the owner of the panel column eliminates the
//            48 x 4 panel (two rows per lane), every warp dumps the raw pivot rows of its columns, ONE __syncthreads,
//            every warp applies the rank-4 update to its columns with DMMAs (buffers alternate by step parity);
//   phase E  w[B] = w_p - W y, rescale, scatter into the CSR row.
// A non-definite S, a rank-deficient P or a non-finite weight raises the redo flag: the batch is then redone by the
// pivoted Gauss-Jordan kernel (weights_mw.cu).
// Scope: collocated rows, 8 <= n <= 64, n + nops <= 64, n - q <= 48, q in the instantiated list, polydeg >= (p-1)/2.
// Replaces the same reference lines as weights.cu (scalestencil.jl:10-20, interpolationmatrix.jl:5-8,
// generate_operator.jl:55-65,89-182, hyperviscosity_operator.jl:97-161).
#include <utility>

#include <cstdlib>
#include "common.cuh"
#include "tables.cuh"
#include "phs.cuh"

namespace {

struct WNArgs {
    const double* X;
    const double* Y;
    const int32_t* stencils;   // [NX][n]
    const int32_t* center;     // [M] stencil of row i (Y != X), or null: row i uses stencil i
    int64_t NS, M;
    int32_t* colind;           // [M][n]
    double* vals;              // [nops][M][n]
    int* redo;                 // set to 1 when any stencil needs the pivoted fallback
    int32_t gzcol[8];          // polynomial right-hand side at eta == 0: DERIV operator o hits exactly one monomial
    double gzval[8];
    int32_t lapcol[3];         // column of x_a^2 (-1: degree < 2)
    int32_t bs;                // row stride of the RBF right-hand-side tile
    int32_t rot;               // 1: warp roles rotate with the CTA index
    OpTables T;
};

__device__ __forceinline__ void dmma884w(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double rcp3w(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double u = fma(e, e, e);
    return fma(y, u, y);
}
__device__ __forceinline__ double warp_max_nn(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned hmax = __reduce_max_sync(0xffffffffu, hi);
    const unsigned lmax = __reduce_max_sync(0xffffffffu, hi == hmax ? lo : 0u);
    return __hiloint2double((int)hmax, (int)lmax);
}
__device__ __forceinline__ void bar_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// graded monomial table of build_op_tables (weights.cu) in closed form: mono[c] = mono[par(c)] * x[axis(c)]
__host__ __device__ constexpr int mono_par(int D, int c) {
    if (c == 0) return 0;
    if (D == 2) {
        int g = 0;
        while ((g + 1) * (g + 2) / 2 <= c) ++g;
        const int j = c - g * (g + 1) / 2;
        return (g - 1) * g / 2 + (j > 0 ? j - 1 : 0);
    }
    int g = 0;
    while ((g + 1) * (g + 2) * (g + 3) / 6 <= c) ++g;
    const int rem = c - g * (g + 1) * (g + 2) / 6;
    int t = 0;
    while ((t + 1) * (t + 2) / 2 <= rem) ++t;
    const int u = rem - t * (t + 1) / 2;
    const int base = (g - 1) * g * (g + 1) / 6;
    if (u > 0) return base + (t - 1) * t / 2 + (u - 1);
    if (t > 0) return base + (t - 1) * t / 2;
    return base;
}
__host__ __device__ constexpr int mono_axis(int D, int c) {
    if (c == 0) return 0;
    if (D == 2) {
        int g = 0;
        while ((g + 1) * (g + 2) / 2 <= c) ++g;
        return c - g * (g + 1) / 2 > 0 ? 1 : 0;
    }
    int g = 0;
    while ((g + 1) * (g + 2) * (g + 3) / 6 <= c) ++g;
    const int rem = c - g * (g + 1) * (g + 2) / 6;
    int t = 0;
    while ((t + 1) * (t + 2) / 2 <= rem) ++t;
    const int u = rem - t * (t + 1) / 2;
    return u > 0 ? 2 : (t > 0 ? 1 : 0);
}
template <int D, int C>
struct MonoStep {
    static constexpr int par = mono_par(D, C), axis = mono_axis(D, C);
};
template <int D, int... C>
__device__ __forceinline__ void mono_rows(const double* x, double* m, std::integer_sequence<int, C...>) {
    ((m[C] = (C == 0) ? 1.0 : m[MonoStep<D, C>::par] * x[MonoStep<D, C>::axis]), ...);
}

// Unpivoted elimination of one 48 x 4 panel of the definite matrix S (static pivot rows pr0 .. pr0+3), one warp, two rows
// per lane (rows lane and lane + 32).  Pbuf [4][PS] holds the panel by column; the transform columns W of the rank-4
// update X += W X[pivots, :] go to Lb [4][PS], the pivot reciprocals to rinv_s[pr0 ..].  Pivot rows are not scaled
// (y = RHS_row / pivot at the end), so their own entry of W is zero.  Returns the sign-violation bits of the pivots.
// Deliberately not inlined: it is called once per block step and would otherwise be replicated 12 times.
__device__ __forceinline__ int gj_panel48(const double* __restrict__ Pbuf, double* __restrict__ Lb, double* __restrict__ rinv_s,
                                       int pr0, int sgnbits, int PS) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int sl = pr0 >> 5;                              // slot of the four pivot rows (pr0 is a multiple of 4)
    double av[2][4], w[2][4];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            av[rr][cc] = (rr == 0 || lane < 16) ? Pbuf[cc * PS + lane + 32 * rr] : 0.0;
            w[rr][cc] = 0.0;
        }
    int bad = 0;
#pragma unroll
    for (int sidx = 0; sidx < 4; ++sidx) {
        const int pl = (pr0 + sidx) & 31;
        double pv[4], wp[4];
#pragma unroll
        for (int cc = sidx; cc < 4; ++cc) pv[cc] = __shfl_sync(FULL, sl ? av[1][cc] : av[0][cc], pl);
#pragma unroll
        for (int cc = 0; cc < sidx; ++cc) wp[cc] = __shfl_sync(FULL, sl ? w[1][cc] : w[0][cc], pl);
        bad |= __double2hiint(pv[sidx]) ^ sgnbits;        // S not definite: the pivoted kernel must take over
        const double rinv = rcp3w(pv[sidx]);
        if (lane == 0) rinv_s[pr0 + sidx] = rinv;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const double nl = (lane == pl && rr == sl) ? 0.0 : av[rr][sidx] * (-rinv);
#pragma unroll
            for (int cc = sidx + 1; cc < 4; ++cc) av[rr][cc] = fma(nl, pv[cc], av[rr][cc]);
#pragma unroll
            for (int cc = 0; cc < sidx; ++cc) w[rr][cc] = fma(nl, wp[cc], w[rr][cc]);
            w[rr][sidx] = nl;
        }
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        Lb[cc * PS + lane] = w[0][cc];
        if (lane < 16) Lb[cc * PS + lane + 32] = w[1][cc];
    }
    return bad;
}

#ifdef NSW_TIMING
// development aid (make EXTRA=-DNSW_TIMING): wall cycles per phase, summed over stencils (thread 0 of warps 0 and 2)
__device__ unsigned long long nsw_prof[16];
#define NSW_T(slot, cond) do { if (cond) { const long long _t = clock64(); atomicAdd(&nsw_prof[slot], (unsigned long long)(_t - tprev)); tprev = _t; } } while (0)
#else
#define NSW_T(slot, cond) do { } while (0)
#endif

constexpr int WN_LD = 68;      // row stride of Phi (original stencil order); == 4 (mod 16): pair stores 69 l + k, 69 l + 68 k conflict-free
constexpr int WN_NBP = 48;     // padded null-space dimension (6 tiles)

template <int D, int Q, bool FOLD>
struct WnCfg {
    static constexpr int KS = (Q + 3) / 4, QP = 4 * KS;
    static constexpr int NJ = FOLD ? 6 : 7;              // tile columns of [S | t]
    static constexpr int US = ((8 * NJ + 15) & ~15) + 4;  // row stride of the Y tile, == 4 (mod 16): Y_B fragment loads 4 t + g conflict-free
    static constexpr int DP = D == 2 ? 2 : 4;
    static constexpr int CS = (Q + 2) & ~1;              // candidate row: Q entries + the reciprocal of the pivot, even
    // doubles
    static constexpr int G = 64 * WN_LD;                 // Phi, later the Y tile, later the exchange buffers of phase D
    static constexpr int WT = 56 * QP;                   // W' rows by position, then the w_p rows
    static constexpr int SC = 64 * DP;
    static constexpr int CAND = 2 * 2 * CS;              // [parity][warp][CS]
    static_assert(64 * US <= G, "Y tile must fit into the Phi tile");
};

template <int D, int Q, bool FOLD>
__global__ void __launch_bounds__(128, 4) weights_nsw_kernel(WNArgs a) {
    using C = WnCfg<D, Q, FOLD>;
    constexpr int LD = WN_LD, KS = C::KS, QP = C::QP, NJ = C::NJ, US = C::US, DP = C::DP, CS = C::CS, NBP = WN_NBP;
    extern __shared__ __align__(16) unsigned char wsm[];
    // Warp ROLES rotate with the CTA index: a CTA's warp w sits on scheduler partition w of the SM, so without the rotation
    // the same role of all resident CTAs (the P reduction, the owner of the first panels) would share one partition.
    const int tid = threadIdx.x, lane = tid & 31, warp = ((tid >> 5) + a.rot * (int)blockIdx.x) & 3;
    const int g = lane >> 2, t = lane & 3;
#ifdef NSW_TIMING
    const int rtid = warp * 32 + lane;
#endif
    const OpTables& T = a.T;
    const int n = T.n, nops = T.nops, nb = n - Q, BS = a.bs;
    double* G = reinterpret_cast<double*>(wsm);
    double* Yb = G;
    double* Wt = G + C::G;
    double* Sc = Wt + C::WT;
    double* cand = Sc + C::SC;
    double* Bt = cand + C::CAND;                      // [64][BS] RBF right-hand sides, original order
    double* Ys = Bt;                                  // solution y, [op][48] (Bt is dead by then; 8*48 <= 64*BS needs BS >= 6 or nops <= BS)
    double* pf = Bt + 64 * BS;                        // [8] chain-rule factor of every operator
    unsigned* candkey = reinterpret_cast<unsigned*>(pf + 8);           // [2][2]
    int* perm = reinterpret_cast<int*>(candkey + 4);                   // [64] position -> original stencil slot
    int* cnt0 = perm + 64;                                             // non-basic rows held by warp 0
    const double EPS = 2.220446049250313e-16;
    const unsigned FULL = 0xffffffffu;
    const double sgn = (((T.p + 1) >> 1) & 1) ? -1.0 : 1.0;           // (-1)^((p+1)/2) S is positive definite
    const int sgnbits = sgn < 0.0 ? (int)0x80000000 : 0;
    const int hp = (T.p - 1) >> 1;
    const int rc0 = (nb + 3) & ~3;
    const int rcb = FOLD ? rc0 : NBP;                 // first right-hand-side column; also the row of w_p in Wt
    const int L2 = (warp & 1) * 32 + lane;            // the stencil node / row of [P; g'] this thread works on in phase A
    // exchange buffers of phase D (alias the Y tile), alternating by block-step parity
    constexpr int PS = 52, UST = 8 * NJ + 4;          // strides == 4 (mod 16)
    double* Pbuf = G;                                 // [4][PS]            owner-private
    double* Lbuf = Pbuf + 4 * PS;                     // [2][4][PS]
    double* Ubuf = Lbuf + 2 * 4 * PS;                 // [2][4][UST]
    double* rinv_s = Ubuf + 2 * 4 * UST;              // [48]

    for (int64_t i = blockIdx.x; i < a.NS; i += gridDim.x) {
        // ---- 0. scalestencil.jl:10-20: every warp reduces the whole stencil (identical s in all warps) ----
#ifdef NSW_TIMING
        long long tprev = clock64();
#endif
        const int32_t* st = a.stencils + (a.center ? (int64_t)a.center[i] : i) * n;
        const int id0 = st[lane < n ? lane : 0], id1 = st[lane + 32 < n ? lane + 32 : 0];
        double sx[D], s[D], eta[D];
        bool eta_zero = true;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const double x0 = a.X[(int64_t)id0 * D + c], x1 = a.X[(int64_t)id1 * D + c];
            const double xc = __shfl_sync(FULL, x0, 0);
            const double d0 = x0 - xc, d1 = x1 - xc;
            s[c] = 1.0 / warp_max_nn(fmax(fabs(d0), fabs(d1)));
            sx[c] = ((warp & 1) ? d1 : d0) * s[c];
            eta[c] = (a.Y[i * D + c] - xc) * s[c];
            eta_zero = eta_zero && (eta[c] == 0.0);
        }
        unsigned kmin = 0xffffffffu;
        int bad = 0;
        NSW_T(0, rtid == 0);
        NSW_T(9, rtid == 64);
        if (warp < 2) {
            // ---- A1. column reduction of [P; g']: thread L2 < n holds row L2 of P, thread n + o the row g_o' ----
            double prow[Q];
            mono_rows<D>(sx, prow, std::make_integer_sequence<int, Q>{});
            const int go = L2 - n;
            const bool gown = go >= 0 && go < nops;
            if (L2 >= n) {
#pragma unroll
                for (int c = 0; c < Q; ++c) prow[c] = 0.0;
            }
            if (eta_zero) {
                if (gown) {
                    const bool lap = T.kind[go] == RBFFD_OP_LAPLACE;
                    const int col = a.gzcol[go];
                    const double val = a.gzval[go];
#pragma unroll
                    for (int c = 0; c < Q; ++c) {
                        double v = (c == col) ? val : 0.0;
#pragma unroll
                        for (int ax = 0; ax < D; ++ax)
                            if (lap && c == a.lapcol[ax]) v = 2.0 * s[ax] * s[ax];
                        prow[c] = v;
                    }
                }
            } else {
                // general evaluation point: thread c evaluates monomial c for every operator, staged through shared memory
                for (int o = 0; o < nops; ++o)
                    if (L2 < Q) Wt[o * QP + L2] = rhs_poly_entry<D>(T, o, L2, eta, s);
                bar_named(1, 64);
                if (gown) {
#pragma unroll
                    for (int c = 0; c < Q; ++c) prow[c] = Wt[go * QP + c];
                }
                bar_named(1, 64);
            }
            bool basic = false;
            int mybasic = 0;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                const unsigned hi = (unsigned)__double2hiint(prow[j]) & 0x7fffffc0u;
                const unsigned key = (L2 < n && !basic) ? (hi | (unsigned)L2) : 0u;
                const unsigned kw = __reduce_max_sync(FULL, key);
                const double rown = rcp3w(prow[j]);             // every candidate inverts its own entry under the search
                double* cw = cand + ((j & 1) * 2 + warp) * CS;
                if (lane == 0) candkey[(j & 1) * 2 + warp] = kw;
                if (key == kw && kw >= 64u) {                   // the warp's winner publishes its row and 1/pivot
                    double2* dst = reinterpret_cast<double2*>(cw);
#pragma unroll
                    for (int c = 0; c < CS; c += 2)
                        dst[c >> 1] = make_double2(c < Q ? prow[c] : rown, c + 1 < Q ? prow[c + 1] : rown);
                }
                bar_named(1, 64);
                const unsigned k0 = candkey[(j & 1) * 2], k1 = candkey[(j & 1) * 2 + 1];
                const unsigned kmax = max(k0, k1);
                kmin = min(kmin, kmax);                         // < 64: P is rank deficient on this stencil
                const double2* cr = reinterpret_cast<const double2*>(cand + ((j & 1) * 2 + (k1 > k0 ? 1 : 0)) * CS);
                if (L2 == (int)(kmax & 63u) && kmax >= 64u) { basic = true; mybasic = j; }
                double pr[CS];
#pragma unroll
                for (int c = 0; c < CS; c += 2) { const double2 v = cr[c >> 1]; pr[c] = v.x; pr[c + 1] = v.y; }
                const double tl = prow[j] * pr[Q];
#pragma unroll
                for (int c = 0; c < Q; ++c)
                    if (c != j) prow[c] = fma(-tl, pr[c], prow[c]);
                prow[j] = tl;
            }
            // positions: non-basic nodes first (0..nb-1, in stencil order), then the basic ones in pivot order
            const unsigned nbmask = __ballot_sync(FULL, L2 < n && !basic);
            if (warp == 0 && lane == 0) *cnt0 = __popc(nbmask);
            bar_named(1, 64);
            const int before = __popc(nbmask & ((1u << lane) - 1u)) + (warp == 1 ? *cnt0 : 0);
            const int pos = L2 < n ? (basic ? nb + mybasic : before) : -1;
            if (L2 < n) perm[pos] = L2; else perm[L2] = 0;      // positions n..63 read node 0 (rows / columns nobody uses)
            const int wrow = (L2 < n && !basic) ? pos : (gown ? rcb + go : -1);
            if (wrow >= 0) {
                double2* dst = reinterpret_cast<double2*>(Wt + wrow * QP);
#pragma unroll
                for (int c = 0; c < QP; c += 2) dst[c >> 1] = make_double2(c < Q ? prow[c] : 0.0, c + 1 < Q ? prow[c + 1] : 0.0);
            }
            NSW_T(1, rtid == 0);
        } else {
            // ---- A2. Phi in original stencil order by symmetric pairs (round k pairs node l with (l + k) mod n), and
            //          the RBF part of the right-hand sides (generate_operator.jl:123-154) ----
            const int l = L2 < n ? L2 : 0;
            if (L2 < n) {
#pragma unroll
                for (int c = 0; c < D; ++c) Sc[L2 * DP + c] = sx[c];
                G[L2 * LD + L2] = 0.0;
                double del[D];
                double r2 = 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const double dd = eta[c] - sx[c];
                    del[c] = dd == 0.0 ? EPS : dd;
                    r2 = fma(del[c], del[c], r2);
                }
                const double y = phs_rsqrt(r2);
                double rp4 = y;                                 // r^(p-4)
                for (int e = 1; e < hp; ++e) rp4 *= r2;
                const double rp2 = rp4 * r2, rp = rp2 * r2, r = r2 * y;
                for (int o = 0; o < nops; ++o) Bt[L2 * BS + o] = rhs_rbf_entry_fast<D>(T, o, del, s, r, r2, rp, rp2, rp4);
            }
            bar_named(2, 64);
            phs_assemble<D, DP, LD>(Sc, G, sx, l, n, L2 < n, hp);
            NSW_T(2, rtid == 64);
        }
        __syncthreads();
        NSW_T(3, rtid == 0);
        NSW_T(4, rtid == 64);
        // ---- B. Y = Phi~[:, N] - Phi~[:, B] W : row tiles 2*warp, 2*warp+1; permutation applied while gathering ----
        {
            double cy[2][NJ][2];
            const double* rowp[2];
            const double* browp[2];
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
                const int orig = perm[8 * (2 * warp + ii) + g];
                rowp[ii] = G + orig * LD;
                browp[ii] = Bt + orig * BS;
            }
#pragma unroll
            for (int J = 0; J < NJ; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 8 * J + 2 * t + e;
                    const bool isn = col < nb, isr = col >= rcb && col < rcb + nops;
                    const int pc = perm[isn ? col : 0];
#pragma unroll
                    for (int ii = 0; ii < 2; ++ii) cy[ii][J][e] = isn ? rowp[ii][pc] : (isr ? browp[ii][col - rcb] : 0.0);
                }
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const bool kin = 4 * k + t < Q;
                const int pc = perm[kin ? nb + 4 * k + t : 0];
                double af[2], bf[NJ];
#pragma unroll
                for (int ii = 0; ii < 2; ++ii) af[ii] = kin ? -rowp[ii][pc] : 0.0;
#pragma unroll
                for (int J = 0; J < NJ; ++J) bf[J] = Wt[(8 * J + g) * QP + 4 * k + t];
#pragma unroll
                for (int J = 0; J < NJ; ++J)
#pragma unroll
                    for (int ii = 0; ii < 2; ++ii) dmma884w(cy[ii][J][0], cy[ii][J][1], af[ii], bf[J]);
            }
            __syncthreads();                              // Phi is dead in every warp: the Y tile reuses its storage
#pragma unroll
            for (int J = 0; J < NJ; ++J)
#pragma unroll
                for (int ii = 0; ii < 2; ++ii)
                    *reinterpret_cast<double2*>(Yb + (8 * (2 * warp + ii) + g) * US + 8 * J + 2 * t) = make_double2(cy[ii][J][0], cy[ii][J][1]);
        }
        __syncthreads();
        NSW_T(5, rtid == 0);
        // ---- C. [S | t] = Y[N, :] - W' Y[B, :] : this warp owns tile columns warp and warp + 4 ----
        double c[6][2][2];
        const bool has2 = warp + 4 < NJ;
        {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int J = warp + 4 * jj;
#pragma unroll
                for (int I = 0; I < 6; ++I) {
                    const double2 v = (jj == 0 || has2) ? *reinterpret_cast<const double2*>(Yb + (8 * I + g) * US + 8 * J + 2 * t) : make_double2(0.0, 0.0);
                    c[I][jj][0] = v.x; c[I][jj][1] = v.y;
                }
            }
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const bool kin = 4 * k + t < Q;
                double af[6], bf[2];
#pragma unroll
                for (int I = 0; I < 6; ++I) af[I] = -Wt[(8 * I + g) * QP + 4 * k + t];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) bf[jj] = (kin && (jj == 0 || has2)) ? Yb[(nb + 4 * k + t) * US + 8 * (warp + 4 * jj) + g] : 0.0;
#pragma unroll
                for (int I = 0; I < 6; ++I) dmma884w(c[I][0][0], c[I][0][1], af[I], bf[0]);
                if (has2) {
#pragma unroll
                    for (int I = 0; I < 6; ++I) dmma884w(c[I][1][0], c[I][1][1], af[I], bf[1]);
                }
            }
            // identity padding outside the nb x nb block; right-hand-side columns of padded rows are zero
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int J = warp + 4 * jj;
#pragma unroll
                for (int I = 0; I < 6; ++I) {
                    if (8 * I + 8 > nb || 8 * J + 8 > nb) {
                        const int row = 8 * I + g;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int col = 8 * J + 2 * t + e;
                            if (col >= rcb) { if (row >= nb) c[I][jj][e] = 0.0; }
                            else if (row >= nb || col >= nb) c[I][jj][e] = row == col ? sgn : 0.0;
                        }
                    }
                }
            }
        }
        __syncthreads();                                  // the Y tile is dead: its storage becomes the exchange buffers
        NSW_T(6, rtid == 0);
        // ---- D. blocked Gauss-Jordan WITHOUT pivoting on the definite S (static pivot rows 4kb .. 4kb+3), with one
        //         step of lookahead: in step kb the owner of panel kb+1 updates that tile column first and eliminates the
        //         next panel while the other warps are still applying update kb ----
        auto dump_panel = [&](auto JPc, auto Hc) {        // the 4 panel columns (tile column Jp, half h) of all 48 rows
            constexpr int Jp = decltype(JPc)::value, h = decltype(Hc)::value, jp = Jp >> 2;
            if ((t >> 1) == h) {
                double* pw = Pbuf + (2 * (t & 1)) * PS + g;
#pragma unroll
                for (int I = 0; I < 6; ++I) {
                    pw[8 * I] = c[I][jp][0];
                    pw[PS + 8 * I] = c[I][jp][1];
                }
            }
            __syncwarp();
        };
        auto dump_pivot_rows = [&](auto KBc) {            // raw pivot rows of step kb: tile row kb>>1, lanes with g>>2 == kb&1
            constexpr int kb = decltype(KBc)::value, Jp = kb >> 1, h = kb & 1, jlo = h == 0 ? Jp : Jp + 1;
            double* Ub = Ubuf + (kb & 1) * 4 * UST;
            if ((g >> 2) == h) {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int J = warp + 4 * jj;
                    if (J >= jlo && (jj == 0 || has2))
                        *reinterpret_cast<double2*>(Ub + (g & 3) * UST + 8 * J + 2 * t) = make_double2(c[Jp][jj][0], c[Jp][jj][1]);
                }
            }
        };
        if (warp == 0) {
            dump_panel(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
            bad |= gj_panel48(Pbuf, Lbuf, rinv_s, 0, sgnbits, PS);
        }
        dump_pivot_rows(std::integral_constant<int, 0>{});
        __syncthreads();
        auto gj_step = [&](auto KBc) {
            constexpr int kb = decltype(KBc)::value;
            constexpr int Jp = kb >> 1, h = kb & 1, jlo = h == 0 ? Jp : Jp + 1;
            constexpr int kn = kb + 1, Jn = kn >> 1, hn = kn & 1, jn = Jn >> 2;      // next panel: tile column Jn, local index jn
            const double* Lb = Lbuf + (kb & 1) * 4 * PS;
            const double* Ub = Ubuf + (kb & 1) * 4 * UST;
            const bool next = kn < 12 && 4 * kn < nb;
            double af[6];
#pragma unroll
            for (int I = 0; I < 6; ++I) af[I] = Lb[t * PS + 8 * I + g];
            auto update = [&](auto JJc) {
                constexpr int jj = decltype(JJc)::value;
                const int J = warp + 4 * jj;
                if (J >= jlo && (jj == 0 || has2)) {
                    const double bf = Ub[t * UST + 8 * J + g];
#pragma unroll
                    for (int I = 0; I < 6; ++I) dmma884w(c[I][jj][0], c[I][jj][1], af[I], bf);
                }
            };
            if (next && warp == (Jn & 3)) {
                update(std::integral_constant<int, jn>{});
                dump_panel(std::integral_constant<int, Jn>{}, std::integral_constant<int, hn>{});
                bad |= gj_panel48(Pbuf, Lbuf + (kn & 1) * 4 * PS, rinv_s, 4 * kn, sgnbits, PS);
                update(std::integral_constant<int, 1 - jn>{});
            } else {
                update(std::integral_constant<int, 0>{});
                update(std::integral_constant<int, 1>{});
            }
            if (next) dump_pivot_rows(std::integral_constant<int, (kn < 12 ? kn : 0)>{});
            __syncthreads();
        };
        [&]<int... KB>(std::integer_sequence<int, KB...>) {
            (([&] { if (4 * KB < nb) gj_step(std::integral_constant<int, KB>{}); }()), ...);
        }(std::make_integer_sequence<int, 12>{});
        if (bad < 0 || kmin < 64u) *a.redo = 1;
        __syncthreads();
        NSW_T(7, rtid == 0);                                  // every pivot reciprocal is published; Bt is dead (Ys aliases it)
        // y = RHS_row / pivot_row  (right-hand-side column rcb + o lives in tile (rcb + o) / 8)
        if (tid < nops) pf[tid] = op_post_factor<D>(T, tid, s);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int J = warp + 4 * jj;
#pragma unroll
            for (int I = 0; I < 6; ++I) {
                const int row = 8 * I + g;
                const double ri = rinv_s[row < nb ? row : 0];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int o = 8 * J + 2 * t + e - rcb;
                    if (row < nb && (jj == 0 || has2) && o >= 0 && o < nops) Ys[o * NBP + row] = c[I][jj][e] * ri;
                }
            }
        }
        __syncthreads();
        // ---- E. w[N] = y, w[B] = w_p - W y; rescale and scatter into the CSR row (generate_operator.jl:161-182) ----
        {
            bool fin = true;
            const double INF = __longlong_as_double(0x7ff0000000000000ll);
            // non-basic nodes: w = y
            for (int idx = tid; idx < nb * nops; idx += 128) {
                const int o = idx / nb, pos = idx - o * nb;
                const double wv = Ys[o * NBP + pos] * pf[o];
                fin = fin && (fabs(wv) < INF);              // a zero pivot shows up as a non-finite weight
                a.vals[((int64_t)o * a.M + i) * n + perm[pos]] = wv;
            }
            // basic nodes: the Q x nops block  w_p - W' y  as one DMMA row tile per warp (rows = basic nodes 8 warp + g,
            // columns = operators 2t, 2t+1, k = non-basic nodes)
            if (8 * warp < Q) {
                const int cc = 8 * warp + g;
                const bool rin = cc < Q;
                double c0 = (rin && 2 * t < nops) ? Wt[(rcb + 2 * t) * QP + (rin ? cc : 0)] : 0.0;
                double c1 = (rin && 2 * t + 1 < nops) ? Wt[(rcb + 2 * t + 1) * QP + (rin ? cc : 0)] : 0.0;
                const double* wcol = Wt + (rin ? cc : 0);
                const double* ycol = Ys + (g < nops ? g : 0) * NBP;
#pragma unroll 2
                for (int k0 = 0; k0 < nb; k0 += 4) {
                    const int aa = k0 + t;
                    const bool kin = aa < nb;
                    const double af = (rin && kin) ? -wcol[(kin ? aa : 0) * QP] : 0.0;
                    const double bf = (g < nops && kin) ? ycol[kin ? aa : 0] : 0.0;
                    dmma884w(c0, c1, af, bf);
                }
                if (rin) {
                    const int dst = perm[nb + cc];
                    if (2 * t < nops) {
                        const double wv = c0 * pf[2 * t];
                        fin = fin && (fabs(wv) < INF);
                        a.vals[((int64_t)(2 * t) * a.M + i) * n + dst] = wv;
                    }
                    if (2 * t + 1 < nops) {
                        const double wv = c1 * pf[2 * t + 1];
                        fin = fin && (fabs(wv) < INF);
                        a.vals[((int64_t)(2 * t + 1) * a.M + i) * n + dst] = wv;
                    }
                }
            }
            if (!fin) *a.redo = 1;
            if (warp == 0) {
                if (lane < n) a.colind[i * n + lane] = id0;
                if (lane + 32 < n) a.colind[i * n + lane + 32] = id1;
            }
        }
        __syncthreads();
        NSW_T(8, rtid == 0);
    }
}

template <int D, int Q>
bool wn_table_matches(const OpTables& T) {
    for (int c = 1; c < Q; ++c)
        if (T.mpar[c] != mono_par(D, c) || T.maxis[c] != mono_axis(D, c)) return false;
    return true;
}

template <int D, int Q, bool FOLD>
int launch_nsw2(rbffd_context* ctx, WNArgs& a) {
    using C = WnCfg<D, Q, FOLD>;
    a.bs = std::max(6, (a.T.nops + 1) & ~1);             // Ys [nops][48] aliases Bt [64][bs]
    const size_t smem = ((size_t)(C::G + C::WT + C::SC + C::CAND + 64 * a.bs + 8) * 8 + 4 * 4 + 64 * 4 + 16 + 15) & ~(size_t)15;
    if ((int64_t)smem > ctx->max_smem_optin) return RBFFD_ERR_UNSUPPORTED;
    auto kern = weights_nsw_kernel<D, Q, FOLD>;
    // development knobs: RBFFD_NSW_PAD_SMEM = extra dynamic shared memory per CTA (occupancy sweeps), RBFFD_NSW_ROT = 0/1
    static const int pad_smem = [] { const char* e = getenv("RBFFD_NSW_PAD_SMEM"); return e ? atoi(e) : 0; }();
    static const int rot = [] { const char* e = getenv("RBFFD_NSW_ROT"); return e ? atoi(e) : 0; }();   // measured: 1 costs 12 % at config 4 (r02a)
    a.rot = rot;
    const size_t smem_launch = smem + (size_t)std::max(0, pad_smem);
    if ((int64_t)smem_launch > ctx->max_smem_optin) return RBFFD_ERR_UNSUPPORTED;
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_launch));
    const int per_sm = std::max<int>(1, std::min<int>(4, (int)((228 * 1024) / (smem_launch + 1024))));
    // CTAs per resident slot (see weights_ns.cu): 4 -> 44.8 ms, 32 -> 43.9, 256 -> 41.3, one stencil per CTA -> 43.1 (config 4 shape)
    static const int waves = [] { const char* e = getenv("RBFFD_NSW_WAVES"); const int w = e ? atoi(e) : 0; return w > 0 ? w : 256; }();
    const int grid = (int)std::min<int64_t>(a.NS, (int64_t)ctx->sm_count * per_sm * waves);
#ifdef NSW_TIMING
    unsigned long long zero[16] = {};
    cudaMemcpyToSymbolAsync(nsw_prof, zero, sizeof(zero), 0, cudaMemcpyHostToDevice, ctx->stream);
#endif
    kern<<<grid, 128, smem_launch, ctx->stream>>>(a);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
#ifdef NSW_TIMING
    unsigned long long h[16];
    cudaMemcpyFromSymbol(h, nsw_prof, sizeof(h));
    static const char* nm[9] = {"0 load+scale", "A1 P reduction (warp 0)", "A2 Phi+rhs (warp 2)", "A wait (warp 0)", "A wait (warp 2)", "B Y",
                                "C S", "D elimination", "E back+store"};
    for (int k = 0; k < 9; ++k) fprintf(stderr, "[nsw timing] %-26s %9.0f cycles/stencil\n", nm[k], (double)h[k] / (double)a.NS);
#endif
    return RBFFD_OK;
}

template <int D, int Q>
int launch_nsw(rbffd_context* ctx, WNArgs& a) {
    if (!wn_table_matches<D, Q>(a.T)) return RBFFD_ERR_UNSUPPORTED;
    const int nb = a.T.n - Q;
    const bool fold = ((nb + 3) & ~3) + a.T.nops <= WN_NBP;
    return fold ? launch_nsw2<D, Q, true>(ctx, a) : launch_nsw2<D, Q, false>(ctx, a);
}

}  // namespace

// Multi-warp null-space path.  Returns RBFFD_ERR_UNSUPPORTED when the configuration is outside its scope, or when any
// stencil failed its definiteness / rank / finiteness check (the caller then runs the pivoted kernels over the batch).
int rbffd_weights_nsw(rbffd_context* ctx, const OpTables& T, const double* X, int64_t NS, const double* Y, int64_t M,
                      const int32_t* stencils, const int32_t* center, int32_t* colind_out, double* vals_out) {
    const int nb = T.n - T.q;
    if (T.nops > 8 || T.n > 64 || T.n < 8 || T.n + T.nops > 64 || nb < 1 || nb > WN_NBP || T.dim < 2 || NS != M) return RBFFD_ERR_UNSUPPORTED;
    {
        int need = 1;
        const int deg = (T.p - 1) / 2;
        for (int tq = 1; tq <= T.dim; ++tq) need = need * (deg + tq) / tq;
        if (T.q < need) return RBFFD_ERR_UNSUPPORTED;
    }
    WNArgs a;
    a.X = X; a.Y = Y; a.stencils = stencils; a.center = center; a.NS = NS; a.M = M;
    a.colind = colind_out; a.vals = vals_out; a.T = T;
    for (int o = 0; o < 8; ++o) { a.gzcol[o] = -1; a.gzval[o] = 0.0; }
    for (int ax = 0; ax < 3; ++ax) a.lapcol[ax] = -1;
    for (int c = 0; c < T.q; ++c) {
        for (int ax = 0; ax < T.dim; ++ax) {
            bool sq = true;
            for (int b = 0; b < T.dim; ++b) sq = sq && T.mono[c][b] == (b == ax ? 2 : 0);
            if (sq) a.lapcol[ax] = c;
        }
        for (int o = 0; o < T.nops; ++o) {
            if (T.kind[o] != RBFFD_OP_DERIV) continue;
            bool hit = true;
            double v = 1.0;
            for (int ax = 0; ax < T.dim; ++ax) {
                hit = hit && T.mono[c][ax] == T.alpha[o][ax];
                for (int u = 2; u <= T.alpha[o][ax]; ++u) v *= (double)u;
            }
            if (hit) { a.gzcol[o] = c; a.gzval[o] = v; }
        }
    }
    DevBuf<int> redo;
    const bool deferred = ctx->deferred_flags != nullptr;
    if (deferred) a.redo = ctx->deferred_flags + 8 * ctx->deferred_slot + 4;
    else {
        CUDA_TRY(ctx, redo.alloc(1, ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(redo.p, 0, sizeof(int), ctx->stream));
        a.redo = redo.p;
    }
    int rc = RBFFD_ERR_UNSUPPORTED;
    if (T.dim == 2 && T.q == 10) rc = launch_nsw<2, 10>(ctx, a);
    else if (T.dim == 2 && T.q == 15) rc = launch_nsw<2, 15>(ctx, a);
    else if (T.dim == 2 && T.q == 21) rc = launch_nsw<2, 21>(ctx, a);
    else if (T.dim == 3 && T.q == 10) rc = launch_nsw<3, 10>(ctx, a);
    else if (T.dim == 3 && T.q == 20) rc = launch_nsw<3, 20>(ctx, a);
    if (rc != RBFFD_OK || deferred) return rc;
    int h_redo = 0;
    CUDA_TRY(ctx, rbffd_fetch_flags(ctx, redo.p, 1, &h_redo));
    return h_redo ? RBFFD_ERR_UNSUPPORTED : RBFFD_OK;
}
