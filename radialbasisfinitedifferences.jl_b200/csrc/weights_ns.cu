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
//      (Y = Phi[:,N] - Phi[:,B] W,  S = Y[N,:] - W' Y[B,:]; the right-hand sides ride along as extra columns)
//      and eliminated WITHOUT pivoting;
//   3. w[N] = S^-1 Z'(b - Phi w_p),  w[B] = w_p - W w[N].
// Measured against extended precision on the BASELINE stencil families the error is <= 0.06 eps cond(A), the same as
// pivoted LU; a non-definite S or a rank-deficient P raises a flag and the batch is redone by the
// Gauss-Jordan kernel (weights_fast.cu).  Work: ~(n-q)^2 (n+q) flops instead of (n+q)^3, q pivot searches instead
// of n+q, and half the registers, i.e. more resident warps.
//
// The kernel is specialised at compile time on (dimension, number of monomials): the graded monomial table of
// build_op_tables (weights.cu) is spelled out as straight-line code, the pivot row of every reduction step travels by
// SHFL, the right-hand sides ride in the spare columns of the last null-space tile when they fit (nb + nops <= 24),
// and nothing is zero-filled that a later phase overwrites or never reads.
// Scope: collocated rows, n <= 32, (d, q) in {(2,3), (2,6), (2,10), (3,4), (3,10)}, n - q <= 24, <= 8 operators,
// polydeg >= (p-1)/2.   Replaces the same reference lines as weights.cu.
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "common.cuh"
#include "tables.cuh"
#include "phs.cuh"
#include "nullspace.cuh"

namespace {

struct NArgs {
    const double* X;
    const double* Y;
    const int32_t* stencils;   // [NX][n]
    const int32_t* center;     // [M] stencil of row i (Y != X: generate_operator.jl:47,103-108), or null: row i uses stencil i
    int64_t NS, M;             // NS = rows processed (== M)
    int32_t* colind;           // [M][n]
    double* vals;              // [nops][M][n]
    int* fail;                 // flags[0]: singular node + 1
    int* redo;                 // set to 1 when any stencil needs the pivoted fallback
    // polynomial right-hand side at eta == 0: DERIV operator o hits exactly one monomial (column gzcol, value alpha!)
    int32_t gzcol[8];
    double gzval[8];
    int32_t bs;                // row stride of the RBF right-hand-side tile (nops rounded up to even)
    int32_t smem_per_warp;
    // segmented mode (Y != X, several rows per centre: generate_operator.jl:89-95,158 -- the rows of one centre share inv(A)):
    // work item k = { centre, up to three rows of that centre (-1: none) }; the rows ride as extra right-hand-side columns
    const int4* items;
    const int* nitems;         // device: number of work items
    int32_t gr;                // rows per item (2 or 3)
    OpTables T;
};

__device__ __forceinline__ void dmma884n(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 1/x to ~1 ulp: MUFU.RCP64H seed (relative error ~2^-20) + one third-order step y (1 + e + e^2), e = 1 - x y
__device__ __forceinline__ double rcp3(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double u = fma(e, e, e);
    return fma(y, u, y);
}

// max over the warp of a non-negative double: ordering of non-negative doubles == ordering of their bit patterns
__device__ __forceinline__ double warp_max_nonneg(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned hmax = __reduce_max_sync(0xffffffffu, hi);
    const unsigned lmax = __reduce_max_sync(0xffffffffu, hi == hmax ? lo : 0u);
    return __hiloint2double((int)hmax, (int)lmax);
}

constexpr int NS_NB = 24;      // padded null-space dimension (3 tiles)
// Row strides == 4 (mod 16) doubles: the symmetric pair stores of the assembly (addresses 37 l + k and 37 l + 36 k over the
// lanes l), the DMMA fragment loads (36 g + t) and the Y_B fragment loads (36 t + g) are all free of bank conflicts.
constexpr int NS_LD = 36;      // row stride of the Phi tile
constexpr int NS_US = 36;      // row stride of the Y / [S|t] exchange tile (even: 16-byte rows)

// graded monomials of build_op_tables (weights.cu): mono[t] = mono[mpar[t]] * x[maxis[t]]
template <int D, int Q>
__device__ __forceinline__ void mono_row(const double* x, double* m) {
    m[0] = 1.0;
    if constexpr (D == 2) {
        m[1] = x[0]; m[2] = x[1];
        if constexpr (Q > 3) { m[3] = m[1] * x[0]; m[4] = m[1] * x[1]; m[5] = m[2] * x[1]; }
        if constexpr (Q > 6) { m[6] = m[3] * x[0]; m[7] = m[3] * x[1]; m[8] = m[4] * x[1]; m[9] = m[5] * x[1]; }
    } else {
        m[1] = x[0]; m[2] = x[1]; m[3] = x[2];
        if constexpr (Q > 4) { m[4] = m[1] * x[0]; m[5] = m[1] * x[1]; m[6] = m[1] * x[2]; m[7] = m[2] * x[1]; m[8] = m[2] * x[2]; m[9] = m[3] * x[2]; }
    }
}
// column of the monomial x_a^2 (the only one with a non-zero Laplacian at 0), -1 when the degree is below 2
template <int D, int Q>
__host__ __device__ constexpr int lap_col(int a) {
    return D == 2 ? (Q > 3 ? (a == 0 ? 3 : 5) : -1) : (Q > 4 ? (a == 0 ? 4 : (a == 1 ? 7 : 9)) : -1);
}

// NN, NO != 0: stencil size and operator count fixed at compile time (the BASELINE configs[1] shape n = 30, one operator):
// the column classification of the Y tile, the padding tests, the block-step guards and the back-substitution trip counts
// then fold to constants.
// PP != 0: PHS power fixed; COLLOC: every row is evaluated at its own stencil centre (Y == X, no centre indirection), so eta == 0.
// SNJ != 0: segmented mode, see NArgs::items.  One elimination per work item; every (row, operator) pair of the item is one
// right-hand-side column (<= 18 of them), the matrix has SNJ = 3, 4 or 5 tile columns (the smallest count that holds
// rc0 + rows x operators columns; [S | t] up to 24 x 40), only the rows of the basic nodes of Y are staged through shared memory
// (compact tile, stride 40), and the weights of the basic nodes of all columns come out of one DMMA product.
template <int D, int Q, int MINB, bool FOLD, int NN = 0, int NO = 0, int PP = 0, bool COLLOC = false, int SNJ = 0>
__global__ void __launch_bounds__(128, MINB) weights_ns_kernel(NArgs a) {
    constexpr bool SEG = SNJ != 0;
    constexpr int LD = NS_LD, US = SEG ? 40 : NS_US;
    constexpr int GRM = SEG ? 3 : 1;                  // rows per work item (upper bound)
    constexpr int NJ = SEG ? SNJ : (FOLD ? 3 : 4);    // tile columns of [S | t]
    constexpr int NJM = NJ > 4 ? NJ : 4;
    constexpr int WTR = NJ > 4 ? 40 : 32;             // rows of the W' block
    constexpr int YSS = SEG ? 28 : NS_NB;             // stride of the solution columns (SEG: == 12 (mod 16), they are DMMA operands)
    static_assert(!SEG || (!COLLOC && NN == 0 && NO == 0), "segmented mode: generic shape, general evaluation points");
    constexpr int KS = (Q + 3) / 4, QP = 4 * KS;      // k-steps of the DMMAs over the basic nodes
    constexpr int PCS = (Q + 2) & ~1;                 // published pivot row of the column reduction: Q entries + 1 / pivot, even
    extern __shared__ __align__(16) unsigned char nsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const OpTables& T = a.T;
    const int n = NN ? NN : T.n, nops = NO ? NO : T.nops, nb = n - Q, BS = NO ? ((NO + 1) & ~1) : a.bs;
    const int nrt = SEG ? a.gr * nops : nops;         // right-hand-side columns of a work item
    double* G = reinterpret_cast<double*>(nsm + (size_t)warp * a.smem_per_warp);   // Phi~ (stride LD), then Y, then [S|t]
    double* Yb = G;                                   // aliases G: written only after every read of Phi~ is done
    double* Wt = G + 32 * NS_US;                       // [32][QP]: W' rows of the non-basic nodes, then the w_p rows
    double* Bt = Wt + WTR * QP;                       // [32][BS]: RBF right-hand sides by position
    double* Ys = Bt;                                  // solution y, [op][24] (Bt is dead by then)
    constexpr int DP = D == 2 ? 2 : 4;                // doubles per stored point (16-byte aligned)
    double* Sc = Bt + 32 * BS;                        // permuted scaled coordinates, [pos][DP]
    int* perm = reinterpret_cast<int*>(Sc + 32 * DP);
    const double EPS = 2.220446049250313e-16;
    const unsigned FULL = 0xffffffffu;
    const int pw = PP ? PP : T.p;
    const double sgn = (((pw + 1) >> 1) & 1) ? -1.0 : 1.0;        // (-1)^((p+1)/2) S is positive definite
    const int hp = (pw - 1) >> 1;
    const int sgnbits = sgn < 0.0 ? (int)0x80000000 : 0;
    // right-hand-side columns: in the spare columns of the last null-space tile when they fit, else in a 4th tile column
    const int rc0 = (nb + 3) & ~3;
    constexpr bool fold = FOLD || SEG;                // host: rc0 + nops <= NS_NB (SEG: rc0 + gr nops <= 40)
    const int rcb = fold ? rc0 : NS_NB;               // first right-hand-side column; also the row of w_p in Wt
    const bool gl = !SEG && n + nops <= 32;           // g rows fit into spare lanes of the column reduction
    const int go = gl ? lane - n : lane;              // right-hand-side column whose g row this lane owns
    const bool gown = go >= 0 && go < nrt;

    // node id of this lane in the stencil of row i; the ids of the NEXT row are fetched one iteration ahead and its
    // coordinates are pulled into L1 under the elimination, so that phase 0 does not sit on two dependent DRAM round trips
    auto stencil_id = [&](int64_t item) -> int {
        const int64_t ctr = SEG ? (int64_t)a.items[item].x : (a.center ? (int64_t)a.center[item] : item);
        const int32_t* st = a.stencils + ctr * n;
        return st[lane < n ? lane : 0];
    };
    const int64_t NSI = SEG ? (int64_t)*a.nitems : a.NS;
    const int64_t istride = (int64_t)gridDim.x * 4;
    int64_t i = blockIdx.x * 4ll + warp;
    int id_next = i < NSI ? stencil_id(i) : 0;
    for (; i < NSI; i += istride) {
        // ---- 0. scalestencil.jl:10-20: lane l owns stencil node l ----
        const int id = id_next;
        if (i + istride < NSI) id_next = stencil_id(i + istride);
        // rows of this work item (SEG: up to three rows of one centre; a missing row repeats row 0 and is not stored)
        int64_t rowid[GRM];
        int cnt = 1;
        if constexpr (SEG) {
            const int4 it = a.items[i];
            rowid[0] = it.y;
            if constexpr (GRM > 1) { rowid[1] = it.z >= 0 ? it.z : it.y; cnt += it.z >= 0; }
            if constexpr (GRM > 2) { rowid[2] = it.w >= 0 ? it.w : it.y; cnt += it.w >= 0; }
        } else rowid[0] = i;
        double sx[D], s[D], eta[GRM][D];
        bool eta_zero = !SEG;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const double xv = a.X[(int64_t)id * D + c];
            const double xc = __shfl_sync(FULL, xv, 0);
            sx[c] = xv - xc;
            s[c] = 1.0 / warp_max_nonneg(fabs(sx[c]));
            sx[c] = sx[c] * s[c];
            if constexpr (COLLOC) eta[0][c] = 0.0;
            else {
#pragma unroll
                for (int r = 0; r < GRM; ++r) {
                    eta[r][c] = (a.Y[rowid[r] * D + c] - xc) * s[c];
                    eta_zero = eta_zero && (eta[r][c] == 0.0);
                }
            }
        }
        // ---- 1. column reduction of [P; g']: lane l < n holds row l of P, lane n+o (or registers grow) the row g_o' ----
        double prow[Q], grow[Q];
        mono_row<D, Q>(sx, prow);
        if (lane >= n) {
#pragma unroll
            for (int c = 0; c < Q; ++c) prow[c] = 0.0;
        }
        {
            double gv[Q];
            if (eta_zero) {
                const int og = gown ? go : 0;
                const bool lap = T.kind[og] == RBFFD_OP_LAPLACE;
                const int col = a.gzcol[og];
                const double val = a.gzval[og];
#pragma unroll
                for (int c = 0; c < Q; ++c) {
                    double v = (c == col) ? val : 0.0;
#pragma unroll
                    for (int ax = 0; ax < D; ++ax)
                        if (c == lap_col<D, Q>(ax)) v = lap ? 2.0 * s[ax] * s[ax] : v;
                    gv[c] = v;
                }
            } else {
                // general evaluation point: lane c evaluates monomial c for every operator, staged through shared
                // memory so that the row of operator o lands in one lane
#pragma unroll
                for (int r = 0; r < GRM; ++r) {
                    if (r * nops < nrt) {
                        for (int o = 0; o < nops; ++o) {
                            const double v = lane < Q ? rhs_poly_entry<D>(T, o, lane, eta[r], s) : 0.0;
                            if (lane < Q) Wt[(r * nops + o) * QP + lane] = v;
                        }
                    }
                }
                __syncwarp();
#pragma unroll
                for (int c = 0; c < Q; ++c) gv[c] = gown ? Wt[go * QP + c] : 0.0;
                __syncwarp();
            }
#pragma unroll
            for (int c = 0; c < Q; ++c) {
                if (gl) { if (gown) prow[c] = gv[c]; grow[c] = 0.0; }
                else grow[c] = gown ? gv[c] : 0.0;
            }
        }
        unsigned kmin = 0xffffffffu;
        int bad = 0;                                            // sign bit set: a pivot of S had the wrong sign
        bool basic = false;
        int mybasic = 0;
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            const unsigned hi = (unsigned)__double2hiint(prow[j]) & 0x7fffffe0u;
            const unsigned key = (lane < n && !basic) ? (hi | (unsigned)lane) : 0u;
            const unsigned kmax = __reduce_max_sync(FULL, key);
            const double rown = rcp3(prow[j]);                  // every candidate inverts its own entry under the search
            kmin = min(kmin, kmax);                             // < 32: P is rank deficient on this stencil
            const int pl = kmax & 31;
            // the winner publishes its row and the reciprocal of its pivot through shared memory (the Phi~ tile is idle in this
            // phase; two buffers alternate by step parity, so one __syncwarp per step suffices): Q/2 + 1 one-lane stores and as
            // many broadcast loads instead of 2 Q + 2 SHFLs per step
            double* cw = G + (j & 1) * PCS;
            if (lane == pl) {
                basic = true; mybasic = j;
                double2* dst = reinterpret_cast<double2*>(cw);
#pragma unroll
                for (int c = 0; c < PCS; c += 2) dst[c >> 1] = make_double2(c < Q ? prow[c] : rown, c + 1 < Q ? prow[c + 1] : rown);
            }
            __syncwarp();
            double pr[Q];
#pragma unroll
            for (int c = 0; c < Q; c += 2) {
                const double2 v = reinterpret_cast<const double2*>(cw)[c >> 1];
                pr[c] = v.x;
                if (c + 1 < Q) pr[c + 1] = v.y;
            }
            const double rinv = cw[Q];
            const double tl = prow[j] * rinv;
#pragma unroll
            for (int c = 0; c < Q; ++c)
                if (c != j) prow[c] = fma(-tl, pr[c], prow[c]);
            prow[j] = tl;
            if (!gl) {                                          // g rows kept in the second register set
                const double tg = grow[j] * rinv;
#pragma unroll
                for (int c = 0; c < Q; ++c)
                    if (c != j) grow[c] = fma(-tg, pr[c], grow[c]);
                grow[j] = tg;
            }
        }
        // positions: non-basic nodes first (0..nb-1, in stencil order), then the basic ones in pivot order
        const unsigned nbmask = __ballot_sync(FULL, lane < n && !basic);
        const int pos = lane < n ? (basic ? nb + mybasic : __popc(nbmask & ((1u << lane) - 1u))) : 31;
        // ---- 2. stage W', w_p, permuted coordinates and the RBF right-hand sides ----
        {
            // rows of Wt: non-basic node -> its position; g row of operator o -> rcb + o
            const bool wnb = lane < n && !basic;
            const int wrow = wnb ? pos : (gown ? rcb + go : -1);
            if (wrow >= 0) {
                double2* dst = reinterpret_cast<double2*>(Wt + wrow * QP);
                const bool useg = !gl && !wnb;
#pragma unroll
                for (int c = 0; c < QP; c += 2) {
                    const double v0 = c < Q ? (useg ? grow[c] : prow[c]) : 0.0;
                    const double v1 = c + 1 < Q ? (useg ? grow[c + 1] : prow[c + 1]) : 0.0;
                    dst[c >> 1] = make_double2(v0, v1);
                }
            }
            if (!gl && gown && lane < n && !basic) {            // this lane owns a W' row AND a g row
                double2* dst = reinterpret_cast<double2*>(Wt + (rcb + go) * QP);
#pragma unroll
                for (int c = 0; c < QP; c += 2)
                    dst[c >> 1] = make_double2(c < Q ? grow[c] : 0.0, c + 1 < Q ? grow[c + 1] : 0.0);
            }
        }
        if (lane < n) {
            perm[pos] = lane;
            G[pos * LD + pos] = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) Sc[pos * DP + c] = sx[c];
            // RBF part of the right-hand sides at this node (generate_operator.jl:123-154)
#pragma unroll
            for (int rw = 0; rw < GRM; ++rw) {
                if (rw * nops < nrt) {
                    double del[D];
                    double r2 = 0.0;
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        const double dd = eta[rw][c] - sx[c];
                        del[c] = dd == 0.0 ? EPS : dd;
                        r2 = fma(del[c], del[c], r2);
                    }
                    const double y = phs_rsqrt(r2);
                    double rp4 = y;                             // r^(p-4)
                    for (int e = 1; e < hp; ++e) rp4 *= r2;
                    const double rp2 = rp4 * r2, rp = rp2 * r2, r = r2 * y;
                    for (int o = 0; o < nops; ++o) Bt[pos * BS + rw * nops + o] = rhs_rbf_entry_fast<D>(T, o, del, s, r, r2, rp, rp2, rp4);
                }
            }
        }
        __syncwarp();
        // ---- 3. Phi~ in permuted order, by symmetric pairs (phs.cuh): lane l owns position l ----
        {
            double me[D];
            const int l = lane < n ? lane : 0;
#pragma unroll
            for (int c = 0; c < D; ++c) me[c] = Sc[l * DP + c];
            phs_assemble<D, DP, LD>(Sc, G, me, l, n, lane < n, hp);
        }
        __syncwarp();
        // ---- 4. Y = Phi~[:, N] - Phi~[:, B] W   (4 x NJ tiles; the right-hand-side columns start from b and use w_p) ----
        double c[4][NJM][2];
        {
#pragma unroll
            for (int J = 0; J < NJM; ++J) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 8 * J + 2 * t + e;
                    // source of column `col`: Phi~ (null-space column), b (right-hand side) or zero (G[0] is a zero)
                    const bool isn = col < nb, isr = col >= rcb && col < rcb + nrt;
                    const double* src = isn ? G + g * LD + col : (isr ? Bt + g * BS + (col - rcb) : G);
                    const int str = isn ? 8 * LD : (isr ? 8 * BS : 0);
#pragma unroll
                    for (int I = 0; I < 4; ++I) c[I][J][e] = J < NJ ? src[I * str] : 0.0;
                }
            }
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                double af[4], bf[NJM];
#pragma unroll
                for (int I = 0; I < 4; ++I) af[I] = (4 * k + t < Q) ? -G[(8 * I + g) * LD + nb + 4 * k + t] : 0.0;
#pragma unroll
                for (int J = 0; J < NJM; ++J) bf[J] = J < NJ ? Wt[(8 * J + g) * QP + 4 * k + t] : 0.0;
#pragma unroll
                for (int J = 0; J < NJM; ++J)
                    if (J < NJ) {
#pragma unroll
                        for (int I = 0; I < 4; ++I) dmma884n(c[I][J][0], c[I][J][1], af[I], bf[J]);
                    }
            }
        }
        // ---- 5. [S | t] = Y[N, :] - W' Y[B, :] ----
        __syncwarp();                                     // Phi~ is dead: the Y tile reuses its storage
        // only the rows of the basic nodes (Y_B, positions nb .. n-1) are read back; Y_N stays in the accumulators
        // (SEG: compact tile, row = position - nb, stride 40 -- 40 columns do not fit the stride of the Phi~ rows)
        constexpr int YR0 = SEG ? 1 : 0;                  // 1: rows are stored relative to nb
#pragma unroll
        for (int J = 0; J < NJM; ++J)
            if (J < NJ) {
#pragma unroll
                for (int I = 0; I < 4; ++I)
                    if (8 * I + 8 > nb && (!SEG || (8 * I + g >= nb && 8 * I + g < n)))
                        *reinterpret_cast<double2*>(Yb + (8 * I + g - YR0 * nb) * US + 8 * J + 2 * t) = make_double2(c[I][J][0], c[I][J][1]);
            }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            double af[3], bf[NJM];
#pragma unroll
            for (int I = 0; I < 3; ++I) af[I] = -Wt[(8 * I + g) * QP + 4 * k + t];
#pragma unroll
            for (int J = 0; J < NJM; ++J) bf[J] = (J < NJ && 4 * k + t < Q) ? Yb[(nb - YR0 * nb + 4 * k + t) * US + 8 * J + g] : 0.0;
#pragma unroll
            for (int J = 0; J < NJM; ++J)
                if (J < NJ) {
#pragma unroll
                    for (int I = 0; I < 3; ++I) dmma884n(c[I][J][0], c[I][J][1], af[I], bf[J]);
                }
        }
        __syncwarp();
        // identity padding outside the nb x nb block; right-hand-side columns of padded rows are zero.  Only tiles that
        // reach past row / column nb are touched (warp-uniform tests).
#pragma unroll
        for (int I = 0; I < 3; ++I) {
            const int row = 8 * I + g;
#pragma unroll
            for (int J = 0; J < 3; ++J) {
                if (8 * I + 8 > nb || 8 * J + 8 > nb) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int col = 8 * J + 2 * t + e;
                        if (col >= rcb) { if (row >= nb) c[I][J][e] = 0.0; }
                        else if (row >= nb || col >= nb) c[I][J][e] = row == col ? sgn : 0.0;
                    }
                }
            }
#pragma unroll
            for (int J = 3; J < NJM; ++J)
                if (J < NJ && row >= nb) { c[I][J][0] = 0.0; c[I][J][1] = 0.0; }
        }
        // ---- 6. blocked Gauss-Jordan WITHOUT pivoting on the definite S: the block step of weights_fast.cu with
        //         static pivot rows (row 4kb+s), so no pivot search, no row selects and a static pivot-row dump ----
        if (i + istride < NSI) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.X + (int64_t)id_next * D));
            if (!COLLOC && !SEG && lane < D) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.Y + (i + istride) * D + lane));
        }
        // 4 x 4 block pivots (nullspace.cuh): the pivot block is inverted in every lane and applied by DMMAs, operands change
        // fragment layout by SHFL -- no panel / pivot-row dumps through shared memory, one __syncwarp per block step
        __syncwarp();                                     // every lane is done with the Y tile: its first 256 B carry the pivot blocks
        bad |= nsp::block_gj_warp<3, NJ, 4, NJM, true>(c, nb, G, sgnbits);
        // the matrix is now [I | y]  (right-hand-side column rcb + o lives in tile (rcb + o) / 8)
#pragma unroll
        for (int I = 0; I < 3; ++I) {
            const int row = 8 * I + g;
#pragma unroll
            for (int J = 0; J < NJM; ++J)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int o = 8 * J + 2 * t + e - rcb;
                    if (row < nb && J < NJ && o >= 0 && o < nrt) Ys[o * YSS + row] = c[I][J][e];
                }
        }
        __syncwarp();
        // ---- 7. w[N] = y, w[B] = w_p - W y; rescale and scatter into the CSR row (generate_operator.jl:161-182) ----
        const int dstj = perm[lane < n ? lane : 0];
        bool ok = kmin >= 32u && bad >= 0;
        const int nout = SEG ? cnt * nops : nops;         // columns of rows that exist
        if constexpr (SEG) {
            // all (row, operator) columns at once: the non-basic weights are y itself, the Q x nrt block of the basic nodes
            // w_p - W' y is a DMMA product (rows = basic nodes, columns = right-hand sides, k = non-basic positions)
            const double INF = __longlong_as_double(0x7ff0000000000000ll);
            double* pf = Sc;                                // chain-rule factor of every column (the coordinates are dead)
            if (lane < nrt) pf[lane] = op_post_factor<D>(T, lane % nops, s);
            __syncwarp();
            bool fin = true;
            auto out_row = [&](int j, int& o) -> int64_t {
                const int rw = j >= 2 * nops ? 2 : (j >= nops ? 1 : 0);
                o = j - rw * nops;
                return rw == 0 ? rowid[0] : (rw == 1 ? rowid[GRM > 1 ? 1 : 0] : rowid[GRM > 2 ? 2 : 0]);
            };
            if (lane < nb) {
                for (int j = 0; j < nout; ++j) {
                    int o;
                    const int64_t orow = out_row(j, o);
                    const double wv = Ys[j * YSS + lane] * pf[j];
                    fin = fin && (fabs(wv) < INF);          // a zero pivot shows up as a non-finite weight
                    a.vals[((int64_t)o * a.M + orow) * n + dstj] = wv;
                }
            }
            constexpr int QT = (Q + 7) / 8;
            double acc[QT][3][2];
#pragma unroll
            for (int I = 0; I < QT; ++I)
#pragma unroll
                for (int J = 0; J < 3; ++J)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int cc = 8 * I + g, j = 8 * J + 2 * t + e;
                        acc[I][J][e] = (cc < Q && j < nrt) ? Wt[(rcb + j) * QP + cc] : 0.0;
                    }
            for (int k0 = 0; k0 < nb; k0 += 4) {
                const int aa = k0 + t;
                const bool kin = aa < nb;
                double af[QT], bf[3];
#pragma unroll
                for (int I = 0; I < QT; ++I) af[I] = (8 * I + g < Q && kin) ? -Wt[(kin ? aa : 0) * QP + 8 * I + g] : 0.0;
#pragma unroll
                for (int J = 0; J < 3; ++J) bf[J] = (8 * J + g < nrt && kin) ? Ys[(8 * J + g) * YSS + aa] : 0.0;
#pragma unroll
                for (int J = 0; J < 3; ++J)
                    if (8 * J < nrt) {
#pragma unroll
                        for (int I = 0; I < QT; ++I) dmma884n(acc[I][J][0], acc[I][J][1], af[I], bf[J]);
                    }
            }
#pragma unroll
            for (int I = 0; I < QT; ++I) {
                const int cc = 8 * I + g;
                const int dstb = perm[cc < Q ? nb + cc : 0];
#pragma unroll
                for (int J = 0; J < 3; ++J)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = 8 * J + 2 * t + e;
                        if (cc < Q && j < nout) {
                            int o;
                            const int64_t orow = out_row(j, o);
                            const double wv = acc[I][J][e] * pf[j];
                            fin = fin && (fabs(wv) < INF);
                            a.vals[((int64_t)o * a.M + orow) * n + dstb] = wv;
                        }
                    }
            }
            ok = ok && __all_sync(FULL, fin);
        } else
        for (int j = 0; j < nout; ++j) {
            int rw = 0, o = j;
            if constexpr (SEG) { rw = j >= 2 * nops ? 2 : (j >= nops ? 1 : 0); o = j - rw * nops; }
            const double f = op_post_factor<D>(T, o, s);
            const int64_t orow = SEG ? (rw == 0 ? rowid[0] : (rw == 1 ? rowid[GRM > 1 ? 1 : 0] : rowid[GRM > 2 ? 2 : 0])) : i;
            double* vrow = a.vals + ((int64_t)o * a.M + orow) * n;
            double wv = 0.0;
            if (lane < nb) wv = Ys[j * NS_NB + lane];
            else if (lane < n) {
                const int cc = lane - nb;
                double acc0 = Wt[(rcb + j) * QP + cc], acc1 = 0.0;
                const double* wcol = Wt + cc;
                const double* yv = Ys + j * NS_NB;
                int aa = 0;
                for (; aa + 1 < nb; aa += 2) {
                    acc0 = fma(-wcol[aa * QP], yv[aa], acc0);
                    acc1 = fma(-wcol[(aa + 1) * QP], yv[aa + 1], acc1);
                }
                if (aa < nb) acc0 = fma(-wcol[aa * QP], yv[aa], acc0);
                wv = acc0 + acc1;
            }
            wv *= f;
            // a zero pivot shows up as a non-finite weight
            ok = ok && __all_sync(FULL, fabs(wv) < __longlong_as_double(0x7ff0000000000000ll));
            if (lane < n) vrow[dstj] = ok ? wv : nan("");
        }
        if (lane < n) {
            if constexpr (SEG) {
#pragma unroll
                for (int r = 0; r < GRM; ++r)
                    if (r < cnt) a.colind[rowid[r] * n + lane] = id;
            } else a.colind[i * n + lane] = id;
        }
        if (!ok && lane == 0) *a.redo = 1;
        __syncwarp();
    }
}

// ---- segmented mode: rows grouped by centre, `gr` rows per work item ----
__global__ void ns_iota_kernel(int* p, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int)i;
}
// start[c] = first position of centre c in the sorted key array (start[nseg] = n)
__global__ void ns_seg_start_kernel(const int* __restrict__ key_sorted, int64_t n, int nseg, int* __restrict__ start) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i > n) return;
    const int prev = i == 0 ? -1 : key_sorted[i - 1];
    const int cur = i == n ? nseg : key_sorted[i];
    for (int c = prev + 1; c <= cur; ++c) start[c] = (int)i;
}
__global__ void ns_item_count_kernel(const int* __restrict__ start, int nseg, int gr, int* __restrict__ cnt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= nseg) cnt[c] = c < nseg ? (start[c + 1] - start[c] + gr - 1) / gr : 0;
}
__global__ void ns_item_fill_kernel(const int* __restrict__ start, const int* __restrict__ rows, const int* __restrict__ off, int nseg, int gr,
                                    int4* __restrict__ items) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nseg) return;
    const int s0 = start[c], s1 = start[c + 1];
    int k = off[c];
    for (int r = s0; r < s1; r += gr, ++k) {
        int4 it;
        it.x = c;
        it.y = rows[r];
        it.z = (gr > 1 && r + 1 < s1) ? rows[r + 1] : -1;
        it.w = (gr > 2 && r + 2 < s1) ? rows[r + 2] : -1;
        items[k] = it;
    }
}

// rows of one centre share the elimination (generate_operator.jl:89-95,158): sort the rows by centre, cut every centre's rows
// into work items of `gr` rows, run the 5-tile-column instance.  NX = number of stencils (centres).
template <int D, int Q>
int launch_ns_segmented(rbffd_context* ctx, NArgs& a, int64_t NX, int gr) {
    constexpr int QP = 4 * ((Q + 3) / 4);
    cudaStream_t st = ctx->stream;
    const int64_t M = a.M;
    DevBuf<int> keys_sorted, ident, rows, start, cnt, off;
    DevBuf<unsigned char> tmp;
    CUDA_TRY(ctx, keys_sorted.alloc(M, st));
    CUDA_TRY(ctx, ident.alloc(M, st));
    CUDA_TRY(ctx, rows.alloc(M, st));
    CUDA_TRY(ctx, start.alloc(NX + 1, st));
    CUDA_TRY(ctx, cnt.alloc(NX + 1, st));
    CUDA_TRY(ctx, off.alloc(NX + 1, st));
    ns_iota_kernel<<<ceil_div_i64(M, 256), 256, 0, st>>>(ident.p, M);
    KLAUNCH(ctx);
    int bits = 1;
    while ((1ll << bits) < NX) ++bits;
    size_t sort_bytes = 0, scan_bytes = 0;
    CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, a.center, keys_sorted.p, ident.p, rows.p, (int)M, 0, bits, st));
    CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, cnt.p, off.p, (int)(NX + 1), st));
    CUDA_TRY(ctx, tmp.alloc(std::max(sort_bytes, scan_bytes), st));
    CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, sort_bytes, a.center, keys_sorted.p, ident.p, rows.p, (int)M, 0, bits, st));
    ns_seg_start_kernel<<<ceil_div_i64(M + 1, 256), 256, 0, st>>>(keys_sorted.p, M, (int)NX, start.p);
    ns_item_count_kernel<<<ceil_div_i64(NX + 1, 256), 256, 0, st>>>(start.p, (int)NX, gr, cnt.p);
    KLAUNCH(ctx); KLAUNCH(ctx);
    CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, scan_bytes, cnt.p, off.p, (int)(NX + 1), st));
    const int64_t max_items = std::min<int64_t>(M, M / gr + NX + 1);
    DevBuf<int4> items;
    CUDA_TRY(ctx, items.alloc(max_items, st));
    ns_item_fill_kernel<<<ceil_div_i64(NX, 256), 256, 0, st>>>(start.p, rows.p, off.p, (int)NX, gr, items.p);
    KLAUNCH(ctx);
    a.items = items.p;
    a.nitems = off.p + NX;
    a.gr = gr;
    a.bs = (gr * a.T.nops + 1) & ~1;
    const int rc0 = (a.T.n - Q + 3) & ~3;
    const int snj = (rc0 + gr * a.T.nops + 7) / 8;       // tile columns: 3, 4 or 5
    a.smem_per_warp = ((32 * NS_US + (snj > 4 ? 40 : 32) * QP + 32 * a.bs + 32 * (D == 2 ? 2 : 4)) * 8 + 32 * 4 + 15) & ~15;
    const size_t smem = (size_t)a.smem_per_warp * 4;
    if ((int64_t)smem > ctx->max_smem_optin || snj > 5) return RBFFD_ERR_UNSUPPORTED;
    auto kern = snj <= 3 ? weights_ns_kernel<D, Q, 4, false, 0, 0, 0, false, 3>
                         : (snj == 4 ? weights_ns_kernel<D, Q, 4, false, 0, 0, 0, false, 4> : weights_ns_kernel<D, Q, 3, false, 0, 0, 0, false, 5>);
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int resident = std::max<int>(1, std::min<int>(snj > 4 ? 3 : 4, (int)((228 * 1024) / (smem + 1024))));
    // the item count lives on the device: the grid is sized for the upper bound, surplus CTAs find nothing to do
    const int grid = (int)std::min<int64_t>((max_items + 3) / 4, (int64_t)ctx->sm_count * resident * 128);
    kern<<<grid, 128, smem, st>>>(a);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

template <int D, int Q>
int launch_ns(rbffd_context* ctx, NArgs& a, int64_t NX) {
    constexpr int QP = 4 * ((Q + 3) / 4);
    // Y != X: the rows of a centre share one elimination whenever at least two of them fit the 40 columns of the segmented instance
    if (a.center != nullptr && NX > 0) {
        static const bool seg_on = [] { const char* e = getenv("RBFFD_NS_SEGMENTED"); return !e || atoi(e) != 0; }();
        const int rc0 = (a.T.n - Q + 3) & ~3;
        const int gr = std::min(3, (40 - rc0) / a.T.nops);
        if constexpr (Q >= 6) {                             // (polynomial degree 1 keeps the row-by-row path)
            if (seg_on && gr >= 2) return launch_ns_segmented<D, Q>(ctx, a, NX, gr);
        }
    }
    a.bs = (a.T.nops + 1) & ~1;
    a.smem_per_warp = ((32 * NS_US + 32 * QP + 32 * a.bs + 32 * (D == 2 ? 2 : 4)) * 8 + 32 * 4 + 15) & ~15;
    const size_t smem_used = (size_t)a.smem_per_warp * 4;
    // development knob: RBFFD_NS_PAD_SMEM = extra dynamic shared memory per CTA (occupancy sweeps)
    static const int pad_smem = [] { const char* e = getenv("RBFFD_NS_PAD_SMEM"); return e ? atoi(e) : 0; }();
    const size_t smem = smem_used + (size_t)std::max(0, pad_smem);
    if ((int64_t)smem > ctx->max_smem_optin) return RBFFD_ERR_UNSUPPORTED;
    // 4 CTAs (16 warps) per SM when the shared-memory tile allows it, else 3
    const bool four = (smem + 1024) * 4 <= 228 * 1024;
    const int nb = a.T.n - Q;
    const bool fold = ((nb + 3) & ~3) + a.T.nops <= NS_NB;
    auto kern = four ? (fold ? weights_ns_kernel<D, Q, 4, true> : weights_ns_kernel<D, Q, 4, false>)
                     : (fold ? weights_ns_kernel<D, Q, 3, true> : weights_ns_kernel<D, Q, 3, false>);
    if constexpr (D == 2 && Q == 10) {
        static const bool no_spec = [] { const char* e = getenv("RBFFD_NS_SPECIALIZE"); return e && atoi(e) == 0; }();
        if (four && fold && a.T.n == 30 && a.T.nops == 1 && !no_spec) {
            const bool colloc = a.center == nullptr && (a.Y == a.X || ctx->collocated_rows);      // row i is evaluated at the centre of stencil i
            kern = (a.T.p == 5 && colloc) ? weights_ns_kernel<2, 10, 4, true, 30, 1, 5, true> : weights_ns_kernel<2, 10, 4, true, 30, 1>;
        }
    }
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks_needed = (a.NS + 3) / 4;
    // CTAs per resident slot: many short CTAs balance better than a few long grid-stride loops and keep the concurrently
    // processed stencils in a compact window of the node array (sweep 1..100000: 8 -> 5.38 ms, 128 -> 5.05 ms, more is flat)
    static const int waves = [] { const char* e = getenv("RBFFD_NS_WAVES"); const int w = e ? atoi(e) : 0; return w > 0 ? w : 128; }();
    const int resident = std::max<int>(1, std::min<int>(four ? 4 : 3, (int)((228 * 1024) / (smem + 1024))));
    const int grid = (int)std::min<int64_t>(blocks_needed, (int64_t)ctx->sm_count * resident * waves);
    kern<<<grid, 128, smem, ctx->stream>>>(a);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

// the straight-line monomial code of mono_row must agree with the table the other kernels (and the g rows) use
template <int D, int Q>
bool table_matches(const OpTables& T) {
    static const int8_t e2[10][2] = {{0, 0}, {1, 0}, {0, 1}, {2, 0}, {1, 1}, {0, 2}, {3, 0}, {2, 1}, {1, 2}, {0, 3}};
    static const int8_t e3[10][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {2, 0, 0}, {1, 1, 0}, {1, 0, 1}, {0, 2, 0}, {0, 1, 1}, {0, 0, 2}};
    for (int c = 0; c < Q; ++c)
        for (int ax = 0; ax < D; ++ax)
            if (T.mono[c][ax] != (D == 2 ? e2[c][ax] : e3[c][ax])) return false;
    return true;
}

}  // namespace

// Null-space fast path.  Returns RBFFD_ERR_UNSUPPORTED when the configuration is outside its scope, or when any stencil
// failed its definiteness / rank check (the caller then runs the pivoted Gauss-Jordan kernel over the batch).
// NX: number of stencils when `center` is given (0: unknown, rows are then solved one by one).
int rbffd_weights_ns(rbffd_context* ctx, const OpTables& T, const double* X, int64_t NS, const double* Y, int64_t M,
                     const int32_t* stencils, const int32_t* center, int32_t* colind_out, double* vals_out, int* fail_flag, int64_t NX) {
    const int nb = T.n - T.q;
    if (T.nops > 8 || T.n > 32 || nb < 1 || nb > NS_NB || T.dim < 2 || NS != M) return RBFFD_ERR_UNSUPPORTED;
    // conditional definiteness needs polynomial degree >= (p-1)/2: q >= C((p-1)/2 + d, d)
    {
        int need = 1;
        const int deg = (T.p - 1) / 2;
        for (int tq = 1; tq <= T.dim; ++tq) need = need * (deg + tq) / tq;
        if (T.q < need) return RBFFD_ERR_UNSUPPORTED;
    }
    NArgs a;
    a.X = X; a.Y = Y; a.stencils = stencils; a.center = center; a.NS = NS; a.M = M;
    a.colind = colind_out; a.vals = vals_out; a.fail = fail_flag; a.T = T;
    a.items = nullptr; a.nitems = nullptr; a.gr = 1;
    // d^alpha x^e (0) = alpha! [e == alpha]
    for (int o = 0; o < 8; ++o) { a.gzcol[o] = -1; a.gzval[o] = 0.0; }
    for (int o = 0; o < T.nops; ++o) {
        if (T.kind[o] != RBFFD_OP_DERIV) continue;
        for (int c = 0; c < T.q; ++c) {
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
    if (T.dim == 2 && T.q == 3 && table_matches<2, 3>(T)) rc = launch_ns<2, 3>(ctx, a, NX);
    else if (T.dim == 2 && T.q == 6 && table_matches<2, 6>(T)) rc = launch_ns<2, 6>(ctx, a, NX);
    else if (T.dim == 2 && T.q == 10 && table_matches<2, 10>(T)) rc = launch_ns<2, 10>(ctx, a, NX);
    else if (T.dim == 3 && T.q == 4 && table_matches<3, 4>(T)) rc = launch_ns<3, 4>(ctx, a, NX);
    else if (T.dim == 3 && T.q == 10 && table_matches<3, 10>(T)) rc = launch_ns<3, 10>(ctx, a, NX);
    if (rc != RBFFD_OK || deferred) return rc;
    int h_redo = 0;
    CUDA_TRY(ctx, rbffd_fetch_flags(ctx, redo.p, 1, &h_redo));
    return h_redo ? RBFFD_ERR_UNSUPPORTED : RBFFD_OK;
}
