// weights_ns2.cu -- two-stage null-space weight path (K4) for the larger stencils of BASELINE configs 1, 3, 4, 5
// (32 < n <= 64).  Same mathematics as weights_ns.cu / weights_nsw.cu (see the header of weights_ns.cu): column
// reduction of P with row pivoting, S = Z' Phi Z by FP64 DMMAs, unpivoted blocked Gauss-Jordan on the definite S,
// w[N] = y, w[B] = w_p - W y.  What is new is HOW the two serial chains of that algorithm are scheduled:
//
//   stage A  ns2_pred_kernel   ONE WARP per stencil, two rows of [P; g'] per lane: the q pivoted reduction steps need no
//            barrier (warp-wide REDUX, winner row broadcast through 176 B of shared memory) and the kernel keeps no tile
//            in shared memory, so 16 stencils per SM are reduced concurrently instead of 4 (in weights_nsw.cu this chain
//            ran inside the 4-warp CTA at 850 clocks per step while two warps waited).  It writes, per stencil, a record
//            { s, x_c, node ids by POSITION (non-basic | basic), position -> stencil slot, W' rows, w_p rows }:
//            (n - q + nops) q doubles + 384 B = 7.4 KB at config 4, into a scratch buffer that is reused chunk by chunk
//            (nothing of size m^2 ever reaches HBM).
//   stage B  ns2_solve_kernel  one CTA of 4 warps per stencil: all four warps assemble Phi directly in PERMUTED order
//            (circulant pairing, two half-sets of rounds), so the DMMA fragment loads of Y = Phi~[:,N] - Phi~[:,B] W
//            are contiguous and conflict-free (the permuted gather cost 11 % of the samples before); S, the elimination
//            and the back substitution work on tiles sized for the stencil at hand (NT = ceil(nb/8) row tiles, NJ
//            tile columns) instead of the fixed 48 x 56 padding.
//            The instances compiled for the BASELINE shapes form S from the rows of the basic nodes of Y only:
//            S = V + V',  V = Phi~[N,N]/2 - W' T,  T = Phi~[B,N] - Phi~[B,B] W / 2  (Phi is symmetric), stored in place in the
//            Phi~ tile -- about half the DMMAs of the Y phase (phase B' / C' of ns2_solve_kernel).
//   stage C  ns2_elim1_kernel  ONE WARP per stencil: [S | t] register-resident from the load to the last pivot, unpivoted
//            Gauss-Jordan by 4 x 4 block pivots (nullspace.cuh), back substitution and the CSR row.
//   The three stages work through the rows in chunks sized by a scratch budget (min(12 GiB, 1/8 of the device memory)).
//
// Compiled once per (dimension, monomial count): -DNS2_D=<d> -DNS2_Q=<q> emits the kernels and
// rbffd_ns2_launch_<d>_<q>; without those macros the file emits the dispatcher rbffd_weights_ns2.
// Replaces the same reference lines as weights.cu (scalestencil.jl:10-20, interpolationmatrix.jl:5-8,
// generate_operator.jl:55-65,89-182, hyperviscosity_operator.jl:97-161).
#include <algorithm>
#include <cstdlib>

#include "nullspace.cuh"
#include "phs.cuh"

struct Ns2Args {
    const double* X;
    const double* Y;
    const int32_t* stencils;   // [NX][n]
    const int32_t* center;     // [M] stencil of row i (Y != X), or null: row i uses stencil i
    int64_t row0, cnt, M;      // rows [row0, row0 + cnt) of the M rows
    unsigned char* rec;        // stage-A records of this chunk, record k at rec + k * rec_stride
    int64_t rec_stride;
    int32_t* colind;           // [M][n]
    double* vals;              // [nops][M][n]
    int* redo;                 // set to 1 when any stencil needs the pivoted fallback
    int32_t gzcol[8];          // polynomial right-hand side at eta == 0: DERIV operator o hits exactly one monomial
    double gzval[8];
    int32_t lapcol[3];         // column of x_a^2 (-1: degree < 2)
    int32_t bs;                // row stride of the RBF right-hand-side tile
    int32_t ws, rcb;           // row stride of the transposed W' block, position of the first w_p row
    double* stile;             // split path: [S | t] of every stencil of the chunk as DMMA accumulator tiles,
    int64_t stile_stride;      //   tile (I, J) of record k at stile + k * stile_stride + (J * NT + I) * 64, lane l holds doubles 2l, 2l+1
    // pure even-order axis derivatives d^K/dx_a^K r^p (the hyperviscosity operators of rbfbasis_k.jl:9-18): the term list
    // sum_j c_j x^(K-2j) r^(p-2K+2j) is r^(p-K) times a polynomial of degree K/2 in u = (x/r)^2, evaluated by Horner
    int8_t hv_axis[8];         // -1: not of this form
    int8_t hv_half[8];         // K / 2
    int8_t hv_rexp[8];         // p - K (odd, >= 1)
    double hv_c[8][4];         // coefficient of u^k
    OpTables T;
};

// record layout (bytes).  The W' block is stored TRANSPOSED, [Q][ws] doubles with ws = 8 NJ + 4: entry (c, pos) holds
// W'[pos][c] for the non-basic positions pos < nb and w_p[o][c] at pos = rcb + o.  Stage A's lanes (one row each, consecutive
// lanes = consecutive positions) then store consecutive addresses, stage B copies the block into shared memory as it is, and
// both DMMA operand loads of it (t ws + g) are free of bank conflicts for ws == 4 or 12 (mod 16).  Positions no row maps to
// are zeroed by stage A itself (the scratch buffer is recycled from call to call and never cleared by the host).
constexpr int NS2_REC_S = 0;        // double s[3]
constexpr int NS2_REC_XC = 24;      // double xc[3]
constexpr int NS2_REC_ETA = 48;     // double eta[3]
constexpr int NS2_REC_PID = 128;    // int32 node id by position [64]
constexpr int NS2_REC_PERM = 384;   // uint8 position -> stencil slot [64]
constexpr int NS2_REC_HDR = 448;    // bytes of the header (multiple of 16)
constexpr int NS2_REC_W = 512;      // double [Q][ws]

#if defined(NS2_D) && defined(NS2_Q)

namespace {
using namespace nsp;

#ifdef NS2_TIMING
__device__ unsigned long long ns2_prof[16];
#define NS2_T(slot, cond) do { if (cond) { const long long _t = clock64(); atomicAdd(&ns2_prof[slot], (unsigned long long)(_t - tprev)); tprev = _t; } } while (0)
#else
#define NS2_T(slot, cond) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------------------------------
// stage A: pivoted column reduction of [P; g'], one warp per stencil, rows `lane` and `lane + 32`
// ---------------------------------------------------------------------------------------------------------------------
template <int D, int Q, int NN = 0, int NO = 0>
__global__ void __launch_bounds__(128, 4) ns2_pred_kernel(Ns2Args a) {
    constexpr int KS = (Q + 3) / 4, QP = 4 * KS;
    constexpr int CS = (Q + 2) & ~1;                  // published row: Q entries + the reciprocal of the pivot, even
    __shared__ __align__(16) double cand_s[4][2][CS];
    __shared__ __align__(16) double stage_s[4][8 * QP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    const OpTables& T = a.T;
    const int n = NN ? NN : T.n, nops = NO ? NO : T.nops, nb = n - Q;      // NN, NO != 0: compile-time shape (see ns2_solve_kernel)
    double* cand = &cand_s[warp][0][0];
    double* stage = &stage_s[warp][0];
    const int slot0 = lane, slot1 = lane + 32;
    for (int64_t k = blockIdx.x * 4ll + warp; k < a.cnt; k += (int64_t)gridDim.x * 4) {
        const int64_t row = a.row0 + k;
        // ---- scalestencil.jl:10-20 ----
        const int32_t* st = a.stencils + (a.center ? (int64_t)a.center[row] : row) * n;
        const int id0 = st[slot0 < n ? slot0 : 0], id1 = st[slot1 < n ? slot1 : 0];
        double sx0[D], sx1[D], s[D], xcv[D], eta[D];
        bool eta_zero = true;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const double x0 = a.X[(int64_t)id0 * D + c], x1 = a.X[(int64_t)id1 * D + c];
            const double xc = __shfl_sync(FULL, x0, 0);
            const double d0 = x0 - xc, d1 = x1 - xc;
            s[c] = 1.0 / warp_max_nn(fmax(fabs(d0), fabs(d1)));
            sx0[c] = d0 * s[c];
            sx1[c] = d1 * s[c];
            xcv[c] = xc;
            eta[c] = (a.Y[row * D + c] - xc) * s[c];
            eta_zero = eta_zero && (eta[c] == 0.0);
        }
        double p0[Q], p1[Q];
        mono_rows<D>(sx0, p0, std::make_integer_sequence<int, Q>{});
        mono_rows<D>(sx1, p1, std::make_integer_sequence<int, Q>{});
        // rows n .. n+nops-1 carry the polynomial right-hand sides g_o' (reduced by the same column operations)
        const int go0 = slot0 - n, go1 = slot1 - n;
        const bool gown0 = go0 >= 0 && go0 < nops, gown1 = go1 >= 0 && go1 < nops;
        if (slot0 >= n) {
#pragma unroll
            for (int c = 0; c < Q; ++c) p0[c] = 0.0;
        }
        if (slot1 >= n) {
#pragma unroll
            for (int c = 0; c < Q; ++c) p1[c] = 0.0;
        }
        if (eta_zero) {
            auto fill = [&](int go, double* pr) {
                const bool lap = T.kind[go] == RBFFD_OP_LAPLACE;
                const int col = a.gzcol[go];
                const double val = a.gzval[go];
#pragma unroll
                for (int c = 0; c < Q; ++c) {
                    double v = (c == col) ? val : 0.0;
#pragma unroll
                    for (int ax = 0; ax < D; ++ax)
                        if (lap && c == a.lapcol[ax]) v = 2.0 * s[ax] * s[ax];
                    pr[c] = v;
                }
            };
            if (gown0) fill(go0, p0);
            if (gown1) fill(go1, p1);
        } else {
            // general evaluation point: lane c evaluates monomial c for every operator, staged through shared memory
            for (int o = 0; o < nops; ++o)
                if (lane < Q) stage[o * QP + lane] = rhs_poly_entry<D>(T, o, lane, eta, s);
            __syncwarp();
            if (gown0) {
#pragma unroll
                for (int c = 0; c < Q; ++c) p0[c] = stage[go0 * QP + c];
            }
            if (gown1) {
#pragma unroll
                for (int c = 0; c < Q; ++c) p1[c] = stage[go1 * QP + c];
            }
            __syncwarp();
        }
        bool b0 = false, b1 = false;
        int mb0 = 0, mb1 = 0;
        unsigned kmin = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            const unsigned h0 = (unsigned)__double2hiint(p0[j]) & 0x7fffffc0u;
            const unsigned h1 = (unsigned)__double2hiint(p1[j]) & 0x7fffffc0u;
            const unsigned k0 = (slot0 < n && !b0) ? (h0 | (unsigned)slot0) : 0u;
            const unsigned k1 = (slot1 < n && !b1) ? (h1 | (unsigned)slot1) : 0u;
            const unsigned kw = __reduce_max_sync(FULL, max(k0, k1));
            const double r0 = rcp3(p0[j]), r1 = rcp3(p1[j]);       // every candidate inverts its own entry under the search
            double* cw = cand + (j & 1) * CS;
            // the winning slot is in the key, so which of the two register rows is published is a warp-uniform branch
            const int wl = (int)(kw & 31u);
            if (kw >= 64u) {
                if ((kw & 32u) == 0u) {
                    if (lane == wl) {
                        double2* dst = reinterpret_cast<double2*>(cw);
#pragma unroll
                        for (int c = 0; c < CS; c += 2) dst[c >> 1] = make_double2(c < Q ? p0[c] : r0, c + 1 < Q ? p0[c + 1] : r0);
                        b0 = true; mb0 = j;
                    }
                } else {
                    if (lane == wl) {
                        double2* dst = reinterpret_cast<double2*>(cw);
#pragma unroll
                        for (int c = 0; c < CS; c += 2) dst[c >> 1] = make_double2(c < Q ? p1[c] : r1, c + 1 < Q ? p1[c + 1] : r1);
                        b1 = true; mb1 = j;
                    }
                }
            }
            __syncwarp();
            kmin = min(kmin, kw);                                   // < 64: P is rank deficient on this stencil
            const double2* cr = reinterpret_cast<const double2*>(cw);
            const double rinv = cw[Q];
            const double t0 = p0[j] * rinv, t1 = p1[j] * rinv;
#pragma unroll
            for (int c = 0; c < Q; c += 2) {
                const double2 v = cr[c >> 1];
                if (c != j) { p0[c] = fma(-t0, v.x, p0[c]); p1[c] = fma(-t1, v.x, p1[c]); }
                if (c + 1 < Q && c + 1 != j) { p0[c + 1] = fma(-t0, v.y, p0[c + 1]); p1[c + 1] = fma(-t1, v.y, p1[c + 1]); }
            }
            p0[j] = t0;
            p1[j] = t1;
        }
        // positions: non-basic nodes first (0..nb-1, in stencil order), then the basic ones in pivot order
        const unsigned m0 = __ballot_sync(FULL, slot0 < n && !b0), m1 = __ballot_sync(FULL, slot1 < n && !b1);
        const unsigned lt = (1u << lane) - 1u;
        const int pos0 = slot0 < n ? (b0 ? nb + mb0 : __popc(m0 & lt)) : -1;
        const int pos1 = slot1 < n ? (b1 ? nb + mb1 : __popc(m0) + __popc(m1 & lt)) : -1;
        unsigned char* rec = a.rec + k * a.rec_stride;
        int* pid = reinterpret_cast<int*>(rec + NS2_REC_PID);
        unsigned char* perm = rec + NS2_REC_PERM;
        if (lane < D) {
            reinterpret_cast<double*>(rec + NS2_REC_S)[lane] = lane == 0 ? s[0] : (lane == 1 ? s[1] : s[D - 1]);
            reinterpret_cast<double*>(rec + NS2_REC_XC)[lane] = lane == 0 ? xcv[0] : (lane == 1 ? xcv[1] : xcv[D - 1]);
            reinterpret_cast<double*>(rec + NS2_REC_ETA)[lane] = lane == 0 ? eta[0] : (lane == 1 ? eta[1] : eta[D - 1]);
        }
        const int idc = __shfl_sync(FULL, id0, 0);
        if (pos0 >= 0) { pid[pos0] = id0; perm[pos0] = (unsigned char)slot0; }
        else { pid[slot0] = idc; perm[slot0] = 0; }                       // positions n..63: rows / columns nobody uses read the centre
        if (pos1 >= 0) { pid[pos1] = id1; perm[pos1] = (unsigned char)slot1; }
        else { pid[slot1] = idc; perm[slot1] = 0; }
        double* W = reinterpret_cast<double*>(rec + NS2_REC_W);
        const int ws = a.ws;
        const int wrow0 = (slot0 < n && !b0) ? pos0 : (gown0 ? a.rcb + go0 : -1);
        const int wrow1 = (slot1 < n && !b1) ? pos1 : (gown1 ? a.rcb + go1 : -1);
        if (wrow0 >= 0) {
#pragma unroll
            for (int c = 0; c < Q; ++c) W[c * ws + wrow0] = p0[c];
        }
        if (wrow1 >= 0) {
#pragma unroll
            for (int c = 0; c < Q; ++c) W[c * ws + wrow1] = p1[c];
        }
        // positions of the W'^T block that no row maps to ([nb, rcb) and [rcb + nops, ws)) are read by stage B into padded tile
        // rows / columns: they must be finite, so they are zeroed here (the scratch buffer is recycled, never cleared by the host)
        {
            const int z0 = a.rcb - nb, nz = z0 + (ws - a.rcb - nops);
            for (int e = lane; e < Q * nz; e += 32) {
                const int c = e / nz, z = e - c * nz;
                W[c * ws + (z < z0 ? nb + z : a.rcb + nops + (z - z0))] = 0.0;
            }
        }
        if (kmin < 64u && lane == 0) *a.redo = 1;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage B: Phi~ assembly, S = Z' Phi Z, elimination, back substitution, CSR row
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SV_LD = 68;      // row stride of Phi~; == 4 (mod 16): pair stores 69 l + k, 69 l + 68 k and fragment loads 68 g + t conflict-free

#ifndef NS2_PAIR_ILP
#define NS2_PAIR_ILP 4
#endif
#ifndef NS2_SYM
#define NS2_SYM 1
#endif
#ifndef NS2_CFG3_MINB
#define NS2_CFG3_MINB 6
#endif
#ifndef NS2_CFG4_MINB
#define NS2_CFG4_MINB 5
#endif
template <int D, int Q, int NT, int NJ, int NN = 0, int NO = 0>
struct SvCfg {
    static constexpr int KS = (Q + 3) / 4;
    // compile-time n: the Phi~ tile shrinks to the stencil -- row stride = the smallest value >= n that is == 4 (mod 16), n rows
    // (the fragment loads of the padded rows n .. 8 NR - 1 run on into the W' block: finite garbage that stays in padded rows of
    // Y), and the RBF right-hand sides move into the spare columns n .. LD - 1 of the tile when they fit.  Shared memory per
    // CTA: 46.0 -> 34.5 KB at n = 50 (6 CTAs per SM), 50.6 -> 45.3 KB at n = 60 (5 CTAs per SM).
    static constexpr int LD = NN ? ((NN + 11) / 16) * 16 + 4 : SV_LD;
    static constexpr int BSC = NO > 6 ? ((NO + 1) & ~1) : 6;      // row stride of a separate right-hand-side tile (Ys aliases it)
    static constexpr bool BTG = NN != 0 && NO != 0 && LD - NN >= NO;      // right-hand sides in the spare columns of Phi~
    static constexpr int MINB = NN == 0 ? 4 : (LD < SV_LD ? NS2_CFG3_MINB : (BTG ? NS2_CFG4_MINB : 4));
    static constexpr int NBP = 8 * NT;                    // padded null-space dimension
    static constexpr int US = ((8 * NJ + 15) & ~15) + 4;  // row stride of the Y tile, == 4 (mod 16)
    static constexpr int DP = D == 2 ? 2 : 4;
    static constexpr int JJ = NJ > 4 ? 2 : 1;             // tile columns of [S | t] per warp (column J lives in warp J % 4)
    static constexpr int WS = 8 * NJ + 4;                 // row stride of the transposed W' block, == 4 or 12 (mod 16)
    static constexpr int NRT = NN ? (NN + 7) / 8 : 8;     // row tiles of Y
    static constexpr int G = NN ? (NN * LD > 8 * NRT * US ? NN * LD : 8 * NRT * US) : 64 * SV_LD;     // Phi~, later the Y tile, later the exchange buffers of the elimination
    static constexpr int WT = Q * WS;                     // W'^T: [monomial][position], the w_p entries at positions rcb ..
    static constexpr int SC = 64 * DP;
    static constexpr int HDR = NS2_REC_HDR / 8;           // header of the NEXT record (prefetched)
    static constexpr int XN = 64 * D;                     // node coordinates of the NEXT stencil by position (prefetched)
    static_assert(8 * NRT * US <= G && LD >= (NN ? NN : 64), "Y tile must fit into the Phi tile");
    static_assert(NN == 0 || 8 * NRT * LD <= G + Q * WS, "fragment loads of the padded rows must stay inside the W' block");
    static_assert(NJ == NT || NJ == NT + 1, "right-hand sides ride in the last null-space tile column or in one more");
};

// Unpivoted elimination of one (8 NT) x 4 panel of the definite matrix S (static pivot rows pr0 .. pr0+3), one warp, two rows
// per lane (rows lane and lane + 32).  Pbuf [4][PS] holds the panel by column; the transform columns W of the rank-4 update
// X += W X[pivots, :] go to Lb [4][PS], the pivot reciprocals to rinv_s[pr0 ..].  Pivot rows are not scaled (y = RHS_row /
// pivot at the end), so their own entry of W is zero.  Returns the sign-violation bits of the pivots.
template <int ROWS>
__device__ __forceinline__ int gj_panel(const double* __restrict__ Pbuf, double* __restrict__ Lb, double* __restrict__ rinv_s,
                                        int pr0, int sgnbits, int PS) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int sl = pr0 >> 5;                              // slot of the four pivot rows (pr0 is a multiple of 4)
    constexpr int R1 = ROWS - 32;                         // rows held in the second slot
    double av[2][4], w[2][4];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            av[rr][cc] = (rr == 0 || (R1 > 0 && lane < R1)) ? Pbuf[cc * PS + lane + 32 * rr] : 0.0;
            w[rr][cc] = 0.0;
        }
    int bad = 0;
#pragma unroll
    for (int sidx = 0; sidx < 4; ++sidx) {
        const int pl = (pr0 + sidx) & 31;
        double pv[4], wp[4];
#pragma unroll
        for (int cc = sidx; cc < 4; ++cc) pv[cc] = __shfl_sync(FULL, (R1 > 0 && sl) ? av[1][cc] : av[0][cc], pl);
#pragma unroll
        for (int cc = 0; cc < sidx; ++cc) wp[cc] = __shfl_sync(FULL, (R1 > 0 && sl) ? w[1][cc] : w[0][cc], pl);
        bad |= __double2hiint(pv[sidx]) ^ sgnbits;        // S not definite: the pivoted kernel must take over
        const double rinv = rcp3(pv[sidx]);
        if (lane == 0) rinv_s[pr0 + sidx] = rinv;
#pragma unroll
        for (int rr = 0; rr < (R1 > 0 ? 2 : 1); ++rr) {
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
        if (R1 > 0 && lane < R1) Lb[cc * PS + lane + 32] = w[1][cc];
    }
    return bad;
}

// Phi~ by symmetric pairs, split between two thread sets: the thread that owns position l pairs in round k with position
// (l + k) mod n; this call runs rounds k0, k0 + 2, ... <= n/2 (see phs_assemble_t in phs.cuh for the pairing).
template <int D, int DP, int LD, int HP>
__device__ __forceinline__ void phs_assemble_half(const double* __restrict__ Sc, double* __restrict__ G, const double* me, int l, int n,
                                                  bool active, int hp, int k0) {
    double* const grow_l = G + l * LD;
    double* const gcol_l = G + l;
    const int rounds = n >> 1;
    auto phi = [&](int k, int& ib) -> double {
        ib = l + k;
        ib = ib >= n ? ib - n : ib;
        double o[D];
        // (x, y) pairs at stride 16 B, z in its own array behind them: both loads are free of bank conflicts (one point per
        // 32 B cost 8 wavefronts per load instead of 4 and 2)
        const double2 v = *reinterpret_cast<const double2*>(Sc + ib * 2);
        o[0] = v.x; o[1] = v.y;
        if constexpr (D == 3) o[2] = Sc[128 + ib];
        double r2 = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) { const double dd = me[c] - o[c]; r2 = fma(dd, dd, r2); }
        return phs_pow_t<HP>(r2, phs_rsqrt(r2), hp);
    };
    int k = k0;
#if NS2_PAIR_ILP >= 4
    for (; k + 6 <= rounds; k += 8) {                       // four independent dependency chains per trip
        int b0, b1, b2, b3;
        const double v0 = phi(k, b0);
        const double v1 = phi(k + 2, b1);
        const double v2 = phi(k + 4, b2);
        const double v3 = phi(k + 6, b3);
        if (active) {
            grow_l[b0] = v0; gcol_l[b0 * LD] = v0; grow_l[b1] = v1; gcol_l[b1 * LD] = v1;
            grow_l[b2] = v2; gcol_l[b2 * LD] = v2; grow_l[b3] = v3; gcol_l[b3 * LD] = v3;
        }
    }
#endif
    for (; k + 2 <= rounds; k += 4) {
        int b0, b1;
        const double v0 = phi(k, b0);
        const double v1 = phi(k + 2, b1);
        if (active) { grow_l[b0] = v0; gcol_l[b0 * LD] = v0; grow_l[b1] = v1; gcol_l[b1 * LD] = v1; }
    }
    if (k <= rounds) {
        int b0;
        const double v0 = phi(k, b0);
        if (active) { grow_l[b0] = v0; gcol_l[b0 * LD] = v0; }
    }
}

// NN, NO != 0: stencil size and operator count fixed at compile time (the BASELINE configs[2] and configs[3] shapes): the column
// classification of the Y tile, the padding tests, the block-step guards and the trip counts of the node and store phases fold
// to constants (before, 45 % of this kernel's instructions were index arithmetic and predicates on run-time n, n - q, r).
template <int D, int Q, int NT, int NJ, bool SPLIT, int NN = 0, int NO = 0>
__global__ void __launch_bounds__(128, (SvCfg<D, Q, NT, NJ, NN, NO>::MINB)) ns2_solve_kernel(Ns2Args a) {
    using C = SvCfg<D, Q, NT, NJ, NN, NO>;
    static_assert(NN == 0 || SPLIT, "the shrunk Phi~ tile has no room for the exchange buffers of the in-kernel elimination");
    constexpr int LD = C::LD, KS = C::KS, US = C::US, DP = C::DP, NBP = C::NBP, JJ = C::JJ, WS = C::WS;
    constexpr bool SYM = NN != 0 && NO != 0 && NS2_SYM != 0;      // S from the rows of the basic nodes of Y only (see phase B')
    extern __shared__ __align__(16) unsigned char wsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const OpTables& T = a.T;
    const int n = NN ? NN : T.n, nops = NO ? NO : T.nops, nb = n - Q, BS = C::BTG ? LD : (NO ? C::BSC : a.bs);
    double* G = reinterpret_cast<double*>(wsm);
    double* Yb = G;
    double* Wt = G + C::G;                            // [Q][WS]
    double* Sc = Wt + C::WT;
    double* Bt = C::BTG ? G + NN : Sc + C::SC;        // [64][BS] RBF right-hand sides by position (BTG: columns n .. of the Phi~ rows)
    double* Ys = Bt;                                  // solution y, [op][NBP] (Bt is dead by then; 8 * 48 <= 64 * BS needs BS >= 6; not with BTG)
    double* pf = C::BTG ? Sc + C::SC : Bt + 64 * BS;  // [8] chain-rule factor of every operator
    double* hdr = pf + 8;                             // header of the next record
    double* Xn = hdr + C::HDR;                        // [64][D] coordinates of the next stencil's nodes
    int* perm = reinterpret_cast<int*>(Xn + C::XN);   // [64] position -> stencil slot
    const double EPS = 2.220446049250313e-16;
    const double sgn = (((T.p + 1) >> 1) & 1) ? -1.0 : 1.0;           // (-1)^((p+1)/2) S is positive definite
    const int sgnbits = sgn < 0.0 ? (int)0x80000000 : 0;
    const int hp = (T.p - 1) >> 1;
    const int rcb = NN ? (NJ == NT ? ((NN - Q + 3) & ~3) : 8 * NT) : a.rcb;      // first right-hand-side column; also the position of w_p in Wt
    const int NR = (n + 7) >> 3;                      // row tiles of Y
    // exchange buffers of the elimination (alias the Y tile), alternating by block-step parity
    constexpr int PS = 52, UST = 8 * NJ + 4;          // strides == 4 (mod 16)
    double* Pbuf = G;                                 // [4][PS]            owner-private
    double* Lbuf = Pbuf + 4 * PS;                     // [2][4][PS]
    double* Ubuf = Lbuf + 2 * 4 * PS;                 // [2][4][UST]
    double* rinv_s = Ubuf + 2 * 4 * UST;              // [48]
    const bool has2 = JJ > 1 && warp + 4 < NJ;
    const int P = tid & 63;                           // position whose node this thread handles in the node phase

    // header and node coordinates of a record -> shared memory (cp.async: no registers held across the elimination)
    auto fetch_header = [&](const unsigned char* rec) {
        if (tid < NS2_REC_HDR / 16) cp_async16(reinterpret_cast<unsigned char*>(hdr) + 16 * tid, rec + 16 * tid);
    };
    auto fetch_nodes = [&]() {                        // needs the header in shared memory
        if (tid < 64) {
            const int idn = reinterpret_cast<const int*>(reinterpret_cast<const unsigned char*>(hdr) + NS2_REC_PID)[tid];
            const double* src = a.X + (int64_t)idn * D;
            if constexpr (D == 2) cp_async16(Xn + 2 * tid, src);
            else {
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const unsigned d = (unsigned)__cvta_generic_to_shared(Xn + D * tid + c);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src + c) : "memory");
                }
            }
        }
    };
    if ((int64_t)blockIdx.x < a.cnt) {
        fetch_header(a.rec + (int64_t)blockIdx.x * a.rec_stride);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        fetch_nodes();
        cp_async_commit();
    }

    for (int64_t i = blockIdx.x; i < a.cnt; i += gridDim.x) {
#ifdef NS2_TIMING
        long long tprev = clock64();
#endif
        const int64_t row = a.row0 + i;
        const unsigned char* rec = a.rec + i * a.rec_stride;
        cp_async_wait_all();                          // header and nodes of THIS stencil (prefetched during the previous one)
        __syncthreads();
        NS2_T(6, tid == 0);
        // ---- 0. node phase: everything comes from shared memory ----
        const int id = reinterpret_cast<const int*>(reinterpret_cast<const unsigned char*>(hdr) + NS2_REC_PID)[P];
        double sx[D], s[D], eta[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const double xc = hdr[NS2_REC_XC / 8 + c];
            s[c] = hdr[NS2_REC_S / 8 + c];
            eta[c] = hdr[NS2_REC_ETA / 8 + c];
            sx[c] = (Xn[D * P + c] - xc) * s[c];
        }
        if (tid < 64) {
            const int slot = reinterpret_cast<const unsigned char*>(hdr)[NS2_REC_PERM + P];
            if constexpr (!SPLIT) perm[P] = slot;           // split path: the scatter (and the chain-rule factors) belong to the elimination kernel
#pragma unroll
            for (int c = 0; c < D; ++c) Sc[c < 2 ? 2 * P + c : 128 + P] = sx[c];
            if (P < n) {
                G[P * LD + P] = 0.0;
                a.colind[row * n + slot] = id;              // the pattern row in stencil order (generate_operator.jl:171-176)
            }
            if constexpr (!SPLIT) { if (tid < nops) pf[tid] = op_post_factor<D>(T, tid, s); }
        }
        if (P < n) {
            // RBF part of the right-hand sides at this node (generate_operator.jl:123-154); thread set tid/64 takes every
            // other operator (the hyperviscosity closed forms are long term lists)
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
            auto rhs_one = [&](const int o) {
                int order = 0;
#pragma unroll
                for (int c = 0; c < D; ++c) order += T.alpha[o][c];
                double val;
                if (a.hv_axis[o] >= 0) {                    // hyperviscosity closed form: r^(p-K) * poly((x/r)^2)
                    const int ax = a.hv_axis[o];
                    const double xa = ax == 0 ? del[0] : (ax == 1 ? del[1] : del[D - 1]);
                    const double xi = xa * y, u = xi * xi;
                    // Horner from the top of the padded coefficient row (hv_c[o][k] == 0 for k > half): no loop, same value
                    double pv = fma(a.hv_c[o][3], u, a.hv_c[o][2]);
                    pv = fma(pv, u, a.hv_c[o][1]);
                    pv = fma(pv, u, a.hv_c[o][0]);
                    double rr = r;
                    for (int e = 1; e < a.hv_rexp[o]; e += 2) rr *= r2;
                    val = rr * pv;
                } else if (T.kind[o] == RBFFD_OP_DERIV && order > 2) {
                    val = eval_rbf_terms_rinv<D>(T, T.tb[3 * o], T.tb[3 * o + 1], del, r, r2, y);                  // other closed forms: term lists
                } else {
                    val = rhs_rbf_entry_fast<D>(T, o, del, s, r, r2, rp, rp2, rp4);
                }
                Bt[P * BS + o] = val;
            };
            // (unrolling this loop over a compile-time operator index was measured neutral: profiles/r02av)
            for (int o = tid >> 6; o < nops; o += 2) rhs_one(o);
        }
        __syncthreads();                              // Sc complete; header and Xn are consumed
        NS2_T(7, tid == 0);
        // W'^T of this stencil and the header of the next one stream in under the assembly
        {
            const double* Wg = reinterpret_cast<const double*>(rec + NS2_REC_W);
            for (int idx = 2 * tid; idx < C::WT; idx += 256) cp_async16(Wt + idx, Wg + idx);
            if (i + gridDim.x < a.cnt) fetch_header(rec + (int64_t)gridDim.x * a.rec_stride);
            cp_async_commit();
        }
        NS2_T(0, tid == 0);
        // ---- A. Phi~ in permuted order by symmetric pairs: thread set tid/64 runs every other round ----
        {
            const int l = P < n ? P : 0;
            const int k0 = 1 + (tid >> 6);
            switch (hp) {                                   // CTA-uniform
                case 1: phs_assemble_half<D, DP, LD, 1>(Sc, G, sx, l, n, P < n, hp, k0); break;     // r^3
                case 2: phs_assemble_half<D, DP, LD, 2>(Sc, G, sx, l, n, P < n, hp, k0); break;     // r^5
                case 3: phs_assemble_half<D, DP, LD, 3>(Sc, G, sx, l, n, P < n, hp, k0); break;     // r^7
                default: phs_assemble_half<D, DP, LD, -1>(Sc, G, sx, l, n, P < n, hp, k0); break;
            }
        }
        cp_async_wait_all();
        __syncthreads();
        if (i + gridDim.x < a.cnt) fetch_nodes();     // the next stencil's coordinates arrive under phases B .. E
        cp_async_commit();
        NS2_T(1, tid == 0);
        double c[NT][JJ][2];
        if constexpr (SYM) {
            // ---- B'. only what S = Z' Phi Z needs of Y (Phi is symmetric):  S = (Phi_NN/2 - W' T) + (Phi_NN/2 - W' T)'  with
            //          T = Phi~[B, N] - Phi~[B, B] W / 2,  so Y is formed for the ROWS OF THE BASIC NODES only (tile rows RB0 ..,
            //          all tile columns; the factor 1/2 rides on the W operand of the null-space columns), plus the
            //          right-hand-side tile column JR of the other rows (b - Phi~[N, B] w_p).  Per warp: tile columns warp and
            //          warp + 4 of the basic rows; the right-hand-side tiles of the non-basic rows go to the warps that own one
            //          tile column only.  Everything is stored IN PLACE (row stride LD): the rows of the basic nodes and the
            //          columns nb .. of the other rows are dead once every warp has its operands, Phi_NN stays.
            constexpr int NB = NN - Q, RC = NJ == NT ? ((NB + 3) & ~3) : 8 * NT;
            constexpr int RB0 = NB >> 3, NBT = C::NRT - RB0, JR = RC >> 3;
            constexpr int W1 = NJ > 4 ? NJ - 4 : 0, NW1 = 4 - W1, NU = (RB0 + NW1 - 1) / NW1;
            static_assert((RC & 7) + NO <= 8, "right-hand sides must sit in one tile column");
            static_assert(8 * NJ <= (C::BTG ? NN : LD), "T rows are stored in place");
            {
                double cb[NBT][JJ][2], cn[NU][2];
                const int un0 = warp - W1;                  // first right-hand-side tile of a non-basic row tile owned by this warp
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj) {
                    const int J = warp + 4 * jj;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int col = 8 * J + 2 * t + e;
                        const bool isn = col < NB, isr = col >= RC && col < RC + NO, on = jj == 0 ? warp < NJ : has2;
#pragma unroll
                        for (int rr = 0; rr < NBT; ++rr) {
                            const int rw = 8 * (RB0 + rr) + g;
                            cb[rr][jj][e] = (on && isn) ? G[rw * LD + col] : ((on && isr) ? Bt[rw * BS + col - RC] : 0.0);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int I = un0 + u * NW1;
                    const bool on = un0 >= 0 && I < RB0;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int col = 8 * JR + 2 * t + e;
                        cn[u][e] = (on && col >= RC && col < RC + NO) ? Bt[(8 * I + g) * BS + col - RC] : 0.0;
                    }
                }
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    const bool kin = 4 * k + t < Q;
                    const double* wk = Wt + (kin ? 4 * k + t : 0) * WS + g;
                    const int pc = NB + (kin ? 4 * k + t : 0);
                    double af[NBT];
#pragma unroll
                    for (int rr = 0; rr < NBT; ++rr) af[rr] = kin ? -G[(8 * (RB0 + rr) + g) * LD + pc] : 0.0;
#pragma unroll
                    for (int jj = 0; jj < JJ; ++jj) {
                        const int J = warp + 4 * jj;
                        if (jj == 0 ? warp < NJ : has2) {
                            double bf = kin ? wk[8 * J] : 0.0;
                            if (8 * J < NB) bf *= (8 * J + g < NB) ? 0.5 : 1.0;
#pragma unroll
                            for (int rr = 0; rr < NBT; ++rr) dmma884(cb[rr][jj][0], cb[rr][jj][1], af[rr], bf);
                        }
                    }
                    if (un0 >= 0) {
                        const double bfr = kin ? wk[8 * JR] : 0.0;
#pragma unroll
                        for (int u = 0; u < NU; ++u) {
                            const int I = un0 + u * NW1;
                            if (I < RB0) {
                                const double afn = kin ? -G[(8 * I + g) * LD + pc] : 0.0;
                                dmma884(cn[u][0], cn[u][1], afn, bfr);
                            }
                        }
                    }
                }
                __syncthreads();                            // every warp has read its operands of Phi~
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj) {
                    const int J = warp + 4 * jj;
                    if (jj == 0 ? warp < NJ : has2) {
#pragma unroll
                        for (int rr = 0; rr < NBT; ++rr) {
                            const int rw = 8 * (RB0 + rr) + g;
                            double* dst = G + rw * LD + 8 * J + 2 * t;
                            if (8 * (RB0 + rr) >= NB) {         // basic rows only
                                if (rw < NN) *reinterpret_cast<double2*>(dst) = make_double2(cb[rr][jj][0], cb[rr][jj][1]);
                            } else {                            // mixed tile row: a non-basic row keeps its Phi_NN entries
#pragma unroll
                                for (int e = 0; e < 2; ++e)
                                    if (rw < NN && (rw >= NB || 8 * J + 2 * t + e >= NB)) dst[e] = cb[rr][jj][e];
                            }
                        }
                    }
                }
                if (un0 >= 0) {
#pragma unroll
                    for (int u = 0; u < NU; ++u) {
                        const int I = un0 + u * NW1;
                        if (I < RB0) {
                            double* dst = G + (8 * I + g) * LD + 8 * JR + 2 * t;
#pragma unroll
                            for (int e = 0; e < 2; ++e)
                                if (8 * JR + 2 * t + e >= NB) dst[e] = cn[u][e];
                        }
                    }
                }
            }
            __syncthreads();
            NS2_T(2, tid == 0);
            // ---- C'. V = Phi_NN / 2 - W' T (right-hand-side columns: t = Y_N - W' Y_B as before), then S = V + V' through the tile ----
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) {
                const int J = warp + 4 * jj;
                const bool on = jj == 0 ? warp < NJ : has2;
#pragma unroll
                for (int I = 0; I < NT; ++I) {
                    const double2 v = on ? *reinterpret_cast<const double2*>(G + (8 * I + g) * LD + 8 * J + 2 * t) : make_double2(0.0, 0.0);
                    c[I][jj][0] = (8 * J + 2 * t < NB) ? 0.5 * v.x : v.x;
                    c[I][jj][1] = (8 * J + 2 * t + 1 < NB) ? 0.5 * v.y : v.y;
                }
            }
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const bool kin = 4 * k + t < Q;
                const double* wk = Wt + (kin ? 4 * k + t : 0) * WS + g;
                double af[NT], bf[JJ];
#pragma unroll
                for (int I = 0; I < NT; ++I) af[I] = kin ? -wk[8 * I] : 0.0;
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj) bf[jj] = (kin && (jj == 0 ? warp < NJ : has2)) ? G[(NB + 4 * k + t) * LD + 8 * (warp + 4 * jj) + g] : 0.0;
#pragma unroll
                for (int I = 0; I < NT; ++I) dmma884(c[I][0][0], c[I][0][1], af[I], bf[0]);
                if constexpr (JJ > 1) {
                    if (has2) {
#pragma unroll
                        for (int I = 0; I < NT; ++I) dmma884(c[I][1][0], c[I][1][1], af[I], bf[1]);
                    }
                }
            }
            // V in place (this warp's own tile columns of the rows < nb: nobody else reads them in this phase)
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) {
                const int J = warp + 4 * jj;
                if ((jj == 0 ? warp < NJ : has2) && 8 * J < NB) {
#pragma unroll
                    for (int I = 0; I < NT; ++I) {
                        const int rw = 8 * I + g;
                        double* dst = G + rw * LD + 8 * J + 2 * t;
                        if (8 * I + 8 <= NB && 8 * J + 8 <= NB) *reinterpret_cast<double2*>(dst) = make_double2(c[I][jj][0], c[I][jj][1]);
                        else {
#pragma unroll
                            for (int e = 0; e < 2; ++e)
                                if (rw < NB && 8 * J + 2 * t + e < NB) dst[e] = c[I][jj][e];
                        }
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) {
                const int J = warp + 4 * jj;
                if ((jj == 0 ? warp < NJ : has2) && 8 * J < NB) {
#pragma unroll
                    for (int I = 0; I < NT; ++I) {
                        const int rw = 8 * I + g;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int col = 8 * J + 2 * t + e;
                            if ((8 * I + 8 <= NB && 8 * J + 8 <= NB) || (rw < NB && col < NB)) c[I][jj][e] += G[col * LD + rw];
                        }
                    }
                }
            }
            // identity padding outside the nb x nb block; right-hand-side columns of padded rows are zero
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) {
                const int J = warp + 4 * jj;
#pragma unroll
                for (int I = 0; I < NT; ++I) {
                    if (8 * I + 8 > NB || 8 * J + 8 > NB) {
                        const int rw = 8 * I + g;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int col = 8 * J + 2 * t + e;
                            if (col >= RC) { if (rw >= NB) c[I][jj][e] = 0.0; }
                            else if (rw >= NB || col >= NB) c[I][jj][e] = rw == col ? sgn : 0.0;
                        }
                    }
                }
            }
        } else {
            // ---- B. Y = Phi~[:, N] - Phi~[:, B] W : row tiles 2*warp, 2*warp+1 (right-hand-side columns start from b, use w_p) ----
            {
                double cy[2][NJ][2];
                const bool r1ok = 2 * warp + 1 < NR;
                if (2 * warp < NR) {
                    const double* rowp[2];
                    const double* browp[2];
    #pragma unroll
                    for (int ii = 0; ii < 2; ++ii) {
                        const int rr = 8 * (2 * warp + ((ii == 0 || r1ok) ? ii : 0)) + g;
                        rowp[ii] = G + rr * LD;
                        browp[ii] = Bt + rr * BS;
                    }
    #pragma unroll
                    for (int J = 0; J < NJ; ++J)
    #pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int col = 8 * J + 2 * t + e;
                            const bool isn = col < nb, isr = col >= rcb && col < rcb + nops;
    #pragma unroll
                            for (int ii = 0; ii < 2; ++ii) cy[ii][J][e] = isn ? rowp[ii][col] : (isr ? browp[ii][col - rcb] : 0.0);
                        }
    #pragma unroll
                    for (int k = 0; k < KS; ++k) {
                        const bool kin = 4 * k + t < Q;
                        const int pc = kin ? nb + 4 * k + t : 0;
                        const double* wk = Wt + (kin ? 4 * k + t : 0) * WS + g;
                        double af[2];
    #pragma unroll
                        for (int ii = 0; ii < 2; ++ii) af[ii] = kin ? -rowp[ii][pc] : 0.0;
    #pragma unroll
                        for (int J = 0; J < NJ; ++J) {
                            const double bf = kin ? wk[8 * J] : 0.0;
                            dmma884(cy[0][J][0], cy[0][J][1], af[0], bf);
                            if (r1ok) dmma884(cy[1][J][0], cy[1][J][1], af[1], bf);
                        }
                    }
                }
                __syncthreads();                              // Phi~ is dead in every warp: the Y tile reuses its storage
                if (2 * warp < NR) {
    #pragma unroll
                    for (int J = 0; J < NJ; ++J)
    #pragma unroll
                        for (int ii = 0; ii < 2; ++ii)
                            if (ii == 0 || r1ok)
                                *reinterpret_cast<double2*>(Yb + (8 * (2 * warp + ii) + g) * US + 8 * J + 2 * t) = make_double2(cy[ii][J][0], cy[ii][J][1]);
                }
            }
            __syncthreads();
            NS2_T(2, tid == 0);
            // ---- C. [S | t] = Y[N, :] - W' Y[B, :] : this warp owns tile columns warp and warp + 4 ----
            {
    #pragma unroll
                for (int jj = 0; jj < JJ; ++jj) {
                    const int J = warp + 4 * jj;
    #pragma unroll
                    for (int I = 0; I < NT; ++I) {
                        const double2 v = (jj == 0 || has2) ? *reinterpret_cast<const double2*>(Yb + (8 * I + g) * US + 8 * J + 2 * t) : make_double2(0.0, 0.0);
                        c[I][jj][0] = v.x; c[I][jj][1] = v.y;
                    }
                }
    #pragma unroll
                for (int k = 0; k < KS; ++k) {
                    const bool kin = 4 * k + t < Q;
                    const double* wk = Wt + (kin ? 4 * k + t : 0) * WS + g;
                    double af[NT], bf[JJ];
    #pragma unroll
                    for (int I = 0; I < NT; ++I) af[I] = kin ? -wk[8 * I] : 0.0;
    #pragma unroll
                    for (int jj = 0; jj < JJ; ++jj) bf[jj] = (kin && (jj == 0 || has2)) ? Yb[(nb + 4 * k + t) * US + 8 * (warp + 4 * jj) + g] : 0.0;
    #pragma unroll
                    for (int I = 0; I < NT; ++I) dmma884(c[I][0][0], c[I][0][1], af[I], bf[0]);
                    if constexpr (JJ > 1) {
                        if (has2) {
    #pragma unroll
                            for (int I = 0; I < NT; ++I) dmma884(c[I][1][0], c[I][1][1], af[I], bf[1]);
                        }
                    }
                }
                // identity padding outside the nb x nb block; right-hand-side columns of padded rows are zero
    #pragma unroll
                for (int jj = 0; jj < JJ; ++jj) {
                    const int J = warp + 4 * jj;
    #pragma unroll
                    for (int I = 0; I < NT; ++I) {
                        if (8 * I + 8 > nb || 8 * J + 8 > nb) {
                            const int rw = 8 * I + g;
    #pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int col = 8 * J + 2 * t + e;
                                if (col >= rcb) { if (rw >= nb) c[I][jj][e] = 0.0; }
                                else if (rw >= nb || col >= nb) c[I][jj][e] = rw == col ? sgn : 0.0;
                            }
                        }
                    }
                }
            }
        }
        if constexpr (SPLIT) {
            // split path: [S | t] leaves as accumulator tiles (one coalesced 512-byte store per tile); the elimination runs
            // in ns2_elim_kernel with two warps per stencil and twice as many stencils in flight per SM
            double* Sg = a.stile + i * a.stile_stride;
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) {
                const int J = warp + 4 * jj;
                if (jj == 0 || has2) {
#pragma unroll
                    for (int I = 0; I < NT; ++I)
                        *reinterpret_cast<double2*>(Sg + (J * NT + I) * 64 + 2 * lane) = make_double2(c[I][jj][0], c[I][jj][1]);
                }
            }
            NS2_T(3, tid == 0);
            continue;                                     // the barrier at the top of the loop separates the Y tile from the next assembly
        }
        __syncthreads();                                  // the Y tile is dead: its storage becomes the exchange buffers
        NS2_T(3, tid == 0);
        // ---- D. blocked Gauss-Jordan WITHOUT pivoting on the definite S (static pivot rows 4kb .. 4kb+3), with one
        //         step of lookahead: in step kb the owner of panel kb+1 updates that tile column first and eliminates the
        //         next panel while the other warps are still applying update kb ----
        int bad = 0;
        auto dump_panel = [&](auto JPc, auto Hc) {        // the 4 panel columns (tile column Jp, half h) of all 8 NT rows
            constexpr int Jp = decltype(JPc)::value, h = decltype(Hc)::value, jp = Jp >> 2;
            if constexpr (jp < JJ) {
                if ((t >> 1) == h) {
                    double* pw = Pbuf + (2 * (t & 1)) * PS + g;
#pragma unroll
                    for (int I = 0; I < NT; ++I) {
                        pw[8 * I] = c[I][jp][0];
                        pw[PS + 8 * I] = c[I][jp][1];
                    }
                }
            }
            __syncwarp();
        };
        auto dump_pivot_rows = [&](auto KBc) {            // raw pivot rows of step kb: tile row kb>>1, lanes with g>>2 == kb&1
            constexpr int kb = decltype(KBc)::value, Jp = kb >> 1, h = kb & 1, jlo = h == 0 ? Jp : Jp + 1;
            double* Ub = Ubuf + (kb & 1) * 4 * UST;
            if ((g >> 2) == h) {
#pragma unroll
                for (int jj = 0; jj < JJ; ++jj) {
                    const int J = warp + 4 * jj;
                    if (J >= jlo && (jj == 0 || has2))
                        *reinterpret_cast<double2*>(Ub + (g & 3) * UST + 8 * J + 2 * t) = make_double2(c[Jp][jj][0], c[Jp][jj][1]);
                }
            }
        };
        if (warp == 0) {
            dump_panel(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
            bad |= gj_panel<8 * NT>(Pbuf, Lbuf, rinv_s, 0, sgnbits, PS);
        }
        dump_pivot_rows(std::integral_constant<int, 0>{});
        __syncthreads();
        auto gj_step = [&](auto KBc) {
            constexpr int kb = decltype(KBc)::value;
            constexpr int Jp = kb >> 1, h = kb & 1, jlo = h == 0 ? Jp : Jp + 1;
            constexpr int kn = kb + 1, Jn = (kn >> 1) < NT ? (kn >> 1) : 0, hn = kn & 1, jn = Jn >> 2;   // next panel: tile column Jn, local index jn
            const double* Lb = Lbuf + (kb & 1) * 4 * PS;
            const double* Ub = Ubuf + (kb & 1) * 4 * UST;
            const bool next = kn < 2 * NT && 4 * kn < nb;
            double af[NT];
#pragma unroll
            for (int I = 0; I < NT; ++I) af[I] = Lb[t * PS + 8 * I + g];
            auto update = [&](auto JJc) {
                constexpr int jj = decltype(JJc)::value;
                if constexpr (jj < JJ) {
                    const int J = warp + 4 * jj;
                    if (J >= jlo && (jj == 0 || has2)) {
                        const double bf = Ub[t * UST + 8 * J + g];
#pragma unroll
                        for (int I = 0; I < NT; ++I) dmma884(c[I][jj][0], c[I][jj][1], af[I], bf);
                    }
                }
            };
            if (next && warp == (Jn & 3)) {
                update(std::integral_constant<int, jn>{});
                dump_panel(std::integral_constant<int, Jn>{}, std::integral_constant<int, hn>{});
                bad |= gj_panel<8 * NT>(Pbuf, Lbuf + (kn & 1) * 4 * PS, rinv_s, 4 * kn, sgnbits, PS);
                update(std::integral_constant<int, 1 - jn>{});
            } else {
                update(std::integral_constant<int, 0>{});
                update(std::integral_constant<int, 1>{});
            }
            if (next) dump_pivot_rows(std::integral_constant<int, (kn < 2 * NT ? kn : 0)>{});
            __syncthreads();
        };
        [&]<int... KB>(std::integer_sequence<int, KB...>) {
            (([&] { if (4 * KB < nb) gj_step(std::integral_constant<int, KB>{}); }()), ...);
        }(std::make_integer_sequence<int, 2 * NT>{});
        if (bad < 0) *a.redo = 1;
        NS2_T(4, tid == 0);                               // every pivot reciprocal is published; Bt is dead (Ys aliases it)
        // y = RHS_row / pivot_row  (right-hand-side column rcb + o lives in tile (rcb + o) / 8)
#pragma unroll
        for (int jj = 0; jj < JJ; ++jj) {
            const int J = warp + 4 * jj;
            if (8 * J + 8 > rcb && 8 * J < rcb + nops && (jj == 0 || has2)) {    // warp-uniform: tile columns that hold right-hand sides
#pragma unroll
                for (int I = 0; I < NT; ++I) {
                    const int rw = 8 * I + g;
                    const double ri = rinv_s[rw < nb ? rw : 0];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int o = 8 * J + 2 * t + e - rcb;
                        if (rw < nb && o >= 0 && o < nops) Ys[o * NBP + rw] = c[I][jj][e] * ri;
                    }
                }
            }
        }
        __syncthreads();
        // ---- E. w[N] = y, w[B] = w_p - W y; rescale and scatter into the CSR row (generate_operator.jl:161-182) ----
        {
            bool fin = true;
            const double INF = __longlong_as_double(0x7ff0000000000000ll);
            // non-basic nodes: w = y.  Thread set tid/64 takes every other operator, position P = tid % 64
            if (P < nb) {
                const int dst = perm[P];
                for (int o = tid >> 6; o < nops; o += 2) {
                    const double wv = Ys[o * NBP + P] * pf[o];
                    fin = fin && (fabs(wv) < INF);          // a zero pivot shows up as a non-finite weight
                    a.vals[((int64_t)o * a.M + row) * n + dst] = wv;
                }
            }
            // basic nodes: the Q x nops block  w_p - W' y  as one DMMA row tile per warp (rows = basic nodes 8 warp + g,
            // columns = operators 2t, 2t+1, k = non-basic nodes)
            if (8 * warp < Q) {
                const int cc = 8 * warp + g;
                const bool rin = cc < Q;
                const double* wrow = Wt + (rin ? cc : 0) * WS;        // W'^T row of this basic node: [position]
                double c0 = (rin && 2 * t < nops) ? wrow[rcb + 2 * t] : 0.0;
                double c1 = (rin && 2 * t + 1 < nops) ? wrow[rcb + 2 * t + 1] : 0.0;
                const double* ycol = Ys + (g < nops ? g : 0) * NBP;
#pragma unroll 2
                for (int k0 = 0; k0 < nb; k0 += 4) {
                    const int aa = k0 + t;
                    const bool kin = aa < nb;
                    const double af = (rin && kin) ? -wrow[kin ? aa : 0] : 0.0;
                    const double bf = (g < nops && kin) ? ycol[kin ? aa : 0] : 0.0;
                    dmma884(c0, c1, af, bf);
                }
                if (rin) {
                    const int dst = perm[nb + cc];
                    if (2 * t < nops) {
                        const double wv = c0 * pf[2 * t];
                        fin = fin && (fabs(wv) < INF);
                        a.vals[((int64_t)(2 * t) * a.M + row) * n + dst] = wv;
                    }
                    if (2 * t + 1 < nops) {
                        const double wv = c1 * pf[2 * t + 1];
                        fin = fin && (fabs(wv) < INF);
                        a.vals[((int64_t)(2 * t + 1) * a.M + row) * n + dst] = wv;
                    }
                }
            }
            if (!fin) *a.redo = 1;
        }
        NS2_T(5, tid == 0);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage C (split path): unpivoted blocked Gauss-Jordan on [S | t], back substitution, CSR row.  One CTA of TWO warps per
// stencil, tile column J in warp J % 2, the matrix register-resident from the first load to the last pivot; 8 CTAs per SM.
// ---------------------------------------------------------------------------------------------------------------------
template <int D, int Q, int NT, int NJ>
__global__ void __launch_bounds__(64, 8) ns2_elim_kernel(Ns2Args a) {
    constexpr int NW = 2, JJ = (NJ + NW - 1) / NW, NBP = 8 * NT, WS = 8 * NJ + 4;
    constexpr int PS = 52, UST = 8 * NJ + 4;          // strides == 4 (mod 16)
    extern __shared__ __align__(16) unsigned char esm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const OpTables& T = a.T;
    const int n = T.n, nops = T.nops, nb = n - Q;
    double* Pbuf = reinterpret_cast<double*>(esm);    // [4][PS]            owner-private
    double* Lbuf = Pbuf + 4 * PS;                     // [2][4][PS]
    double* Ubuf = Lbuf + 2 * 4 * PS;                 // [2][4][UST]
    double* rinv_s = Ubuf + 2 * 4 * UST;              // [48]
    double* Wt = rinv_s + 48;                         // [Q][WS]
    double* Ys = Wt + Q * WS;                         // [8][NBP]
    double* hdr = Ys + 8 * NBP;                       // record header
    double* pf = hdr + NS2_REC_HDR / 8;               // [8]
    int* perm = reinterpret_cast<int*>(pf + 8);       // [64]
    const double sgn = (((T.p + 1) >> 1) & 1) ? -1.0 : 1.0;
    const int sgnbits = sgn < 0.0 ? (int)0x80000000 : 0;
    const int rcb = a.rcb;
    for (int64_t i = blockIdx.x; i < a.cnt; i += gridDim.x) {
        const int64_t row = a.row0 + i;
        const unsigned char* rec = a.rec + i * a.rec_stride;
        // header and W'^T are needed only after the elimination: they stream in under it
        if (tid < NS2_REC_HDR / 16) cp_async16(reinterpret_cast<unsigned char*>(hdr) + 16 * tid, rec + 16 * tid);
        {
            const double* Wg = reinterpret_cast<const double*>(rec + NS2_REC_W);
            for (int idx = 2 * tid; idx < Q * WS; idx += 128) cp_async16(Wt + idx, Wg + idx);
        }
        cp_async_commit();
        const double* Sg = a.stile + i * a.stile_stride;
        double c[NT][JJ][2];
#pragma unroll
        for (int jj = 0; jj < JJ; ++jj) {
            const int J = warp + NW * jj;
#pragma unroll
            for (int I = 0; I < NT; ++I) {
                const double2 v = J < NJ ? *reinterpret_cast<const double2*>(Sg + (J * NT + I) * 64 + 2 * lane) : make_double2(0.0, 0.0);
                c[I][jj][0] = v.x; c[I][jj][1] = v.y;
            }
        }
        int bad = 0;
        auto dump_panel = [&](auto JPc, auto Hc) {        // the 4 panel columns (tile column Jp, half h) of all 8 NT rows
            constexpr int Jp = decltype(JPc)::value, h = decltype(Hc)::value, jp = Jp / NW;
            if ((t >> 1) == h) {
                double* pw = Pbuf + (2 * (t & 1)) * PS + g;
#pragma unroll
                for (int I = 0; I < NT; ++I) {
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
                for (int jj = 0; jj < JJ; ++jj) {
                    const int J = warp + NW * jj;
                    if (J >= jlo && J < NJ)
                        *reinterpret_cast<double2*>(Ub + (g & 3) * UST + 8 * J + 2 * t) = make_double2(c[Jp][jj][0], c[Jp][jj][1]);
                }
            }
        };
        if (warp == 0) {
            dump_panel(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
            bad |= gj_panel<8 * NT>(Pbuf, Lbuf, rinv_s, 0, sgnbits, PS);
        }
        dump_pivot_rows(std::integral_constant<int, 0>{});
        __syncthreads();
        auto gj_step = [&](auto KBc) {
            constexpr int kb = decltype(KBc)::value;
            constexpr int Jp = kb >> 1, h = kb & 1, jlo = h == 0 ? Jp : Jp + 1;
            constexpr int kn = kb + 1, Jn = (kn >> 1) < NT ? (kn >> 1) : 0, hn = kn & 1, jn = Jn / NW;
            const double* Lb = Lbuf + (kb & 1) * 4 * PS;
            const double* Ub = Ubuf + (kb & 1) * 4 * UST;
            const bool next = kn < 2 * NT && 4 * kn < nb;
            double af[NT];
#pragma unroll
            for (int I = 0; I < NT; ++I) af[I] = Lb[t * PS + 8 * I + g];
            auto update = [&](auto JJc) {
                constexpr int jj = decltype(JJc)::value;
                const int J = warp + NW * jj;
                if (J >= jlo && J < NJ) {
                    const double bf = Ub[t * UST + 8 * J + g];
#pragma unroll
                    for (int I = 0; I < NT; ++I) dmma884(c[I][jj][0], c[I][jj][1], af[I], bf);
                }
            };
            if (next && warp == (Jn % NW)) {
                update(std::integral_constant<int, jn>{});
                dump_panel(std::integral_constant<int, Jn>{}, std::integral_constant<int, hn>{});
                bad |= gj_panel<8 * NT>(Pbuf, Lbuf + (kn & 1) * 4 * PS, rinv_s, 4 * kn, sgnbits, PS);
                [&]<int... JX>(std::integer_sequence<int, JX...>) {
                    (([&] { if constexpr (JX != jn) update(std::integral_constant<int, JX>{}); }()), ...);
                }(std::make_integer_sequence<int, JJ>{});
            } else {
                [&]<int... JX>(std::integer_sequence<int, JX...>) {
                    ((update(std::integral_constant<int, JX>{})), ...);
                }(std::make_integer_sequence<int, JJ>{});
            }
            if (next) dump_pivot_rows(std::integral_constant<int, (kn < 2 * NT ? kn : 0)>{});
            __syncthreads();
        };
        [&]<int... KB>(std::integer_sequence<int, KB...>) {
            (([&] { if (4 * KB < nb) gj_step(std::integral_constant<int, KB>{}); }()), ...);
        }(std::make_integer_sequence<int, 2 * NT>{});
        if (bad < 0) *a.redo = 1;
        // y = RHS_row / pivot_row  (right-hand-side column rcb + o lives in tile (rcb + o) / 8)
#pragma unroll
        for (int jj = 0; jj < JJ; ++jj) {
            const int J = warp + NW * jj;
            if (8 * J + 8 > rcb && 8 * J < rcb + nops && J < NJ) {       // warp-uniform: tile columns that hold right-hand sides
#pragma unroll
                for (int I = 0; I < NT; ++I) {
                    const int rw = 8 * I + g;
                    const double ri = rinv_s[rw < nb ? rw : 0];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int o = 8 * J + 2 * t + e - rcb;
                        if (rw < nb && o >= 0 && o < nops) Ys[o * NBP + rw] = c[I][jj][e] * ri;
                    }
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();
        perm[tid] = reinterpret_cast<const unsigned char*>(hdr)[NS2_REC_PERM + tid];
        if (tid < nops) {
            double s[D];
#pragma unroll
            for (int cdim = 0; cdim < D; ++cdim) s[cdim] = hdr[NS2_REC_S / 8 + cdim];
            pf[tid] = op_post_factor<D>(T, tid, s);
        }
        __syncthreads();
        // ---- w[N] = y, w[B] = w_p - W y; rescale and scatter into the CSR row (generate_operator.jl:161-182) ----
        {
            bool fin = true;
            const double INF = __longlong_as_double(0x7ff0000000000000ll);
            if (tid < nb) {
                const int dst = perm[tid];
                for (int o = 0; o < nops; ++o) {
                    const double wv = Ys[o * NBP + tid] * pf[o];
                    fin = fin && (fabs(wv) < INF);          // a zero pivot shows up as a non-finite weight
                    a.vals[((int64_t)o * a.M + row) * n + dst] = wv;
                }
            }
            for (int wt = warp; 8 * wt < Q; wt += NW) {     // basic nodes: one DMMA row tile per trip (see ns2_solve_kernel)
                const int cc = 8 * wt + g;
                const bool rin = cc < Q;
                const double* wrow = Wt + (rin ? cc : 0) * WS;
                double c0 = (rin && 2 * t < nops) ? wrow[rcb + 2 * t] : 0.0;
                double c1 = (rin && 2 * t + 1 < nops) ? wrow[rcb + 2 * t + 1] : 0.0;
                const double* ycol = Ys + (g < nops ? g : 0) * NBP;
#pragma unroll 2
                for (int k0 = 0; k0 < nb; k0 += 4) {
                    const int aa = k0 + t;
                    const bool kin = aa < nb;
                    const double af = (rin && kin) ? -wrow[kin ? aa : 0] : 0.0;
                    const double bf = (g < nops && kin) ? ycol[kin ? aa : 0] : 0.0;
                    dmma884(c0, c1, af, bf);
                }
                if (rin) {
                    const int dst = perm[nb + cc];
                    if (2 * t < nops) {
                        const double wv = c0 * pf[2 * t];
                        fin = fin && (fabs(wv) < INF);
                        a.vals[((int64_t)(2 * t) * a.M + row) * n + dst] = wv;
                    }
                    if (2 * t + 1 < nops) {
                        const double wv = c1 * pf[2 * t + 1];
                        fin = fin && (fabs(wv) < INF);
                        a.vals[((int64_t)(2 * t + 1) * a.M + row) * n + dst] = wv;
                    }
                }
            }
            if (!fin) *a.redo = 1;
        }
        __syncthreads();                                  // Wt, Ys, the header and the exchange buffers are reused by the next stencil
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage C, second form: ONE WARP per stencil and no barrier at all.  [S | t] stays in the warp's accumulator registers from
// the load to the last pivot; the elimination runs by 4 x 4 block pivots (block_gj_warp, nullspace.cuh): the pivot block is
// inverted redundantly in every lane and applied by DMMAs, so the scalar panel chain (4 pivots x search/scale/update per
// block step), the panel and pivot-row dumps and the CTA barrier of every block step of ns2_elim_kernel are gone.  Warps of
// a CTA are independent (each has its own slice of shared memory); occupancy is set by the registers (NT x NJ x 2 doubles).
// ---------------------------------------------------------------------------------------------------------------------
template <int D, int Q, int NT, int NJ>
struct E1Cfg {
    static constexpr int NBP = 8 * NT, WS = 8 * NJ + 4;
    static constexpr int DSM = 32, WT = Q * WS, YS = 8 * NBP, HDR = NS2_REC_HDR / 8, PF = 8;
    static constexpr int DOUBLES = DSM + WT + YS + HDR + PF;
    static constexpr int BYTES = (DOUBLES * 8 + 64 * 4 + 15) & ~15;       // + perm[64]
    static constexpr int WARPS = NT * NJ > 25 ? 3 : 4;                    // warps per CTA
    static constexpr int MINB = NT * NJ > 25 ? 4 : 3;                     // CTAs per SM the register budget is set for
};

template <int D, int Q, int NT, int NJ, int NN = 0, int NO = 0>
__global__ void __launch_bounds__(32 * E1Cfg<D, Q, NT, NJ>::WARPS, E1Cfg<D, Q, NT, NJ>::MINB) ns2_elim1_kernel(Ns2Args a) {
    using C = E1Cfg<D, Q, NT, NJ>;
    constexpr int NBP = C::NBP, WS = C::WS, NW = C::WARPS;
    extern __shared__ __align__(16) unsigned char esm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const OpTables& T = a.T;
    const int n = NN ? NN : T.n, nops = NO ? NO : T.nops, nb = n - Q;
    double* base = reinterpret_cast<double*>(esm + (size_t)warp * C::BYTES);
    double* dsm = base;                               // [2][16] pivot blocks
    double* Wt = dsm + C::DSM;                        // [Q][WS]
    double* Ys = Wt + C::WT;                          // [8][NBP]
    double* hdr = Ys + C::YS;                         // record header
    double* pf = hdr + C::HDR;                        // [8]
    int* perm = reinterpret_cast<int*>(pf + C::PF);   // [64]
    const double sgn = (((T.p + 1) >> 1) & 1) ? -1.0 : 1.0;
    const int sgnbits = sgn < 0.0 ? (int)0x80000000 : 0;
    const int rcb = NN ? (NJ == NT ? ((NN - Q + 3) & ~3) : 8 * NT) : a.rcb;
    for (int64_t i = (int64_t)blockIdx.x * NW + warp; i < a.cnt; i += (int64_t)gridDim.x * NW) {
        const int64_t row = a.row0 + i;
        const unsigned char* rec = a.rec + i * a.rec_stride;
        // header and W'^T are needed only after the elimination: they stream in under it
        if (lane < NS2_REC_HDR / 16) cp_async16(reinterpret_cast<unsigned char*>(hdr) + 16 * lane, rec + 16 * lane);
        {
            const double* Wg = reinterpret_cast<const double*>(rec + NS2_REC_W);
            for (int idx = 2 * lane; idx < Q * WS; idx += 64) cp_async16(Wt + idx, Wg + idx);
        }
        cp_async_commit();
        const double* Sg = a.stile + i * a.stile_stride;
        double c[NT][NJ][2];
#pragma unroll
        for (int J = 0; J < NJ; ++J)
#pragma unroll
            for (int I = 0; I < NT; ++I) {
                const double2 v = __ldcs(reinterpret_cast<const double2*>(Sg + (J * NT + I) * 64 + 2 * lane));
                c[I][J][0] = v.x; c[I][J][1] = v.y;
            }
        const int bad = block_gj_warp<NT, NJ, NT, NJ, (NT * NJ <= 25)>(c, nb, dsm, sgnbits);
        if (bad < 0) *a.redo = 1;
        // the matrix is now [I | y]: y of operator o sits in column rcb + o
#pragma unroll
        for (int J = 0; J < NJ; ++J) {
            if (8 * J + 8 > rcb && 8 * J < rcb + nops) {             // warp-uniform: tile columns that hold right-hand sides
#pragma unroll
                for (int I = 0; I < NT; ++I) {
                    const int rw = 8 * I + g;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int o = 8 * J + 2 * t + e - rcb;
                        if (rw < nb && o >= 0 && o < nops) Ys[o * NBP + rw] = c[I][J][e];
                    }
                }
            }
        }
        cp_async_wait_all();
        __syncwarp();
        perm[lane] = reinterpret_cast<const unsigned char*>(hdr)[NS2_REC_PERM + lane];
        perm[lane + 32] = reinterpret_cast<const unsigned char*>(hdr)[NS2_REC_PERM + lane + 32];
        if (lane < nops) {
            double s[D];
#pragma unroll
            for (int cdim = 0; cdim < D; ++cdim) s[cdim] = hdr[NS2_REC_S / 8 + cdim];
            pf[lane] = op_post_factor<D>(T, lane, s);
        }
        __syncwarp();
        // ---- w[N] = y, w[B] = w_p - W y; rescale and scatter into the CSR row (generate_operator.jl:161-182) ----
        {
            bool fin = true;
            const double INF = __longlong_as_double(0x7ff0000000000000ll);
            for (int P = lane; P < nb; P += 32) {
                const int dst = perm[P];
                for (int o = 0; o < nops; ++o) {
                    const double wv = Ys[o * NBP + P] * pf[o];
                    fin = fin && (fabs(wv) < INF);          // a zero pivot shows up as a non-finite weight
                    a.vals[((int64_t)o * a.M + row) * n + dst] = wv;
                }
            }
            for (int wt = 0; 8 * wt < Q; ++wt) {            // basic nodes: one DMMA row tile per trip (see ns2_solve_kernel)
                const int cc = 8 * wt + g;
                const bool rin = cc < Q;
                const double* wrow = Wt + (rin ? cc : 0) * WS;
                double c0 = (rin && 2 * t < nops) ? wrow[rcb + 2 * t] : 0.0;
                double c1 = (rin && 2 * t + 1 < nops) ? wrow[rcb + 2 * t + 1] : 0.0;
                const double* ycol = Ys + (g < nops ? g : 0) * NBP;
#pragma unroll 2
                for (int k0 = 0; k0 < nb; k0 += 4) {
                    const int aa = k0 + t;
                    const bool kin = aa < nb;
                    const double af = (rin && kin) ? -wrow[kin ? aa : 0] : 0.0;
                    const double bf = (g < nops && kin) ? ycol[kin ? aa : 0] : 0.0;
                    dmma884(c0, c1, af, bf);
                }
                if (rin) {
                    const int dst = perm[nb + cc];
                    if (2 * t < nops) {
                        const double wv = c0 * pf[2 * t];
                        fin = fin && (fabs(wv) < INF);
                        a.vals[((int64_t)(2 * t) * a.M + row) * n + dst] = wv;
                    }
                    if (2 * t + 1 < nops) {
                        const double wv = c1 * pf[2 * t + 1];
                        fin = fin && (fabs(wv) < INF);
                        a.vals[((int64_t)(2 * t + 1) * a.M + row) * n + dst] = wv;
                    }
                }
            }
            if (!fin) *a.redo = 1;
        }
        __syncwarp();                                     // Wt, Ys, the header and perm are reused by the next stencil
    }
}

// shapes with compile-time (n, operator count): BASELINE configs[2] (2-D, n = 50, degree 4, four operators) and configs[3] / [4]
// (3-D, n = 60, degree 3, four operators).  RBFFD_NS2_SPECIALIZE = bit mask (1 solve, 2 elimination, 4 column reduction; 0 keeps the generic instances: A/B comparisons).
template <int D, int Q, int NT, int NJ>
struct Ns2Shape {
    static constexpr bool cfg3 = D == 2 && Q == 15 && NT == 5 && NJ == 5;
    static constexpr bool cfg4 = D == 3 && Q == 20 && NT == 5 && NJ == 6;
    static constexpr int NN = cfg3 ? 50 : (cfg4 ? 60 : 0), NO = (cfg3 || cfg4) ? 4 : 0;
    static bool matches(const OpTables& T, int which) {       // which: 1 = solve, 2 = elimination, 4 = column reduction kernel (bit mask in the env)
        static const int mask = [] { const char* e = getenv("RBFFD_NS2_SPECIALIZE"); return e ? atoi(e) : 7; }();
        return NN != 0 && (mask & which) && T.n == NN && T.nops == NO;
    }
};

template <int D, int Q, int NT, int NJ>
int launch_elim1(rbffd_context* ctx, Ns2Args& a) {
    using C = E1Cfg<D, Q, NT, NJ>;
    using SH = Ns2Shape<D, Q, NT, NJ>;
    const size_t smem = (size_t)C::BYTES * C::WARPS;
    auto kern = ns2_elim1_kernel<D, Q, NT, NJ>;
    if constexpr (SH::NN != 0) { if (SH::matches(a.T, 2)) kern = ns2_elim1_kernel<D, Q, NT, NJ, SH::NN, SH::NO>; }
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * C::WARPS, smem));
    per_sm = std::max(per_sm, 1);
    static const int waves = [] { const char* e = getenv("RBFFD_NS2_ELIM_WAVES"); const int w = e ? atoi(e) : 0; return w > 0 ? w : 64; }();
    const int grid = (int)std::min<int64_t>((a.cnt + C::WARPS - 1) / C::WARPS, (int64_t)ctx->sm_count * per_sm * waves);
    kern<<<grid, 32 * C::WARPS, smem, ctx->stream>>>(a);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

template <int D, int Q, int NT, int NJ>
int launch_elim(rbffd_context* ctx, Ns2Args& a) {
    constexpr int WS = 8 * NJ + 4, PS = 52, UST = 8 * NJ + 4;
    const size_t smem = ((size_t)(4 * PS + 2 * 4 * PS + 2 * 4 * UST + 48 + Q * WS + 8 * 8 * NT + NS2_REC_HDR / 8 + 8) * 8 + 64 * 4 + 15) & ~(size_t)15;
    static const int pad_smem = [] { const char* e = getenv("RBFFD_NS2_ELIM_PAD_SMEM"); return e ? atoi(e) : 0; }();
    const size_t smem_launch = smem + (size_t)std::max(0, pad_smem);
    auto kern = ns2_elim_kernel<D, Q, NT, NJ>;
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_launch));
    const int per_sm = std::max<int>(1, std::min<int>(8, (int)((228 * 1024) / (smem_launch + 1024))));
    static const int waves = [] { const char* e = getenv("RBFFD_NS2_ELIM_WAVES"); const int w = e ? atoi(e) : 0; return w > 0 ? w : 64; }();
    const int grid = (int)std::min<int64_t>(a.cnt, (int64_t)ctx->sm_count * per_sm * waves);
    kern<<<grid, 64, smem_launch, ctx->stream>>>(a);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

template <int D, int Q, int NT, int NJ>
int launch_solve(rbffd_context* ctx, Ns2Args& a) {
    using C = SvCfg<D, Q, NT, NJ>;
    using SH = Ns2Shape<D, Q, NT, NJ>;
    using CS = SvCfg<D, Q, NT, NJ, SH::NN, SH::NO>;
    const bool split = a.stile != nullptr && NT <= 5;
    auto kern = split ? ns2_solve_kernel<D, Q, NT, NJ, (NT <= 5)> : ns2_solve_kernel<D, Q, NT, NJ, false>;
    int gdoubles = C::G, btdoubles = 64 * a.bs, max_ctas = 4;
    if constexpr (SH::NN != 0) {
        if (split && SH::matches(a.T, 1)) {
            kern = ns2_solve_kernel<D, Q, NT, NJ, true, SH::NN, SH::NO>;
            gdoubles = CS::G; max_ctas = CS::MINB;
            btdoubles = CS::BTG ? 0 : 64 * CS::BSC;
        }
    }
    const size_t smem = ((size_t)(gdoubles + C::WT + C::SC + btdoubles + 8 + C::HDR + C::XN) * 8 + 64 * 4 + 15) & ~(size_t)15;
    static const int pad_smem = [] { const char* e = getenv("RBFFD_NSW_PAD_SMEM"); return e ? atoi(e) : 0; }();
    const size_t smem_launch = smem + (size_t)std::max(0, pad_smem);
    if ((int64_t)smem_launch > ctx->max_smem_optin) return RBFFD_ERR_UNSUPPORTED;
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_launch));
    const int per_sm = std::max<int>(1, std::min<int>(max_ctas, (int)((228 * 1024) / (smem_launch + 1024))));
    // CTAs per resident slot: with chunks of ~600 k stencils 64 (a CTA walks ~3-12 stencils, prefetching the next record) beats 256 by 0.7 %
    static const int waves = [] { const char* e = getenv("RBFFD_NSW_WAVES"); const int w = e ? atoi(e) : 0; return w > 0 ? w : 64; }();
    const int grid = (int)std::min<int64_t>(a.cnt, (int64_t)ctx->sm_count * per_sm * waves);
#ifdef NS2_TIMING
    unsigned long long zero[16] = {};
    cudaMemcpyToSymbolAsync(ns2_prof, zero, sizeof(zero), 0, cudaMemcpyHostToDevice, ctx->stream);
#endif
    kern<<<grid, 128, smem_launch, ctx->stream>>>(a);
    KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
#ifdef NS2_TIMING
    unsigned long long h[16];
    cudaMemcpyFromSymbol(h, ns2_prof, sizeof(h));
    static const char* nm[8] = {"0c W' + next header issue", "A Phi~ (4 warps)", "B Y", "C S", "D elimination", "E back+store", "0a wait for the record", "0b nodes + rhs"};
    for (int k = 0; k < 8; ++k) fprintf(stderr, "[ns2 timing] %-26s %9.0f cycles/stencil\n", nm[k], (double)h[k] / (double)a.cnt);
#endif
    if constexpr (NT <= 5) {
        // RBFFD_NS2_ELIM=2 keeps the two-warp elimination with its scalar panel chain (A/B comparisons)
        static const int elim_env = [] { const char* e = getenv("RBFFD_NS2_ELIM"); return e ? atoi(e) : 1; }();
        if (split) return elim_env == 2 ? launch_elim<D, Q, NT, NJ>(ctx, a) : launch_elim1<D, Q, NT, NJ>(ctx, a);
    }
    return RBFFD_OK;
}

}  // namespace

#define NS2_CAT2(a, b, c) a##b##_##c
#define NS2_CAT(a, b, c) NS2_CAT2(a, b, c)
// Runs stage A and stage B over rows [a.row0, a.row0 + a.cnt).  UNSUPPORTED: shape outside the instantiated tile counts.
int NS2_CAT(rbffd_ns2_launch_, NS2_D, NS2_Q)(rbffd_context* ctx, Ns2Args& a) {
    constexpr int D = NS2_D, Q = NS2_Q;
    if (!nsp::table_matches<D, Q>(a.T)) return RBFFD_ERR_UNSUPPORTED;
    const int nb = a.T.n - Q, nops = a.T.nops;
    const int nt = std::max(4, (nb + 7) / 8);
    if (nt > 6) return RBFFD_ERR_UNSUPPORTED;
    const bool fold = ((nb + 3) & ~3) + nops <= 8 * nt;
    a.bs = std::max(6, (nops + 1) & ~1);                 // Ys [nops][NBP] aliases Bt [64][bs]
    const int njt = fold ? nt : nt + 1;
    a.ws = 8 * njt + 4;
    a.rcb = fold ? ((nb + 3) & ~3) : 8 * nt;
    if (NS2_REC_W + (int64_t)Q * a.ws * 8 > a.rec_stride) return RBFFD_ERR_INVALID;
    {
        // stage A does not prefetch: a finer grid (1-2 stencils per warp) balances its tail better (config 4 shape: -2 % of the whole weight phase)
        static const int pwaves = [] { const char* e = getenv("RBFFD_NS2_PRED_WAVES"); const int w = e ? atoi(e) : 0; return w > 0 ? w : 128; }();
        const int grid = (int)std::min<int64_t>((a.cnt + 3) / 4, (int64_t)ctx->sm_count * 4 * pwaves);
        auto pk = ns2_pred_kernel<D, Q>;
        using SH3 = Ns2Shape<D, Q, 5, 5>;
        using SH4 = Ns2Shape<D, Q, 5, 6>;
        if constexpr (SH3::NN != 0) { if (SH3::matches(a.T, 4)) pk = ns2_pred_kernel<D, Q, SH3::NN, SH3::NO>; }
        if constexpr (SH4::NN != 0) { if (SH4::matches(a.T, 4)) pk = ns2_pred_kernel<D, Q, SH4::NN, SH4::NO>; }
        pk<<<grid, 128, 0, ctx->stream>>>(a);
        KLAUNCH(ctx);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    switch (nt * 2 + (fold ? 0 : 1)) {
        case 8: return launch_solve<D, Q, 4, 4>(ctx, a);
        case 9: return launch_solve<D, Q, 4, 5>(ctx, a);
        case 10: return launch_solve<D, Q, 5, 5>(ctx, a);
        case 11: return launch_solve<D, Q, 5, 6>(ctx, a);
        case 12: return launch_solve<D, Q, 6, 6>(ctx, a);
        default: return launch_solve<D, Q, 6, 7>(ctx, a);
    }
}

#else  // dispatcher

int rbffd_ns2_launch_2_10(rbffd_context* ctx, Ns2Args& a);
int rbffd_ns2_launch_2_15(rbffd_context* ctx, Ns2Args& a);
int rbffd_ns2_launch_2_21(rbffd_context* ctx, Ns2Args& a);
int rbffd_ns2_launch_3_10(rbffd_context* ctx, Ns2Args& a);
int rbffd_ns2_launch_3_20(rbffd_context* ctx, Ns2Args& a);

// Two-stage null-space path.  Returns RBFFD_ERR_UNSUPPORTED when the configuration is outside its scope, or when any
// stencil failed its definiteness / rank / finiteness check (the caller then runs the pivoted kernels over the batch).
int rbffd_weights_ns2(rbffd_context* ctx, const OpTables& T, const double* X, int64_t NS, const double* Y, int64_t M,
                      const int32_t* stencils, const int32_t* center, int32_t* colind_out, double* vals_out) {
    static const int enabled = [] { const char* e = getenv("RBFFD_NS2"); return e ? atoi(e) : 1; }();
    if (!enabled) return RBFFD_ERR_UNSUPPORTED;
    const int nb = T.n - T.q;
    if (T.nops > 8 || T.n > 64 || T.n < 8 || T.n + T.nops > 64 || nb < 1 || nb > 48 || T.dim < 2 || NS != M) return RBFFD_ERR_UNSUPPORTED;
    {
        int need = 1;                                   // conditional definiteness needs polynomial degree >= (p-1)/2
        const int deg = (T.p - 1) / 2;
        for (int tq = 1; tq <= T.dim; ++tq) need = need * (deg + tq) / tq;
        if (T.q < need) return RBFFD_ERR_UNSUPPORTED;
    }
    int (*launch)(rbffd_context*, Ns2Args&) = nullptr;
    if (T.dim == 2 && T.q == 10) launch = rbffd_ns2_launch_2_10;
    else if (T.dim == 2 && T.q == 15) launch = rbffd_ns2_launch_2_15;
    else if (T.dim == 2 && T.q == 21) launch = rbffd_ns2_launch_2_21;
    else if (T.dim == 3 && T.q == 10) launch = rbffd_ns2_launch_3_10;
    else if (T.dim == 3 && T.q == 20) launch = rbffd_ns2_launch_3_20;
    if (!launch) return RBFFD_ERR_UNSUPPORTED;
    Ns2Args a;
    a.X = X; a.Y = Y; a.stencils = stencils; a.center = center; a.M = M;
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
    for (int o = 0; o < 8; ++o) {
        a.hv_axis[o] = -1; a.hv_half[o] = 0; a.hv_rexp[o] = 1;
        for (int k = 0; k < 4; ++k) a.hv_c[o][k] = 0.0;
        if (o >= T.nops || T.kind[o] != RBFFD_OP_DERIV) continue;
        int ax = -1, K = 0, nz = 0;
        for (int b = 0; b < T.dim; ++b) if (T.alpha[o][b] > 0) { ax = b; K = T.alpha[o][b]; ++nz; }
        if (nz != 1 || K < 4 || (K & 1) || K > 6 || K >= T.p) continue;
        bool ok = true;
        double cf[4] = {0, 0, 0, 0};
        for (int t = T.tb[3 * o]; t < T.tb[3 * o + 1] && ok; ++t) {
            const int ea = T.te[t][ax];
            for (int b = 0; b < T.dim; ++b) if (b != ax && T.te[t][b] != 0) ok = false;
            if (ea < 0 || ea > K || (ea & 1) || T.te[t][3] != T.p - K - ea) ok = false;     // x^ea r^(p-K-ea) = r^(p-K) (x/r)^ea
            if (ok) cf[ea / 2] += T.coef[t];
        }
        if (!ok) continue;
        a.hv_axis[o] = (int8_t)ax; a.hv_half[o] = (int8_t)(K / 2); a.hv_rexp[o] = (int8_t)(T.p - K);
        for (int k = 0; k < 4; ++k) a.hv_c[o][k] = cf[k];
    }
    DevBuf<int> redo;
    const bool deferred = ctx->deferred_flags != nullptr;
    if (deferred) a.redo = ctx->deferred_flags + 8 * ctx->deferred_slot + 4;
    else {
        CUDA_TRY(ctx, redo.alloc(1, ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(redo.p, 0, sizeof(int), ctx->stream));
        a.redo = redo.p;
    }
    {
        const int nt = std::max(4, (nb + 7) / 8);       // the tile counts rbffd_ns2_launch_* will pick
        const bool fold = ((nb + 3) & ~3) + T.nops <= 8 * nt;
        const int ws = 8 * (fold ? nt : nt + 1) + 4;
        a.rec_stride = (NS2_REC_W + (int64_t)T.q * ws * 8 + 127) & ~(int64_t)127;
    }
    if ((reinterpret_cast<uintptr_t>(X) & 15) != 0) return RBFFD_ERR_UNSUPPORTED;      // 16-byte cp.async of 2-D coordinates
    // split path (elimination in its own kernel): [S | t] travels as NT x NJ accumulator tiles of 512 B
    static const int split_env = [] { const char* e = getenv("RBFFD_NS2_SPLIT"); return e ? atoi(e) : 1; }();
    {
        const int nt = std::max(4, (nb + 7) / 8);
        const bool fold = ((nb + 3) & ~3) + T.nops <= 8 * nt;
        a.stile_stride = (split_env && nt <= 5) ? (int64_t)nt * (fold ? nt : nt + 1) * 64 : 0;
    }
    // the scratch buffers are reused chunk by chunk.  Every chunk costs three launches with their tails (a solve CTA lives ~9 us,
    // an elimination warp ~40 us), so the chunks are made as large as a budget of min(RBFFD_NS2_SCRATCH_MB = 12 GiB, an eighth of
    // the device memory) allows: 1.5 GiB -> 12 GiB is worth 2-6 % at the BASELINE shapes (profiles/r02bi_*)
    static const int64_t chunk_env = [] { const char* e = getenv("RBFFD_NS2_CHUNK"); return e ? atoll(e) : 0ll; }();
    static const int64_t scratch_mb = [] { const char* e = getenv("RBFFD_NS2_SCRATCH_MB"); const long long v = e ? atoll(e) : 0ll; return v > 0 ? v : 12288ll; }();
    const int64_t per_item = a.rec_stride + a.stile_stride * 8;
    // (the device's TOTAL memory is read once per process: cudaMemGetInfo per call costs milliseconds next to a populated pool)
    static const int64_t total_mem = [] { size_t f = 0, t = 0; return cudaMemGetInfo(&f, &t) == cudaSuccess ? (int64_t)t : (int64_t)0; }();
    int64_t budget = scratch_mb << 20;
    if (total_mem > 0) budget = std::min<int64_t>(budget, total_mem / 8);
    int64_t chunk = std::min<int64_t>(M, chunk_env > 0 ? chunk_env : std::max<int64_t>(4096, budget / per_item));
    DevBuf<double> stile;
    DevBuf<unsigned char> rec;
    a.stile = nullptr;
    for (;;) {                                          // a device short of memory gets smaller chunks, not an error
        cudaError_t e = cudaSuccess;
        if (a.stile_stride > 0) e = stile.alloc((size_t)chunk * a.stile_stride, ctx->stream);
        if (e == cudaSuccess) e = rec.alloc((size_t)chunk * a.rec_stride, ctx->stream);
        if (e == cudaSuccess) break;
        if (e != cudaErrorMemoryAllocation || chunk <= 8192) CUDA_TRY(ctx, e);
        cudaGetLastError();
        stile.reset();
        rec.reset();
        chunk = std::max<int64_t>(4096, chunk / 4);
    }
    if (a.stile_stride > 0) a.stile = stile.p;
    a.rec = rec.p;
    int rc = RBFFD_OK;
    for (int64_t r0 = 0; r0 < M && rc == RBFFD_OK; r0 += chunk) {
        a.row0 = r0;
        a.cnt = std::min<int64_t>(chunk, M - r0);
        rc = launch(ctx, a);
    }
    if (rc != RBFFD_OK || deferred) return rc;
    int h_redo = 0;
    CUDA_TRY(ctx, rbffd_fetch_flags(ctx, redo.p, 1, &h_redo));
    return h_redo ? RBFFD_ERR_UNSUPPORTED : RBFFD_OK;
}

#endif
