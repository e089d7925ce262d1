// nullspace.cuh -- device helpers shared by the two-stage null-space weight path (weights_ns2.cu): FP64 tensor-core
// MMA wrapper, fast reciprocal, warp maximum, named barriers, cp.async, and the graded monomial table in closed form
// (the same table build_op_tables in weights.cu produces; checked on the host by ns2_table_matches).
#pragma once

#include <utility>

#include "common.cuh"
#include "tables.cuh"

namespace nsp {

// D (8x8) += A (8x4, row) * B (4x8, col): lane (g = lane/4, t = lane%4) holds A[g][t], B[t][g], C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
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
__device__ __forceinline__ double warp_max_nn(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned hmax = __reduce_max_sync(0xffffffffu, hi);
    const unsigned lmax = __reduce_max_sync(0xffffffffu, hi == hmax ? lo : 0u);
    return __hiloint2double((int)hmax, (int)lmax);
}

__device__ __forceinline__ void bar_named(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// pivot block of step kb (held by the lanes g / 4 == h, t / 2 == h of tile (kb/2, kb/2) as c0, c1) -> shared memory -> every
// lane: LDL' (upper triangle; S is symmetric to rounding), then the column g % 4 of D^-1 (= row g % 4: symmetric); returns
// element t of that column and ORs the sign violations of the four pivots into `bad`
__device__ __forceinline__ double block_gj_factor(double c0, double c1, const int kb, double* __restrict__ dsm, const int sgnbits, int& bad) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int h = kb & 1;
    double* Db = dsm + (kb & 1) * 16;
    if ((g >> 2) == h && (t >> 1) == h) *reinterpret_cast<double2*>(Db + (g & 3) * 4 + 2 * (t & 1)) = make_double2(c0, c1);
    __syncwarp();
    const double2 a0 = *reinterpret_cast<const double2*>(Db), a0b = *reinterpret_cast<const double2*>(Db + 2);
    const double a11 = Db[5];
    const double2 a1b = *reinterpret_cast<const double2*>(Db + 6), a2b = *reinterpret_cast<const double2*>(Db + 10);
    const double a33 = Db[15];
    const double a00 = a0.x, a01 = a0.y, a02 = a0b.x, a03 = a0b.y, a12 = a1b.x, a13 = a1b.y, a22 = a2b.x, a23 = a2b.y;
    const double r0 = rcp3(a00);
    const double l10 = a01 * r0, l20 = a02 * r0, l30 = a03 * r0;
    const double d1 = fma(-l10, a01, a11);
    const double r1 = rcp3(d1);
    const double v21 = fma(-l20, a01, a12), v31 = fma(-l30, a01, a13);
    const double l21 = v21 * r1, l31 = v31 * r1;
    const double d2 = fma(-l21, v21, fma(-l20, a02, a22));
    const double r2 = rcp3(d2);
    const double v32 = fma(-l31, v21, fma(-l30, a02, a23));
    const double l32 = v32 * r2;
    const double d3 = fma(-l32, v32, fma(-l31, v31, fma(-l30, a03, a33)));
    const double r3 = rcp3(d3);
    bad |= (__double2hiint(a00) ^ sgnbits) | (__double2hiint(d1) ^ sgnbits) | (__double2hiint(d2) ^ sgnbits) | (__double2hiint(d3) ^ sgnbits);
    const int ci = g & 3;
    const double z0 = ci == 0 ? 1.0 : 0.0;
    const double z1 = fma(-l10, z0, ci == 1 ? 1.0 : 0.0);
    const double z2 = fma(-l21, z1, fma(-l20, z0, ci == 2 ? 1.0 : 0.0));
    const double z3 = fma(-l32, z2, fma(-l31, z1, fma(-l30, z0, ci == 3 ? 1.0 : 0.0)));
    const double x3 = z3 * r3;
    const double x2 = fma(-l32, x3, z2 * r2);
    const double x1 = fma(-l31, x3, fma(-l21, x2, z1 * r1));
    const double x0 = fma(-l30, x3, fma(-l20, x2, fma(-l10, x1, z0 * r0)));
    return t == 0 ? x0 : (t == 1 ? x1 : (t == 2 ? x2 : x3));
}

// ---------------------------------------------------------------------------------------------------------------------
// Unpivoted Gauss-Jordan by 4 x 4 BLOCK pivots on a definite matrix held as DMMA accumulator tiles, ONE warp, no barrier:
// c[I][J][0..1] = rows 8I + g, columns 8J + 2t, 8J + 2t + 1 of [S | t] (S: (8 NT)^2, identity-padded beyond nb; t: the
// right-hand sides in the trailing columns).  Block step kb takes the pivot block D = S[K, K], K = 4kb .. 4kb+3:
//   * D travels through 128 B of shared memory to every lane, which factors it (LDL', four reciprocals) and solves for the
//     ONE column of D^-1 it needs -- no per-row panel work, no pivot search, no row exchange;
//   * per tile column J to the right:  R = -D^-1 X[K, J]  (one DMMA, D^-1 sitting in rows 4h .. 4h+3 of the A operand),
//     X[:, J] += S[:, K] R  (NT DMMAs), X[K, J] = -R  (the pivot rows come out already scaled, so the result is [I | y]);
//   * operands change fragment layout (accumulator -> A / B) by SHFL only.
// Returns the OR of (sign bits of the 16 pivots of the LDL' factorisations) ^ sgnbits: negative when S is not definite.
// dsm: 32 doubles of shared memory private to the warp.
// The factorisation of the NEXT pivot block is software-pipelined: that block lives in the first tile column a step updates,
// so it is extracted and factored right after that column, and its dependent scalar chain (LDL', four reciprocals, column
// solve: ~500 clocks) overlaps the DMMAs of the remaining tile columns instead of sitting between two steps.
// PIPE = false factors every pivot block at the top of its own step (fewer live registers: the 5 x 6-tile instance of
// ns2_elim1_kernel spills with the pipelined form and is 5 % slower, the 5 x 5 one is 5.5 % faster).
// (CI, CJ: declared extents of the accumulator array, >= NT, NJ)
template <int NT, int NJ, int CI, int CJ, bool PIPE>
__device__ __forceinline__ int block_gj_warp(double (&c)[CI][CJ][2], const int nb, double* __restrict__ dsm, const int sgnbits) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    int bad = 0;
    double xt = 0.0;
    if constexpr (PIPE) xt = block_gj_factor(c[0][0][0], c[0][0][1], 0, dsm, sgnbits, bad);
#pragma unroll
    for (int kb = 0; kb < 2 * NT; ++kb) {
        if (4 * kb < nb) {                                    // warp-uniform: identity-padded block steps are skipped
            const int Jp = kb >> 1, h = kb & 1, jlo = h == 0 ? Jp : Jp + 1;
            const int kn = kb + 1, Jn = kn >> 1;              // next pivot block: tile (Jn, Jn); Jn == jlo
            const bool next = kn < 2 * NT && 4 * kn < nb;
            if constexpr (!PIPE) xt = block_gj_factor(c[Jp][Jp][0], c[Jp][Jp][1], kb, dsm, sgnbits, bad);
            const double aD = (g >> 2) == h ? -xt : 0.0;      // A operand: -D^-1 in rows 4h .. 4h+3
            // panel columns K of every row tile as A operands: lane (g, t) <- column 4h + t of tile (I, Jp)
            double aP[NT];
            {
                const int src = (g << 2) | (2 * h + (t >> 1));
#pragma unroll
                for (int I = 0; I < NT; ++I) {
                    const double v0 = __shfl_sync(FULL, c[I][Jp][0], src), v1 = __shfl_sync(FULL, c[I][Jp][1], src);
                    aP[I] = (t & 1) ? v1 : v0;
                }
            }
            const int srcK = ((4 * h + t) << 2) | (g >> 1);   // B operand of a 4-row block living in rows 4h .. 4h+3 of a tile
#pragma unroll
            for (int J = 0; J < NJ; ++J) {
                if (J >= jlo) {
                    const double k0 = __shfl_sync(FULL, c[Jp][J][0], srcK), k1 = __shfl_sync(FULL, c[Jp][J][1], srcK);
                    double q0 = 0.0, q1 = 0.0;
                    dmma884(q0, q1, aD, (g & 1) ? k1 : k0);   // R = -D^-1 X[K, J] in rows 4h .. 4h+3
                    const double b0 = __shfl_sync(FULL, q0, srcK), b1 = __shfl_sync(FULL, q1, srcK);
                    const double bR = (g & 1) ? b1 : b0;
#pragma unroll
                    for (int I = 0; I < NT; ++I) dmma884(c[I][J][0], c[I][J][1], aP[I], bR);
                    if ((g >> 2) == h) { c[Jp][J][0] = -q0; c[Jp][J][1] = -q1; }
                    if (PIPE && J == jlo && next && Jn < NT && Jn < NJ) xt = block_gj_factor(c[Jn][Jn][0], c[Jn][Jn][1], kn, dsm, sgnbits, bad);
                }
            }
        }
    }
    return bad;
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

template <int D, int Q>
inline bool table_matches(const OpTables& T) {
    for (int c = 1; c < Q; ++c)
        if (T.mpar[c] != mono_par(D, c) || T.maxis[c] != mono_axis(D, c)) return false;
    return true;
}

}  // namespace nsp
