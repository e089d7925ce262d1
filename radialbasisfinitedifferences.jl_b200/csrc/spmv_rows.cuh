// spmv_rows.cuh -- the row kernel body shared by the plain SpMV (spmv.cu) and the sharded SpMV with its fused halo
// exchange (shard.cu).  Replaces the SparseMatrixCSC products of examples/adv_diff_test.jl:151-152.
#pragma once

#include <cstdint>

constexpr int SPMV_MAXMAT = 6;
// up to six value planes over ONE shared pattern and their coefficients: y = sum_k c[k] * v[k] * x
struct SpmvMats { const double* v[SPMV_MAXMAT]; double c[SPMV_MAXMAT]; };
// what happens to the row sum `acc`:  su == nullptr:  y[row] = acc + beta * y[row]
//                                     su != nullptr:  y[row] = sa * su[row] + sb * (x[row] + sdt * acc)   -- one SSP-RK stage
//                                                     u_new = a u + b (v + dt L(v)) fused into the product (rows = nodes)
struct SpmvEpilogue { double beta; const double* su; double sa, sb, sdt; };

// TPR lanes cooperate on one row; a warp covers 32/TPR consecutive rows = one contiguous chunk of HBM.
// Every lane first issues ALL of its value/index loads (streaming, evict-first), then the gathers of x
// (read-only path, kept in L1/L2), then the FMAs: ITERS*VEC independent loads in flight per lane.
// VEC = 2 uses 16-byte value loads and 8-byte index loads (needs an even row length: rows stay 16-byte aligned).
// HALO: column ids >= n_owned read `xh[col - n_owned]` (the rank's halo inbox) instead of x[col].
// `warp` = index of this warp among the warps working on rows [row_begin, row_end).
template <int TPR, int NMAT, int VEC, int ITERS, bool HALO>
__device__ __forceinline__ void spmv_rows(int64_t warp, int64_t row_begin, int64_t row_end, int n, const int32_t* __restrict__ colind,
                                          const SpmvMats& m, const double* __restrict__ x, const double* __restrict__ xh, int n_owned,
                                          const SpmvEpilogue& ep, double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int t = lane % TPR;
    const int64_t row = row_begin + warp * (32 / TPR) + lane / TPR;
    double acc = 0.0;
    // the epilogue operands of the row are requested together with the matrix stream (lane 0 of the team), not after the
    // reduction: a DRAM round trip at the end of every warp's life would halve the bytes in flight per SM
    double e_u = 0.0, e_x = 0.0;
    if (row < row_end && t == 0) {
        if (ep.su) { e_u = __ldcs(ep.su + row); e_x = __ldg(x + row); }
        else if (ep.beta != 0.0) e_u = __ldcs(y + row);
    }
    if (row < row_end) {
        const int64_t base = row * n;
        int col[ITERS][VEC];
        double w[ITERS][VEC];
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int j = (t + it * TPR) * VEC;
            if (VEC == 2) {
                if (j < n) {
                    const int2 ci = __ldcs(reinterpret_cast<const int2*>(colind + base + j));
                    double2 a0 = __ldcs(reinterpret_cast<const double2*>(m.v[0] + base + j));
                    a0.x *= m.c[0]; a0.y *= m.c[0];
#pragma unroll
                    for (int k = 1; k < NMAT; ++k) {
                        const double2 b = __ldcs(reinterpret_cast<const double2*>(m.v[k] + base + j));
                        a0.x += m.c[k] * b.x; a0.y += m.c[k] * b.y;
                    }
                    col[it][0] = ci.x; col[it][VEC - 1] = ci.y;
                    w[it][0] = a0.x; w[it][VEC - 1] = a0.y;
                } else {
                    col[it][0] = col[it][VEC - 1] = -1;
                    w[it][0] = w[it][VEC - 1] = 0.0;
                }
            } else {
                if (j < n) {
                    col[it][0] = __ldcs(colind + base + j);
                    double a0 = m.c[0] * __ldcs(m.v[0] + base + j);
#pragma unroll
                    for (int k = 1; k < NMAT; ++k) a0 += m.c[k] * __ldcs(m.v[k] + base + j);
                    w[it][0] = a0;
                } else {
                    col[it][0] = -1;
                    w[it][0] = 0.0;
                }
            }
        }
        double xv[ITERS][VEC];
#pragma unroll
        for (int it = 0; it < ITERS; ++it)
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const int cj = col[it][e];
                if (HALO) {
                    // the inbox is written by peer GPUs during this kernel: never through the non-coherent read-only path
                    xv[it][e] = cj < 0 ? 0.0 : (cj < n_owned ? __ldg(x + cj) : __ldcg(xh + (cj - n_owned)));
                } else {
                    xv[it][e] = cj >= 0 ? __ldg(x + cj) : 0.0;
                }
            }
#pragma unroll
        for (int it = 0; it < ITERS; ++it)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc = fma(w[it][e], xv[it][e], acc);
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (row < row_end && t == 0) {
        if (ep.su) y[row] = ep.sa * e_u + ep.sb * (e_x + ep.sdt * acc);
        else y[row] = ep.beta == 0.0 ? acc : acc + ep.beta * e_u;
    }
}
