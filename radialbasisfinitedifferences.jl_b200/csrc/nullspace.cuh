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
