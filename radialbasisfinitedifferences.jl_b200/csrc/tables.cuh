// tables.cuh -- operator descriptor tables passed BY VALUE to the weight kernels (no global state).
//
// The reference derives the PHS derivatives with Symbolics at run time (src/rbfbasis.jl:20-30,
// src/rbfbasis_k.jl:9-18) and the monomial derivative systems with DynamicPolynomials
// (src/polynomialbasis.jl:14-30, src/polynomialbasis_k.jl:14-19).  Here the host builds the closed
// forms once per call as a term list  sum_t coef_t * prod_a x_a^e_ta * r^rpow_t  using
// d/dx_a (x^e r^q) = e_a x^(e-1_a) r^q + q x^(e+1_a) r^(q-2), and the device evaluates them.
#pragma once

#include <cstdint>

#include "../../include/rbffd.h"

constexpr int TAB_MAX_TERMS = 96;
constexpr int TAB_MAX_MONO = 84;    // C(6+3,3): degree <= 6 in 3-D, degree <= 11 in 2-D

struct OpTables {
    int32_t dim, p, n, q, m, nops;
    int32_t kind[RBFFD_MAX_OPS];
    int8_t alpha[RBFFD_MAX_OPS][4];
    // DERIV: terms [tb[3*o], tb[3*o+1]) ; LAPLACE: axis a terms [tb[3*o+a], tb[3*o+a+1])
    int16_t tb[3 * RBFFD_MAX_OPS + 1];
    double coef[TAB_MAX_TERMS];
    int8_t te[TAB_MAX_TERMS][4];    // e0,e1,e2, rpow
    int8_t mono[TAB_MAX_MONO][4];   // exponent table, graded order
    int8_t mpar[TAB_MAX_MONO];      // mono[t] = mono[mpar[t]] * x[maxis[t]]  (t > 0; parents precede children)
    int8_t maxis[TAB_MAX_MONO];
};

int build_op_tables(const rbffd_options* o, OpTables* T, char* err, int errlen);

#ifdef __CUDACC__
// 1/x to ~1 ulp without the IEEE slow path: MUFU.RCP64H seed (about 20 bits) + two Newton steps
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// sqrt(x) for x > 0 (x == 0 -> 0) to ~1 ulp: MUFU.RSQ64H seed + Newton on 1/sqrt + one correction of the root
__device__ __forceinline__ double fast_sqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double h = 0.5 * y;
    double e = fma(-x * y, h, 0.5);       // 0.5 - x*y*y/2
    y = fma(y, e, y);
    h = 0.5 * y;
    e = fma(-x * y, h, 0.5);
    y = fma(y, e, y);
    double r = x * y;
    double d = fma(-r, r, x);
    r = fma(d, 0.5 * y, r);
    return x > 0.0 ? r : 0.0;
}

__device__ __forceinline__ double ipow_u(double x, int e) {
    double r = 1.0;
    for (int t = 0; t < e; ++t) r *= x;
    return r;
}

// r^e for any integer e, given r and r*r
__device__ __forceinline__ double rpow_i(double r, double r2, int e) {
    int ae = e < 0 ? -e : e;
    double v = (ae & 1) ? r : 1.0;
    for (int t = 0; t < (ae >> 1); ++t) v *= r2;
    return e < 0 ? 1.0 / v : v;
}

template <int D>
__device__ __forceinline__ double eval_rbf_terms(const OpTables& T, int b, int e, const double* del, double r, double r2) {
    double s = 0.0;
    for (int t = b; t < e; ++t) {
        double v = T.coef[t];
#pragma unroll
        for (int a = 0; a < D; ++a) v *= ipow_u(del[a], T.te[t][a]);
        v *= rpow_i(r, r2, T.te[t][3]);
        s += v;
    }
    return s;
}

// d^alpha of the monomial with exponents ex at the point x
template <int D>
__device__ __forceinline__ double eval_mono_deriv(const int8_t* ex, const int8_t* alpha, const double* x) {
    double v = 1.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        int e = ex[a], al = alpha[a];
        if (e < al) return 0.0;
        for (int t = 0; t < al; ++t) v *= (double)(e - t);
        v *= ipow_u(x[a], e - al);
    }
    return v;
}

// right-hand-side entry of operator o for the RBF centred at offset del (rows 0..n-1 of the RHS)
template <int D>
__device__ __forceinline__ double rhs_rbf_entry(const OpTables& T, int o, const double* del, const double* s) {
    double r2 = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) r2 += del[a] * del[a];
    double r = sqrt(r2);
    if (T.kind[o] == RBFFD_OP_DERIV) return eval_rbf_terms<D>(T, T.tb[3 * o], T.tb[3 * o + 1], del, r, r2);
    double v = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) v += s[a] * s[a] * eval_rbf_terms<D>(T, T.tb[3 * o + a], T.tb[3 * o + a + 1], del, r, r2);
    return v;
}

// term-list evaluation without a division: rinv = 1/r is already known to the caller (r^e for e < 0 is rinv^|e|)
template <int D>
__device__ __forceinline__ double eval_rbf_terms_rinv(const OpTables& T, int b, int e, const double* del, double r, double r2, double rinv) {
    const double ri2 = rinv * rinv;
    double s = 0.0;
    for (int t = b; t < e; ++t) {
        double v = T.coef[t];
#pragma unroll
        for (int a = 0; a < D; ++a) v *= ipow_u(del[a], T.te[t][a]);
        const int ex = T.te[t][3], ae = ex < 0 ? -ex : ex;
        const double base1 = ex < 0 ? rinv : r, base2 = ex < 0 ? ri2 : r2;
        double pw = (ae & 1) ? base1 : 1.0;
        for (int u = 0; u < (ae >> 1); ++u) pw *= base2;
        s = fma(v, pw, s);
    }
    return s;
}

// closed forms for operators of total order <= 2 and the Laplacian (the reference's standard tuple); higher orders
// (hyperviscosity) go through the term tables.  rp2 = r^(p-2), rp4 = r^(p-4), rp = r^p for this offset.
template <int D>
__device__ __forceinline__ double rhs_rbf_entry_fast(const OpTables& T, int o, const double* del, const double* s,
                                                     double r, double r2, double rp, double rp2, double rp4) {
    const double pp = (double)T.p, pq = (double)(T.p * (T.p - 2));
    if (T.kind[o] == RBFFD_OP_LAPLACE) {
        double v = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) v += s[a] * s[a] * (pp * rp2 + pq * del[a] * del[a] * rp4);
        return v;
    }
    int order = 0, a0 = -1, a1 = -1;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        const int al = T.alpha[o][a];
        order += al;
        if (al >= 1) { if (a0 < 0) a0 = a; else a1 = a; }
        if (al >= 2) a1 = a;
    }
    if (order == 0) return rp;
    if (order > 2) return eval_rbf_terms<D>(T, T.tb[3 * o], T.tb[3 * o + 1], del, r, r2);
    double d0 = 0.0, d1 = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) { if (a == a0) d0 = del[a]; if (a == a1) d1 = del[a]; }
    if (order == 1) return pp * d0 * rp2;
    if (a0 == a1) return pp * rp2 + pq * d0 * d0 * rp4;
    return pq * d0 * d1 * rp4;
}

// polynomial right-hand-side rows at eta == 0 exactly (collocated rows): d^alpha x^e (0) = alpha! [e == alpha]
template <int D>
__device__ __forceinline__ double rhs_poly_entry_at_zero(const OpTables& T, int o, int t, const double* s) {
    if (T.kind[o] == RBFFD_OP_DERIV) {
        double v = 1.0;
#pragma unroll
        for (int a = 0; a < D; ++a) {
            const int e = T.mono[t][a], al = T.alpha[o][a];
            if (e != al) return 0.0;
            for (int u = 2; u <= al; ++u) v *= (double)u;
        }
        return v;
    }
    double v = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        bool hit = true;
#pragma unroll
        for (int b = 0; b < D; ++b) hit = hit && (T.mono[t][b] == (a == b ? 2 : 0));
        if (hit) v += 2.0 * s[a] * s[a];
    }
    return v;
}

// right-hand-side entry of operator o for monomial t at the scaled evaluation point eta (rows n..m-1)
template <int D>
__device__ __forceinline__ double rhs_poly_entry(const OpTables& T, int o, int t, const double* eta, const double* s) {
    if (T.kind[o] == RBFFD_OP_DERIV) return eval_mono_deriv<D>(T.mono[t], T.alpha[o], eta);
    double v = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        int8_t al[4] = {0, 0, 0, 0};
        al[a] = 2;
        v += s[a] * s[a] * eval_mono_deriv<D>(T.mono[t], al, eta);
    }
    return v;
}

// chain-rule factor applied to the weights (generate_operator.jl:161-166, hyperviscosity_operator.jl:159-160)
template <int D>
__device__ __forceinline__ double op_post_factor(const OpTables& T, int o, const double* s) {
    if (T.kind[o] != RBFFD_OP_DERIV) return 1.0;
    double f = 1.0;
#pragma unroll
    for (int a = 0; a < D; ++a) f *= ipow_u(s[a], T.alpha[o][a]);
    return f;
}
#endif
