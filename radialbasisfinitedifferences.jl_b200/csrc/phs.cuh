// phs.cuh -- PHS kernel evaluation shared by the null-space weight kernels (weights_ns.cu, weights_nsw.cu):
// Phi_ij = ||S_i - S_j||^p (src/rbfblock.jl:14-20, src/rbfbasis.jl:9) assembled once per unordered pair.
#pragma once

// 1/sqrt(x), x > 0, to ~1 ulp: MUFU.RSQ64H seed (relative error ~2^-20) + one third-order step
// y (1 + e/2 + 3 e^2/8), e = 1 - x y^2
__device__ __forceinline__ double phs_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    double u = fma(e, 0.375, 0.5);
    u = u * e;
    return fma(y, u, y);
}

// r^p = r2^((p+1)/2) / r for odd p = 2 HP + 1, y = 1/r.  HP is a template parameter so that the pair loops below
// contain no branch (HP < 0: run-time power hp, generic loop).
template <int HP>
__device__ __forceinline__ double phs_pow_t(double r2, double y, int hp) {
    if constexpr (HP == 0) return r2 * y;
    else if constexpr (HP == 1) return (r2 * r2) * y;
    else if constexpr (HP == 2) { const double r4 = r2 * r2; return (r4 * r2) * y; }
    else if constexpr (HP == 3) { const double r4 = r2 * r2; return (r4 * r4) * y; }
    else if constexpr (HP == 4) { const double r4 = r2 * r2; return (r4 * r4) * (r2 * y); }
    else {
        double v = y;
        for (int e = 0; e <= hp; ++e) v *= r2;
        return v;
    }
}

// Symmetric assembly by circulant pairing: the thread that owns point l (coordinates `me`, l < n) pairs in round k with
// point (l + k) mod n; rounds 1 .. n/2 visit every unordered pair (the last round of an even n visits its pairs twice,
// with identical values).  Sc = coordinates [point][DP] in shared memory, G = n x n tile with row stride LD.  Threads
// with active == false run along with l = 0 and store nothing.  Two rounds per trip: two independent dependency chains.
// Coincident nodes give r2 = 0 -> NaN, which the finiteness check of the weights turns into the pivoted fallback.
template <int D, int DP, int LD, int HP>
__device__ __forceinline__ void phs_assemble_t(const double* __restrict__ Sc, double* __restrict__ G, const double* me, int l, int n,
                                               bool active, int hp) {
    double* const grow_l = G + l * LD;
    double* const gcol_l = G + l;
    const int rounds = n >> 1;
    auto phi = [&](int k, int& ib) -> double {
        ib = l + k;
        ib = ib >= n ? ib - n : ib;
        double o[D];
        const double2 v = *reinterpret_cast<const double2*>(Sc + ib * DP);
        o[0] = v.x; o[1] = v.y;
        if constexpr (D == 3) o[2] = Sc[ib * DP + 2];
        double r2 = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) { const double dd = me[c] - o[c]; r2 = fma(dd, dd, r2); }
        return phs_pow_t<HP>(r2, phs_rsqrt(r2), hp);
    };
    int k = 1;
    for (; k + 1 <= rounds; k += 2) {
        int b0, b1;
        const double v0 = phi(k, b0);
        const double v1 = phi(k + 1, b1);
        if (active) { grow_l[b0] = v0; gcol_l[b0 * LD] = v0; grow_l[b1] = v1; gcol_l[b1 * LD] = v1; }
    }
    if (k <= rounds) {
        int b0;
        const double v0 = phi(k, b0);
        if (active) { grow_l[b0] = v0; gcol_l[b0 * LD] = v0; }
    }
}

template <int D, int DP, int LD>
__device__ __forceinline__ void phs_assemble(const double* Sc, double* G, const double* me, int l, int n, bool active, int hp) {
    switch (hp) {                                       // warp-uniform
        case 1: phs_assemble_t<D, DP, LD, 1>(Sc, G, me, l, n, active, hp); break;     // r^3
        case 2: phs_assemble_t<D, DP, LD, 2>(Sc, G, me, l, n, active, hp); break;     // r^5
        case 3: phs_assemble_t<D, DP, LD, 3>(Sc, G, me, l, n, active, hp); break;     // r^7
        default: phs_assemble_t<D, DP, LD, -1>(Sc, G, me, l, n, active, hp); break;
    }
}
