/*
 * rbffd_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic of comp-physics/RadialBasisFiniteDifferences.jl for the
 * hot path  kNN -> per-node RBF-FD weight solve -> fixed-row-length CSR -> SpMV / semidiscrete RHS.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product path (librbffd.so, CUDA) never does.
 *
 * PARITY PINNING: the reference is pure Julia and Julia is not installed in this image, so the
 * reference itself cannot be executed ("oracle/_ref" cannot be built: there is no C/C++ source in the
 * reference at all).  The reference holds no golden weight / index vectors; this oracle is pinned
 * against the known-answer tests the reference does hold (tests/test_oracle_reference_tests.py):
 *   - test/poisson_test.jl:132       rel. l2 error < 0.0027 on the Tominec node set
 *   - test/hyperviscosity_test.jl:32 hyperviscosity(K=2) == Dxx, Dyy
 * plus reference-free identities (polynomial reproduction, row sums).  Index-level and weight-level
 * parity is therefore "pinned by known-answer tests only" (see DESIGN.md).
 *
 * Reference lines restated here (paths relative to /root/reference):
 *   kNN                 src/generate_operator.jl:43-47, src/calculateneighbors.jl:16-42,83-94
 *                       (NearestNeighbors.jl 0.4.13 KDTree/knn, third-party, not vendored: restated as
 *                        a median-split kd-tree, leafsize 10, exact Euclidean, results ascending)
 *   scalestencil        src/scalestencil.jl:10-20
 *   rbfblock            src/rbfblock.jl:14-20,  src/rbfbasis.jl:9
 *   polynomialblock     src/polynomialblock.jl:26-31, src/polynomialbasis.jl:8-11
 *   interpolationmatrix src/interpolationmatrix.jl:5-8   (inv(A) = LAPACK getrf+getri, restated)
 *   RHS + weights       src/generate_operator.jl:89-167, src/hyperviscosity_operator.jl:97-161
 *   poly RHS            src/polylinearoperator.jl:36-44,63-67
 *   RBF derivatives     src/rbfbasis.jl:20-30, src/rbfbasis_k.jl:9-18 (Symbolics closed forms, restated
 *                       by the recurrence d/dx_a (x^e r^q) = e_a x^(e-1_a) r^q + q x^(e+1_a) r^(q-2))
 *   sparse assembly     src/generate_operator.jl:171-182 (every row n entries, zeros kept)
 *   SpMV / RHS          examples/adv_diff_test.jl:144-188
 *
 * All indices are 0-based here (the Julia API is 1-based).
 * Compile with -ffp-contract=off: the kNN distance arithmetic must use separate mul/add roundings.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_DIM 3
#define ORC_MAX_TERMS 64
#define ORC_OP_DERIV 0      /* d^alpha, alpha = (a0,a1,a2); post-factor s^alpha                     */
#define ORC_OP_LAPLACE 1    /* sum_a s_a^2 d_aa in scaled coordinates (= Dxx+Dyy[+Dzz]); no post-factor */

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's reference arm: launchers such as torch.distributed.run export OMP_NUM_THREADS=1, which would silently
 * turn the "all host cores" baseline into a single-thread run */
void orc_set_num_threads(int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * kNN
 * ---------------------------------------------------------------------------------------------- */

/* candidate visibility rule of calculateneighbors.jl:16-42,83-87.
 * kind: 0 interior, 1 boundary, 2 ghost.  bnd: boundary id for kinds 1,2.
 * A boundary/ghost query of boundary i sees interior + ALL boundary nodes + ghosts of boundary i
 * (calculateneighbors.jl:23-30); interior queries (and every Y query) see everything (:83-94). */
static inline int orc_allowed(int qkind, int qbnd, int ckind, int cbnd) {
    if (qkind == 0) return 1;
    if (ckind == 2) return cbnd == qbnd;
    return 1;
}

static inline double orc_dist2(const double* a, const double* b, int d) {
    double dx = a[0] - b[0];
    double s = dx * dx;
    for (int t = 1; t < d; ++t) {
        double dt = a[t] - b[t];
        s = s + dt * dt;
    }
    return s;
}

typedef struct { double d2; int64_t idx; } OrcCand;

static inline int orc_less(double d2a, int64_t ia, double d2b, int64_t ib) {
    return (d2a < d2b) || (d2a == d2b && ia < ib);
}

/* max-heap on (d2, idx) */
static void heap_sift_down(OrcCand* h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, b = i;
        if (l < n && orc_less(h[b].d2, h[b].idx, h[l].d2, h[l].idx)) b = l;
        if (r < n && orc_less(h[b].d2, h[b].idx, h[r].d2, h[r].idx)) b = r;
        if (b == i) return;
        OrcCand t = h[i]; h[i] = h[b]; h[b] = t;
        i = b;
    }
}
static void heap_sift_up(OrcCand* h, int i) {
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!orc_less(h[p].d2, h[p].idx, h[i].d2, h[i].idx)) return;
        OrcCand t = h[i]; h[i] = h[p]; h[p] = t;
        i = p;
    }
}
static inline void heap_offer(OrcCand* h, int* cnt, int k, double d2, int64_t idx) {
    if (*cnt < k) {
        h[*cnt].d2 = d2; h[*cnt].idx = idx;
        heap_sift_up(h, *cnt);
        (*cnt)++;
    } else if (orc_less(d2, idx, h[0].d2, h[0].idx)) {
        h[0].d2 = d2; h[0].idx = idx;
        heap_sift_down(h, k, 0);
    }
}
static void heap_sort_ascending(OrcCand* h, int n) {
    for (int e = n - 1; e > 0; --e) {
        OrcCand t = h[0]; h[0] = h[e]; h[e] = t;
        heap_sift_down(h, e, 0);
    }
}

typedef struct {
    int dim;          /* split dimension, -1 for a leaf */
    double split;
    int64_t lo, hi;   /* range in perm[] */
    int left, right;
} KdNode;

typedef struct {
    const double* X; int d; int64_t N;
    int64_t* perm;
    KdNode* nodes; int nnodes, cap;
} KdTree;

static void kd_select(const double* X, int d, int64_t* a, int64_t n, int64_t kth, int dim) {
    /* quickselect on coordinate dim (ties by index to stay deterministic) */
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        double pv = X[a[mid] * d + dim]; int64_t pi = a[mid];
        int64_t i = lo, j = hi;
        while (i <= j) {
            while (X[a[i] * d + dim] < pv || (X[a[i] * d + dim] == pv && a[i] < pi)) i++;
            while (X[a[j] * d + dim] > pv || (X[a[j] * d + dim] == pv && a[j] > pi)) j--;
            if (i <= j) { int64_t t = a[i]; a[i] = a[j]; a[j] = t; i++; j--; }
        }
        if (kth <= j) hi = j; else if (kth >= i) lo = i; else return;
    }
}

static int kd_build(KdTree* T, int64_t lo, int64_t hi, int leafsize) {
    if (T->nnodes == T->cap) {
        T->cap *= 2;
        T->nodes = (KdNode*)realloc(T->nodes, sizeof(KdNode) * (size_t)T->cap);
    }
    int me = T->nnodes++;
    T->nodes[me].lo = lo; T->nodes[me].hi = hi; T->nodes[me].dim = -1;
    T->nodes[me].left = T->nodes[me].right = -1; T->nodes[me].split = 0.0;
    if (hi - lo <= leafsize) return me;
    /* widest-spread dimension */
    int best = 0; double bw = -1.0;
    for (int t = 0; t < T->d; ++t) {
        double mn = DBL_MAX, mx = -DBL_MAX;
        for (int64_t i = lo; i < hi; ++i) {
            double v = T->X[T->perm[i] * T->d + t];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        if (mx - mn > bw) { bw = mx - mn; best = t; }
    }
    int64_t mid = lo + (hi - lo) / 2;
    kd_select(T->X, T->d, T->perm + lo, hi - lo, mid - lo, best);
    double split = T->X[T->perm[mid] * T->d + best];
    int l = kd_build(T, lo, mid, leafsize);
    int r = kd_build(T, mid, hi, leafsize);
    T->nodes[me].dim = best; T->nodes[me].split = split;
    T->nodes[me].left = l; T->nodes[me].right = r;
    return me;
}

typedef struct {
    const KdTree* T; const double* q; int k; OrcCand* heap; int cnt;
    const int32_t* xkind; const int32_t* xbnd; int qkind, qbnd;
} KdQuery;

static void kd_search(KdQuery* Q, int node) {
    const KdTree* T = Q->T;
    const KdNode* nd = &T->nodes[node];
    if (nd->dim < 0) {
        for (int64_t i = nd->lo; i < nd->hi; ++i) {
            int64_t j = T->perm[i];
            if (Q->xkind && !orc_allowed(Q->qkind, Q->qbnd, Q->xkind[j], Q->xbnd[j])) continue;
            heap_offer(Q->heap, &Q->cnt, Q->k, orc_dist2(Q->q, T->X + j * T->d, T->d), j);
        }
        return;
    }
    double diff = Q->q[nd->dim] - nd->split;
    int near = diff < 0 ? nd->left : nd->right;
    int far = diff < 0 ? nd->right : nd->left;
    kd_search(Q, near);
    /* every point on the far side has d2 >= diff*diff (monotone rounding); "<=" keeps index ties */
    if (Q->cnt < Q->k || diff * diff <= Q->heap[0].d2) kd_search(Q, far);
}

/* Exact kNN of NQ query points among the N points X, results ascending by (d2, idx).
 * xkind/xbnd/qkind/qbnd may all be NULL (no masking).  brute != 0 forces the O(N*NQ) scan.
 * Rows with fewer than k visible candidates are padded with idx -1 / d2 inf.  Returns 0. */
int orc_knn(const double* X, int64_t N, int d, const double* Q, int64_t NQ, int k,
            const int32_t* xkind, const int32_t* xbnd, const int32_t* qkind, const int32_t* qbnd,
            int brute, int64_t* idx_out, double* d2_out) {
    if (d < 1 || d > ORC_MAX_DIM || k < 1 || N < 1) return 1;
    KdTree T; memset(&T, 0, sizeof(T));
    if (!brute) {
        T.X = X; T.d = d; T.N = N;
        T.perm = (int64_t*)malloc(sizeof(int64_t) * (size_t)N);
        for (int64_t i = 0; i < N; ++i) T.perm[i] = i;
        T.cap = 1024; T.nodes = (KdNode*)malloc(sizeof(KdNode) * (size_t)T.cap);
        kd_build(&T, 0, N, 10);
    }
#pragma omp parallel
    {
        OrcCand* heap = (OrcCand*)malloc(sizeof(OrcCand) * (size_t)k);
#pragma omp for schedule(dynamic, 256)
        for (int64_t qi = 0; qi < NQ; ++qi) {
            int cnt = 0;
            int qk = qkind ? qkind[qi] : 0, qb = qbnd ? qbnd[qi] : 0;
            if (brute) {
                for (int64_t j = 0; j < N; ++j) {
                    if (xkind && !orc_allowed(qk, qb, xkind[j], xbnd[j])) continue;
                    heap_offer(heap, &cnt, k, orc_dist2(Q + qi * d, X + j * d, d), j);
                }
            } else {
                KdQuery KQ = { &T, Q + qi * d, k, heap, 0, xkind, xbnd, qk, qb };
                kd_search(&KQ, 0);
                cnt = KQ.cnt;
            }
            /* heap property holds for cnt entries; sort ascending */
            heap_sort_ascending(heap, cnt);
            for (int t = 0; t < k; ++t) {
                idx_out[qi * k + t] = t < cnt ? heap[t].idx : -1;
                if (d2_out) d2_out[qi * k + t] = t < cnt ? heap[t].d2 : INFINITY;
            }
        }
        free(heap);
    }
    if (!brute) { free(T.perm); free(T.nodes); }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * basis functions
 * ---------------------------------------------------------------------------------------------- */

/* number of monomials of total degree <= deg in d variables */
int orc_num_monomials(int d, int deg) {
    int64_t c = 1;
    for (int t = 1; t <= d; ++t) c = c * (deg + t) / t;
    return (int)c;
}

/* graded exponent table (degree 0, 1, ...; within a degree x-major).  The column order of P does not
 * change the first n solution entries except by rounding (reference: polynomialbasis.jl:8-11). */
int orc_monomial_exponents(int d, int deg, int32_t* ex /* q*3 */) {
    int q = 0;
    for (int g = 0; g <= deg; ++g) {
        for (int a = g; a >= 0; --a) {
            if (d == 1) { if (a == g) { ex[q * 3] = a; ex[q * 3 + 1] = 0; ex[q * 3 + 2] = 0; q++; } continue; }
            for (int b = g - a; b >= 0; --b) {
                int c = g - a - b;
                if (d == 2 && c != 0) continue;
                ex[q * 3] = a; ex[q * 3 + 1] = b; ex[q * 3 + 2] = c; q++;
            }
        }
    }
    return q;
}

typedef struct { double coef; int e[3]; int rpow; } OrcTerm;  /* coef * prod x_a^e_a * r^rpow */

/* closed form of d^alpha (r^p) as a term list (restates what Symbolics derives in rbfbasis.jl:20-30,
 * rbfbasis_k.jl:9-18).  Returns the number of terms. */
int orc_rbf_derivative_terms(int p, int d, const int32_t* alpha, OrcTerm* out) {
    OrcTerm cur[ORC_MAX_TERMS], nxt[ORC_MAX_TERMS];
    int nc = 1;
    cur[0].coef = 1.0; cur[0].e[0] = cur[0].e[1] = cur[0].e[2] = 0; cur[0].rpow = p;
    for (int a = 0; a < d; ++a) {
        for (int rep = 0; rep < alpha[a]; ++rep) {
            int nn = 0;
            for (int t = 0; t < nc; ++t) {
                OrcTerm cands[2]; int ncand = 0;
                if (cur[t].e[a] > 0) {
                    cands[ncand] = cur[t]; cands[ncand].coef *= cur[t].e[a]; cands[ncand].e[a] -= 1; ncand++;
                }
                if (cur[t].rpow != 0) {
                    cands[ncand] = cur[t]; cands[ncand].coef *= cur[t].rpow; cands[ncand].e[a] += 1;
                    cands[ncand].rpow -= 2; ncand++;
                }
                for (int c = 0; c < ncand; ++c) {
                    int found = -1;
                    for (int u = 0; u < nn; ++u)
                        if (nxt[u].e[0] == cands[c].e[0] && nxt[u].e[1] == cands[c].e[1] &&
                            nxt[u].e[2] == cands[c].e[2] && nxt[u].rpow == cands[c].rpow) { found = u; break; }
                    if (found >= 0) nxt[found].coef += cands[c].coef;
                    else if (nn < ORC_MAX_TERMS) nxt[nn++] = cands[c];
                }
            }
            nc = 0;
            for (int u = 0; u < nn; ++u) if (nxt[u].coef != 0.0) cur[nc++] = nxt[u];
        }
    }
    for (int t = 0; t < nc; ++t) out[t] = cur[t];
    return nc;
}

/* C-callable flat version for tests: out = [coef, e0, e1, e2, rpow] * nterms */
int orc_rbf_derivative_table(int p, int d, const int32_t* alpha, double* out) {
    OrcTerm T[ORC_MAX_TERMS];
    int n = orc_rbf_derivative_terms(p, d, alpha, T);
    for (int t = 0; t < n; ++t) {
        out[5 * t] = T[t].coef; out[5 * t + 1] = T[t].e[0]; out[5 * t + 2] = T[t].e[1];
        out[5 * t + 3] = T[t].e[2]; out[5 * t + 4] = T[t].rpow;
    }
    return n;
}

static inline double ipow(double x, int e) {
    double r = 1.0;
    for (int t = 0; t < e; ++t) r *= x;
    return r;
}

static double eval_terms(const OrcTerm* T, int nt, const double* x, int d) {
    double r2 = 0.0;
    for (int a = 0; a < d; ++a) r2 += x[a] * x[a];
    double r = sqrt(r2);
    double s = 0.0;
    for (int t = 0; t < nt; ++t) {
        double v = T[t].coef;
        for (int a = 0; a < d; ++a) v *= ipow(x[a], T[t].e[a]);
        v *= pow(r, (double)T[t].rpow);
        s += v;
    }
    return s;
}

static double eval_mono_deriv(const int32_t* e, const int32_t* alpha, const double* x, int d) {
    double v = 1.0;
    for (int a = 0; a < d; ++a) {
        if (e[a] < alpha[a]) return 0.0;
        for (int t = 0; t < alpha[a]; ++t) v *= (double)(e[a] - t);
        v *= ipow(x[a], e[a] - alpha[a]);
    }
    return v;
}

/* ------------------------------------------------------------------------------------------------
 * dense kernels (LAPACK getrf / getri / getrs restated, column-major-agnostic: row-major here)
 * ---------------------------------------------------------------------------------------------- */

/* partial-pivot LU in place, row-major m x m, piv[k] = row swapped with k.  returns 0 or k+1 if singular */
static int lu_factor(double* A, int m, int* piv) {
    for (int k = 0; k < m; ++k) {
        int p = k; double best = fabs(A[k * m + k]);
        for (int i = k + 1; i < m; ++i) {
            double v = fabs(A[i * m + k]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (best == 0.0) return k + 1;
        if (p != k) for (int j = 0; j < m; ++j) { double t = A[k * m + j]; A[k * m + j] = A[p * m + j]; A[p * m + j] = t; }
        double rinv = 1.0 / A[k * m + k];
        for (int i = k + 1; i < m; ++i) {
            double l = A[i * m + k] * rinv;
            A[i * m + k] = l;
            for (int j = k + 1; j < m; ++j) A[i * m + j] -= l * A[k * m + j];
        }
    }
    return 0;
}

static void lu_solve(const double* LU, const int* piv, int m, double* b, int nrhs /* b is m x nrhs row-major */) {
    /* getrs: all row interchanges first (lu_factor swaps whole rows, L part included), then L, then U */
    for (int k = 0; k < m; ++k)
        if (piv[k] != k) for (int c = 0; c < nrhs; ++c) { double t = b[k * nrhs + c]; b[k * nrhs + c] = b[piv[k] * nrhs + c]; b[piv[k] * nrhs + c] = t; }
    for (int k = 0; k < m; ++k) {
        for (int i = k + 1; i < m; ++i) {
            double l = LU[i * m + k];
            for (int c = 0; c < nrhs; ++c) b[i * nrhs + c] -= l * b[k * nrhs + c];
        }
    }
    for (int k = m - 1; k >= 0; --k) {
        for (int c = 0; c < nrhs; ++c) {
            double s = b[k * nrhs + c];
            for (int j = k + 1; j < m; ++j) s -= LU[k * m + j] * b[j * nrhs + c];
            b[k * nrhs + c] = s / LU[k * m + k];
        }
    }
}

/* inv(A) from its LU, the getri way: inv(U), then solve X*L = inv(U), then undo the row swaps as
 * column swaps (interpolationmatrix.jl:8 -> LinearAlgebra.inv -> getrf!+getri!). */
static void lu_inverse(const double* LU, const int* piv, int m, double* Ainv, double* work) {
    /* work = inv(U), upper triangular */
    memset(work, 0, sizeof(double) * (size_t)m * m);
    for (int j = 0; j < m; ++j) {
        work[j * m + j] = 1.0 / LU[j * m + j];
        for (int i = j - 1; i >= 0; --i) {
            double s = 0.0;
            for (int t = i + 1; t <= j; ++t) s += LU[i * m + t] * work[t * m + j];
            work[i * m + j] = -s / LU[i * m + i];
        }
    }
    /* X * L = inv(U)  (L unit lower): process columns right to left */
    memcpy(Ainv, work, sizeof(double) * (size_t)m * m);
    for (int j = m - 2; j >= 0; --j) {
        for (int i = 0; i < m; ++i) {
            double s = Ainv[i * m + j];
            for (int t = j + 1; t < m; ++t) s -= Ainv[i * m + t] * LU[t * m + j];
            Ainv[i * m + j] = s;
        }
    }
    for (int k = m - 1; k >= 0; --k) {
        if (piv[k] != k) for (int i = 0; i < m; ++i) { double t = Ainv[i * m + k]; Ainv[i * m + k] = Ainv[i * m + piv[k]]; Ainv[i * m + piv[k]] = t; }
    }
}

/* extended-precision LU solve (adjudication mode: "who is closer to the truth") */
static int lu_solve_ld(const double* A, int m, double* b, int nrhs) {
    long double* W = (long double*)malloc(sizeof(long double) * (size_t)m * (m + nrhs));
    int w = m + nrhs;
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < m; ++j) W[i * w + j] = A[i * m + j];
        for (int c = 0; c < nrhs; ++c) W[i * w + m + c] = b[i * nrhs + c];
    }
    for (int k = 0; k < m; ++k) {
        int p = k; long double best = fabsl(W[k * w + k]);
        for (int i = k + 1; i < m; ++i) if (fabsl(W[i * w + k]) > best) { best = fabsl(W[i * w + k]); p = i; }
        if (best == 0.0L) { free(W); return k + 1; }
        if (p != k) for (int j = 0; j < w; ++j) { long double t = W[k * w + j]; W[k * w + j] = W[p * w + j]; W[p * w + j] = t; }
        for (int i = k + 1; i < m; ++i) {
            long double l = W[i * w + k] / W[k * w + k];
            for (int j = k + 1; j < w; ++j) W[i * w + j] -= l * W[k * w + j];
        }
    }
    for (int k = m - 1; k >= 0; --k)
        for (int c = 0; c < nrhs; ++c) {
            long double s = W[k * w + m + c];
            for (int j = k + 1; j < m; ++j) s -= W[k * w + j] * W[j * w + m + c];
            W[k * w + m + c] = s / W[k * w + k];
        }
    for (int i = 0; i < m; ++i) for (int c = 0; c < nrhs; ++c) b[i * nrhs + c] = (double)W[i * w + m + c];
    free(W);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * weights
 * ---------------------------------------------------------------------------------------------- */

/*
 * X[N][d], Y[M][d]; idx[N][n] stencil of every X node (idx[i][0] is the stencil centre, scalestencil.jl:10);
 * center[M] = nearest X node of every Y row.  ops[nops][4] = {kind, a0, a1, a2}.
 * mode: 0 = literal reference arithmetic  W = inv(A) * RHS   (generate_operator.jl:158)
 *       1 = LU solve in double, 2 = LU solve in long double.
 * variant: 0 = scaled two-set method (generate_operator.jl:29-190)
 *          1 = legacy collocated method (generate_operator.jl:354-491, hyperviscosity_operator.jl:314-440): no
 *              scaling, the centre node becomes (eps,eps) in A and in both right-hand-side blocks, RBF rows are
 *              evaluated at X_shift[j] = X_j - x_c (:433), so odd derivatives change sign; weights not rescaled.
 * vals[nops][M][n] (row-major), cond1[N] (1-norm condition number of A_i; NULL to skip; needs mode 0).
 * returns 0, or 1+node index of the first singular stencil.
 */
int orc_weights(const double* X, int64_t N, int d, const double* Y, int64_t M,
                const int64_t* idx, const int64_t* center, int p, int n, int polydeg,
                int nops, const int32_t* ops, int mode, int variant, double* vals, double* cond1) {
    if (d < 1 || d > ORC_MAX_DIM) return -1;
    int q = orc_num_monomials(d, polydeg);
    int m = n + q;
    int32_t* ex = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)q);
    orc_monomial_exponents(d, polydeg, ex);
    const double EPS = 2.220446049250313e-16;

    /* term tables per op (for ORC_OP_LAPLACE: one table per axis) */
    OrcTerm* terms = (OrcTerm*)malloc(sizeof(OrcTerm) * ORC_MAX_TERMS * (size_t)nops * 3);
    int* nterms = (int*)calloc((size_t)nops * 3, sizeof(int));
    for (int o = 0; o < nops; ++o) {
        if (ops[4 * o] == ORC_OP_DERIV) {
            nterms[3 * o] = orc_rbf_derivative_terms(p, d, ops + 4 * o + 1, terms + (size_t)(3 * o) * ORC_MAX_TERMS);
        } else {
            for (int a = 0; a < d; ++a) {
                int32_t al[3] = {0, 0, 0}; al[a] = 2;
                nterms[3 * o + a] = orc_rbf_derivative_terms(p, d, al, terms + (size_t)(3 * o + a) * ORC_MAX_TERMS);
            }
        }
    }

    /* rows grouped by centre (counting sort) */
    int64_t* start = (int64_t*)calloc((size_t)N + 1, sizeof(int64_t));
    int64_t* rows = (int64_t*)malloc(sizeof(int64_t) * (size_t)(M > 0 ? M : 1));
    for (int64_t k = 0; k < M; ++k) start[center[k] + 1]++;
    for (int64_t i = 0; i < N; ++i) start[i + 1] += start[i];
    {
        int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)N);
        memcpy(fill, start, sizeof(int64_t) * (size_t)N);
        for (int64_t k = 0; k < M; ++k) rows[fill[center[k]]++] = k;
        free(fill);
    }

    int64_t fail = 0;
#pragma omp parallel
    {
        double* S = (double*)malloc(sizeof(double) * (size_t)n * d);
        double* A = (double*)malloc(sizeof(double) * (size_t)m * m);
        double* LU = (double*)malloc(sizeof(double) * (size_t)m * m);
        double* Ainv = (double*)malloc(sizeof(double) * (size_t)m * m);
        double* work = (double*)malloc(sizeof(double) * (size_t)m * m);
        double* rhs = (double*)malloc(sizeof(double) * (size_t)m * nops);
        double* sol = (double*)malloc(sizeof(double) * (size_t)m * nops);
        int* piv = (int*)malloc(sizeof(int) * (size_t)m);
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < N; ++i) {
            if (start[i + 1] == start[i]) { if (cond1) cond1[i] = 0.0; continue; }   /* no row uses this stencil */
            const int64_t* st = idx + i * n;
            double s[3] = {1.0, 1.0, 1.0};
            const double* xc = X + st[0] * d;
            /* scalestencil.jl:10-20 */
            for (int j = 0; j < n; ++j) for (int a = 0; a < d; ++a) S[j * d + a] = X[st[j] * d + a] - xc[a];
            /* legacy collocated method: no scaling, centre node replaced by (eps, eps)  generate_operator.jl:403-410 */
            if (variant != 0) for (int a = 0; a < d; ++a) S[a] = EPS;
            if (variant == 0) {
                for (int a = 0; a < d; ++a) {
                    double mx = 0.0;
                    for (int j = 0; j < n; ++j) { double v = fabs(S[j * d + a]); if (v > mx) mx = v; }
                    s[a] = 1.0 / mx;
                }
                for (int j = 0; j < n; ++j) for (int a = 0; a < d; ++a) S[j * d + a] = S[j * d + a] * s[a];
            }
            /* interpolationmatrix.jl:5, rbfblock.jl:14-20, polynomialblock.jl:26-31 */
            memset(A, 0, sizeof(double) * (size_t)m * m);
            for (int a_ = 0; a_ < n; ++a_) {
                for (int b_ = 0; b_ < n; ++b_) {
                    double r2 = 0.0;
                    for (int a = 0; a < d; ++a) { double t = S[a_ * d + a] - S[b_ * d + a]; r2 += t * t; }
                    A[a_ * m + b_] = pow(sqrt(r2), (double)p);
                }
                for (int t = 0; t < q; ++t) {
                    double v = 1.0;
                    for (int a = 0; a < d; ++a) v *= ipow(S[a_ * d + a], ex[3 * t + a]);
                    A[a_ * m + n + t] = v;
                    A[(n + t) * m + a_] = v;
                }
            }
            memcpy(LU, A, sizeof(double) * (size_t)m * m);
            int info = lu_factor(LU, m, piv);
            if (info) {
#pragma omp critical
                { if (!fail || i + 1 < fail) fail = i + 1; }
                continue;
            }
            if (mode == 0 || cond1) lu_inverse(LU, piv, m, Ainv, work);
            if (cond1) {
                double na = 0.0, ni = 0.0;
                for (int j = 0; j < m; ++j) {
                    double ca = 0.0, ci = 0.0;
                    for (int r_ = 0; r_ < m; ++r_) { ca += fabs(A[r_ * m + j]); ci += fabs(Ainv[r_ * m + j]); }
                    if (ca > na) na = ca;
                    if (ci > ni) ni = ci;
                }
                cond1[i] = na * ni;
            }
            for (int64_t rr = start[i]; rr < start[i + 1]; ++rr) {
                int64_t k = rows[rr];
                double eta[3] = {0, 0, 0};
                /* generate_operator.jl:110-120 (two-set) / :410 (legacy: centre := (eps, eps)) */
                for (int a = 0; a < d; ++a) eta[a] = variant == 0 ? (Y[k * d + a] - xc[a]) * s[a] : EPS;
                for (int o = 0; o < nops; ++o) {
                    const int32_t* op = ops + 4 * o;
                    for (int j = 0; j < n; ++j) {
                        double del[3];
                        for (int a = 0; a < d; ++a) {
                            /* :125-133 (two-set: eta - S_j, 0 -> eps) ; :433 (legacy: S_j - centre) */
                            double t = variant == 0 ? eta[a] - S[j * d + a] : S[j * d + a];
                            if (variant == 0 && t == 0.0) t = EPS;
                            del[a] = t;
                        }
                        double v;
                        if (op[0] == ORC_OP_DERIV) {
                            v = eval_terms(terms + (size_t)(3 * o) * ORC_MAX_TERMS, nterms[3 * o], del, d);
                        } else {
                            v = 0.0;
                            for (int a = 0; a < d; ++a)
                                v += s[a] * s[a] * eval_terms(terms + (size_t)(3 * o + a) * ORC_MAX_TERMS, nterms[3 * o + a], del, d);
                        }
                        rhs[j * nops + o] = v;
                    }
                    /* polylinearoperator.jl:36-44: polynomial rows at the scaled evaluation point */
                    double pe[3];
                    for (int a = 0; a < d; ++a) pe[a] = eta[a];   /* legacy: polylinearoperator([X_shift[1]], ...) = (eps, eps)  :429 */
                    for (int t = 0; t < q; ++t) {
                        double v;
                        if (op[0] == ORC_OP_DERIV) v = eval_mono_deriv(ex + 3 * t, op + 1, pe, d);
                        else {
                            v = 0.0;
                            for (int a = 0; a < d; ++a) {
                                int32_t al[3] = {0, 0, 0}; al[a] = 2;
                                v += s[a] * s[a] * eval_mono_deriv(ex + 3 * t, al, pe, d);
                            }
                        }
                        rhs[(n + t) * nops + o] = v;
                    }
                }
                if (mode == 0) {
                    /* stenc = M_inv * RHS  (generate_operator.jl:158) */
                    for (int r_ = 0; r_ < n; ++r_)
                        for (int o = 0; o < nops; ++o) {
                            double acc = 0.0;
                            for (int t = 0; t < m; ++t) acc += Ainv[r_ * m + t] * rhs[t * nops + o];
                            sol[r_ * nops + o] = acc;
                        }
                } else if (mode == 1) {
                    memcpy(sol, rhs, sizeof(double) * (size_t)m * nops);
                    lu_solve(LU, piv, m, sol, nops);
                } else {
                    memcpy(sol, rhs, sizeof(double) * (size_t)m * nops);
                    lu_solve_ld(A, m, sol, nops);
                }
                /* generate_operator.jl:161-166, hyperviscosity_operator.jl:159-160 */
                for (int o = 0; o < nops; ++o) {
                    const int32_t* op = ops + 4 * o;
                    double f = 1.0;
                    if (op[0] == ORC_OP_DERIV)
                        for (int a = 0; a < d; ++a) f *= ipow(s[a], op[1 + a]);
                    for (int j = 0; j < n; ++j) vals[((size_t)o * M + k) * n + j] = f * sol[j * nops + o];
                }
            }
        }
        free(S); free(A); free(LU); free(Ainv); free(work); free(rhs); free(sol); free(piv);
    }
    free(ex); free(terms); free(nterms); free(start); free(rows);
    return (int)fail;
}

/* ------------------------------------------------------------------------------------------------
 * operator application (examples/adv_diff_test.jl:151-152)
 * ---------------------------------------------------------------------------------------------- */

/* y = alpha * A x + beta * y,  A fixed-row-length CSR (M rows, n entries per row) */
void orc_spmv(int64_t M, int n, const int64_t* colind, const double* vals, const double* x,
              double alpha, double beta, double* y) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < M; ++k) {
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc += vals[k * n + j] * x[colind[k * n + j]];
        y[k] = beta == 0.0 ? alpha * acc : alpha * acc + beta * y[k];
    }
}

/* y (length N) = alpha * A' v + beta * y   (E' * v, adv_diff_test.jl:151); serial scatter like CSC */
void orc_spmv_t(int64_t M, int64_t N, int n, const int64_t* colind, const double* vals, const double* v,
                double alpha, double beta, double* y) {
    for (int64_t i = 0; i < N; ++i) y[i] = beta == 0.0 ? 0.0 : beta * y[i];
    for (int64_t k = 0; k < M; ++k)
        for (int j = 0; j < n; ++j) y[colind[k * n + j]] += alpha * vals[k * n + j] * v[k];
}

/* du = E' * (alpha*Dxx*u + alpha*Dyy*u - ux*Dx*u - uy*Dy*u) - gamma * (Dxk + Dyk) * u   (adv_diff_test.jl:151-152)
 * all operators share colind (M rows, n per row, N columns). work has M doubles. */
void orc_rhs_advdiff(int64_t M, int64_t N, int n, const int64_t* colind,
                     const double* E, const double* Dx, const double* Dy, const double* Dxx, const double* Dyy,
                     const double* Dxk, const double* Dyk, double alpha, double ux, double uy, double gamma,
                     const double* u, double* du, double* work) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < M; ++k) {
        double axx = 0, ayy = 0, ax = 0, ay = 0;
        for (int j = 0; j < n; ++j) {
            double uj = u[colind[k * n + j]];
            axx += Dxx[k * n + j] * uj; ayy += Dyy[k * n + j] * uj;
            ax += Dx[k * n + j] * uj; ay += Dy[k * n + j] * uj;
        }
        work[k] = alpha * axx + alpha * ayy - ux * ax - uy * ay;
    }
    orc_spmv_t(M, N, n, colind, E, work, 1.0, 0.0, du);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < M; ++k) {
        if (k >= N) continue;
        double hv = 0;
        for (int j = 0; j < n; ++j) hv += (Dxk[k * n + j] + Dyk[k * n + j]) * u[colind[k * n + j]];
        du[k] -= gamma * hv;
    }
}
