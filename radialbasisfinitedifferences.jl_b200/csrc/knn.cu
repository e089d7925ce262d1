// knn.cu -- K1/K2/K3 of SURVEY.md §2: grid-binned EXACT k-nearest-neighbour search on the GPU.
//
// Replaces  KDTree(X); knn(tree, X, n, true); knn(tree, Y, 1)      (src/generate_operator.jl:43-47,
// src/hyperviscosity_operator.jl:54-58) and the masked per-boundary searches of
// src/calculateneighbors.jl:16-42,83-94.
//
// Results are ordered by (squared distance, caller's node index); the squared distance is accumulated
// as ((dx*dx)+dy*dy)+dz*dz with separate roundings (no FMA) so the CPU oracle reproduces it bit for bit.
//
// Layout in HBM: points are counting-sorted by grid cell (x fastest), coordinates stored SoA
// (xs[a*N + i]) so a run of cells along x is ONE contiguous range of points; perm[i] maps a sorted
// slot back to the caller's index (tie-breaking and outputs use the caller's index, the internal
// order is invisible at the API).  One thread owns one query and keeps its k-candidate max-heap in
// shared memory as a column ([slot][thread], padded stride) so heap traffic is bank-conflict free;
// threads of a warp are spatial neighbours (queries are issued in cell order) so their candidate
// loads hit the same L1 lines.
#include <cub/device/device_radix_sort.cuh>
#include <cfloat>
#include <cmath>

#include <cstdlib>
#include "common.cuh"

namespace {

struct Grid {
    double lo[3], h[3], inv_h[3];
    int n[3];
    int ncells;
};

__device__ __forceinline__ unsigned long long enc_double(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
inline double dec_double(unsigned long long u) {
    unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    double v;
    memcpy(&v, &b, 8);
    return v;
}

__global__ void bbox_kernel(const double* __restrict__ X, int64_t N, int d, unsigned long long* mn,
                            unsigned long long* mx) {
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        for (int a = 0; a < d; ++a) {
            double v = X[i * d + a];
            if (!isfinite(v)) mx[3] = 1ull;       // fmin/fmax drop NaN silently: flag it
            lo[a] = fmin(lo[a], v);
            hi[a] = fmax(hi[a], v);
        }
    }
    // warp reduction, then ONE pair of atomics per axis and CTA (one per warp put 38 k atomics on six addresses at 1 M points:
    // their serialisation in L2 was most of this kernel's 32 us)
    __shared__ double wlo[8][3], whi[8][3];
    for (int a = 0; a < d; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) { wlo[(threadIdx.x >> 5) & 7][a] = lo[a]; whi[(threadIdx.x >> 5) & 7][a] = hi[a]; }
    }
    __syncthreads();
    if ((int)threadIdx.x < d) {
        const int a = threadIdx.x, nw = (blockDim.x + 31) >> 5;
        double l = wlo[0][a], h = whi[0][a];
        for (int w = 1; w < nw; ++w) { l = fmin(l, wlo[w][a]); h = fmax(h, whi[w][a]); }
        atomicMin(&mn[a], enc_double(l));
        atomicMax(&mx[a], enc_double(h));
    }
}

__device__ __forceinline__ int cell_coord(double x, double lo, double inv_h, int n) {
    double t = (x - lo) * inv_h;
    int c = t > 0.0 ? (t < (double)n ? (int)t : n - 1) : 0;
    return c;
}

__global__ void cell_id_kernel(const double* __restrict__ X, int64_t N, int d, Grid g, int* __restrict__ cid,
                               int* __restrict__ ident) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    int c = cell_coord(X[i * d], g.lo[0], g.inv_h[0], g.n[0]);
    if (d > 1) c += g.n[0] * cell_coord(X[i * d + 1], g.lo[1], g.inv_h[1], g.n[1]);
    if (d > 2) c += g.n[0] * g.n[1] * cell_coord(X[i * d + 2], g.lo[2], g.inv_h[2], g.n[2]);
    cid[i] = c;
    ident[i] = (int)i;
}

__global__ void gather_sorted_kernel(const double* __restrict__ X, int64_t N, int d, const int* __restrict__ perm,
                                     const int32_t* __restrict__ xgroup, double* __restrict__ xs,
                                     int32_t* __restrict__ sgroup) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    int p = perm[i];
    for (int a = 0; a < d; ++a) xs[a * N + i] = X[(int64_t)p * d + a];
    if (sgroup) sgroup[i] = xgroup[p];
}

// cell_start[c] = first sorted slot whose cell id >= c;  cell_start[ncells] = N
__global__ void cell_start_kernel(const int* __restrict__ cid_sorted, int64_t N, int ncells, int* __restrict__ cell_start) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i > N) return;
    int prev = i == 0 ? -1 : cid_sorted[i - 1];
    int cur = i == N ? ncells : cid_sorted[i];
    for (int c = prev + 1; c <= cur; ++c) cell_start[c] = (int)i;
}

// visibility rule of calculateneighbors.jl:16-42,83-87.  group code: 0 interior, 1+2b boundary b, 2+2b ghost b
__device__ __forceinline__ bool visible(int qg, int cg) {
    if (qg == 0) return true;                        // interior queries see every node (:83-87)
    if (cg > 0 && (cg & 1) == 0) return cg == 2 + 2 * ((qg - 1) >> 1);   // ghosts only of the query's boundary (:30)
    return true;                                     // interior + all boundary nodes (:24-28)
}

struct KnnArgs {
    Grid g;
    const double* xs;          // SoA sorted coordinates [D][N]
    const int* perm;           // sorted slot -> caller index
    const int* cell_start;     // [ncells+1]
    const int32_t* sgroup;     // sorted group codes or null
    const double* Q;           // queries AoS [NQ][D]; null => queries are the sorted points themselves
    const int32_t* qgroup;     // group code per query (caller order) or null
    int64_t N, NQ;
    int k;
    int32_t* idx_out;          // [NQ][k] caller order
    double* d2_out;            // [NQ][k] or null
    int* short_rows;           // counts rows with fewer than k visible candidates
};

constexpr int KNN_BS = 128;

// One thread per query.  The k-candidate max-heap of a thread is a column of shared memory holding, per entry, the squared
// distance ROUNDED TO FLOAT and the candidate's slot in the cell-sorted arrays: 8 bytes instead of the 12 of (double, caller
// index), i.e. 1.5x - 2x the resident warps of this latency-bound kernel (at k = 60 the heap alone is 62 KB per 128 queries).
// Rounding to float is monotone, so float keys that differ order the candidates exactly as the doubles do; only when two float
// keys are EQUAL are the two squared distances recomputed in double from the coordinates (same ((dx*dx)+dy*dy)+dz*dz arithmetic)
// and, on an exact tie, the caller indices compared -- results are bit-identical to a heap of doubles, exact-tie lattices
// included (they just take the slow comparison more often).
template <int D>
__global__ void __launch_bounds__(KNN_BS) knn_kernel(KnnArgs a) {
    extern __shared__ unsigned char knn_smem[];
    const int tid = threadIdx.x;
    const int S = KNN_BS + 1;
    const int k = a.k;
    float* hf = reinterpret_cast<float*>(knn_smem) + tid;
    int* hs = reinterpret_cast<int*>(knn_smem + sizeof(float) * (size_t)S * k) + tid;
    const int64_t t = blockIdx.x * (int64_t)KNN_BS + tid;
    if (t >= a.NQ) return;
    const Grid& g = a.g;

    double q[D];
    int64_t row;
    int qg = 0;
    if (a.Q) {
        row = t;
#pragma unroll
        for (int c = 0; c < D; ++c) q[c] = a.Q[t * D + c];
        if (a.qgroup) qg = a.qgroup[t];
    } else {
        row = a.perm[t];
#pragma unroll
        for (int c = 0; c < D; ++c) q[c] = a.xs[c * a.N + t];
        if (a.sgroup) qg = a.sgroup[t];
    }
    int cc[D];
    double slack = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        cc[c] = cell_coord(q[c], g.lo[c], g.inv_h[c], g.n[c]);
        slack = fmax(slack, 1e-9 * g.h[c]);
    }

    auto exact_d2 = [&](int i) -> double {
        double dx = __dsub_rn(q[0], a.xs[i]);
        double d2 = __dmul_rn(dx, dx);
        if (D > 1) {
            double dy = __dsub_rn(q[1], a.xs[a.N + i]);
            d2 = __dadd_rn(d2, __dmul_rn(dy, dy));
        }
        if (D > 2) {
            double dz = __dsub_rn(q[D - 1], a.xs[2 * a.N + i]);
            d2 = __dadd_rn(d2, __dmul_rn(dz, dz));
        }
        return d2;
    };
    // (d2, caller index) order on (float key, slot) pairs
    auto less_fs = [&](float fa, int sa, float fb, int sb) -> bool {
        if (fa != fb) return fa < fb;
        const double da = exact_d2(sa), db = exact_d2(sb);
        if (da != db) return da < db;
        return a.perm[sa] < a.perm[sb];
    };
    // put (vf, vs) into the hole at i0 of a max-heap of size n and sift it down
    auto sift_down = [&](int n, int i0, float vf, int vs) {
        int i = i0;
        for (;;) {
            int l = 2 * i + 1;
            if (l >= n) break;
            int c = l;
            float cf = hf[l * S];
            int cs = hs[l * S];
            if (l + 1 < n) {
                const float rf = hf[(l + 1) * S];
                const int rs = hs[(l + 1) * S];
                if (less_fs(cf, cs, rf, rs)) { c = l + 1; cf = rf; cs = rs; }
            }
            if (!less_fs(vf, vs, cf, cs)) break;
            hf[i * S] = cf;
            hs[i * S] = cs;
            i = c;
        }
        hf[i * S] = vf;
        hs[i * S] = vs;
    };

    int cnt = 0;
    float wf = FLT_MAX;            // float key of the current k-th candidate (heap root) once cnt == k
    int ws = 0;
    double wd_ub = DBL_MAX;        // an upper bound of its exact squared distance: the next float above wf

    auto set_root = [&]() {
        wf = hf[0];
        ws = hs[0];
        wd_ub = wf < FLT_MAX ? (double)__int_as_float(__float_as_int(wf) + 1) : DBL_MAX;
    };

    auto scan_run = [&](int y, int z, int x0, int x1) {
        if (cnt == k) {
            double m2 = 0.0;
            {
                double lox = g.lo[0] + x0 * g.h[0], hix = g.lo[0] + (x1 + 1) * g.h[0];
                double gp = fmax(lox - q[0], q[0] - hix) - slack;
                if (gp > 0.0) m2 += gp * gp;
            }
            if (D > 1) {
                double loy = g.lo[1] + y * g.h[1], hiy = loy + g.h[1];
                double gp = fmax(loy - q[1], q[1] - hiy) - slack;
                if (gp > 0.0) m2 += gp * gp;
            }
            if (D > 2) {
                double loz = g.lo[2] + z * g.h[2], hiz = loz + g.h[2];
                double gp = fmax(loz - q[D - 1], q[D - 1] - hiz) - slack;
                if (gp > 0.0) m2 += gp * gp;
            }
            if (m2 > wd_ub) return;
        }
        const int base = (D == 3 ? (z * g.n[1] + y) : (D == 2 ? y : 0)) * g.n[0];
        const int s = a.cell_start[base + x0], e = a.cell_start[base + x1 + 1];
        for (int i = s; i < e; ++i) {
            const double d2 = exact_d2(i);
            const float f = __double2float_rn(d2);
            if (cnt == k && f > wf) continue;                  // a larger float key is a larger squared distance
            if (a.sgroup && !visible(qg, a.sgroup[i])) continue;
            if (cnt < k) {
                // sift up
                int j = cnt;
                while (j > 0) {
                    int p = (j - 1) >> 1;
                    const float pf = hf[p * S];
                    const int ps = hs[p * S];
                    if (!less_fs(pf, ps, f, i)) break;
                    hf[j * S] = pf;
                    hs[j * S] = ps;
                    j = p;
                }
                hf[j * S] = f;
                hs[j * S] = i;
                if (++cnt == k) set_root();
            } else if (less_fs(f, i, wf, ws)) {
                sift_down(k, 0, f, i);
                set_root();
            }
        }
    };

    for (int R = 0;; ++R) {
        if (R > 0) {
            // lower bound on the distance of every point in rings >= R
            bool any = false;
            double lb = DBL_MAX;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                if (cc[c] - R >= 0) {
                    any = true;
                    lb = fmin(lb, q[c] - (g.lo[c] + (cc[c] - R + 1) * g.h[c]));
                }
                if (cc[c] + R <= g.n[c] - 1) {
                    any = true;
                    lb = fmin(lb, (g.lo[c] + (cc[c] + R) * g.h[c]) - q[c]);
                }
            }
            if (!any) break;                       // grid exhausted
            lb -= slack;
            // exact test against the k-th candidate (its distance recomputed in double; the float bound first)
            if (cnt == k && lb > 0.0 && lb * lb > wd_ub) break;
            if (cnt == k && lb > 0.0 && lb * lb > exact_d2(ws)) break;
        }
        const int x0 = max(cc[0] - R, 0), x1 = min(cc[0] + R, g.n[0] - 1);
        if (D == 1) {
            if (R == 0) scan_run(0, 0, x0, x1);
            else {
                if (cc[0] - R >= 0) scan_run(0, 0, cc[0] - R, cc[0] - R);
                if (cc[0] + R < g.n[0]) scan_run(0, 0, cc[0] + R, cc[0] + R);
            }
        } else if (D == 2) {
            for (int dy = -R; dy <= R; ++dy) {
                const int y = cc[1] + dy;
                if (y < 0 || y >= g.n[1]) continue;
                if (dy == -R || dy == R) {
                    if (x0 <= x1) scan_run(y, 0, x0, x1);
                } else {
                    if (cc[0] - R >= 0) scan_run(y, 0, cc[0] - R, cc[0] - R);
                    if (cc[0] + R < g.n[0]) scan_run(y, 0, cc[0] + R, cc[0] + R);
                }
            }
        } else {
            for (int dz = -R; dz <= R; ++dz) {
                const int z = cc[D - 1] + dz;
                if (z < 0 || z >= g.n[D - 1]) continue;
                const bool zface = (dz == -R || dz == R);
                for (int dy = -R; dy <= R; ++dy) {
                    const int y = cc[1] + dy;
                    if (y < 0 || y >= g.n[1]) continue;
                    if (zface || dy == -R || dy == R) {
                        if (x0 <= x1) scan_run(y, z, x0, x1);
                    } else {
                        if (cc[0] - R >= 0) scan_run(y, z, cc[0] - R, cc[0] - R);
                        if (cc[0] + R < g.n[0]) scan_run(y, z, cc[0] + R, cc[0] + R);
                    }
                }
            }
        }
    }

    // heap sort -> ascending by (d2, idx)
    for (int e = cnt - 1; e > 0; --e) {
        const float tf = hf[e * S];
        const int ts = hs[e * S];
        hf[e * S] = hf[0];
        hs[e * S] = hs[0];
        sift_down(e, 0, tf, ts);
    }
    if (cnt < k) atomicAdd(a.short_rows, 1);
    int32_t* orow = a.idx_out + row * k;
    for (int j = 0; j < k; ++j) orow[j] = j < cnt ? a.perm[hs[j * S]] : -1;
    if (a.d2_out) {
        double* drow = a.d2_out + row * k;
        for (int j = 0; j < k; ++j) drow[j] = j < cnt ? exact_d2(hs[j * S]) : INFINITY;
    }
}

struct Bins {
    Grid g;
    DevBuf<double> xs;
    DevBuf<int> perm;
    DevBuf<int> cell_start;
    DevBuf<int32_t> sgroup;
    bool has_groups = false;
    int64_t N = 0;
    int dim = 0;
};

int build_bins(rbffd_context* ctx, const double* X, int64_t N, int dim, int k_hint, const int32_t* xgroup, Bins& B) {
    cudaStream_t st = ctx->stream;
    B.N = N;
    B.dim = dim;
    // bounding box
    DevBuf<unsigned long long> mm;
    CUDA_TRY(ctx, mm.alloc(7, st));
    unsigned long long init[7] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull, 0ull};
    CUDA_TRY(ctx, cudaMemcpyAsync(mm.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    int blocks = (int)std::min<int64_t>((N + 255) / 256, (int64_t)ctx->sm_count * 8);
    bbox_kernel<<<blocks, 256, 0, st>>>(X, N, dim, mm.p, mm.p + 3);
    KLAUNCH(ctx);
    unsigned long long h_mm[7];
    CUDA_TRY(ctx, cudaMemcpyAsync(h_mm, mm.p, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (h_mm[6]) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "non-finite node coordinates");
    double lo[3] = {0, 0, 0}, ext[3] = {0, 0, 0};
    double vol = 1.0;
    int deff = 0;
    for (int a = 0; a < dim; ++a) {
        lo[a] = dec_double(h_mm[a]);
        double hi = dec_double(h_mm[3 + a]);
        if (!std::isfinite(lo[a]) || !std::isfinite(hi))
            RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "non-finite node coordinates");
        ext[a] = hi - lo[a];
        if (ext[a] > 0) { vol *= ext[a]; deff++; }
    }
    // points per cell so that two rings of cells usually enclose the k-ball (kNN)
    const double unit_ball[4] = {1.0, 2.0, 3.141592653589793, 4.18879020478639};
    // points per cell: two rings of cells usually enclose the k-ball; RBFFD_KNN_PPC rescales the default factor 1.15 (tuning knob)
    static const double ppc_factor = [] { const char* e = getenv("RBFFD_KNN_PPC"); const double f = e ? atof(e) : 0.0; return f > 0.0 ? f : 1.15; }();
    double ppc = ppc_factor * k_hint / (unit_ball[dim] * std::pow(2.0, dim));
    ppc = std::min(std::max(ppc, 1.0), 16.0);
    double h = 1.0;
    if (deff > 0) h = std::pow(vol * ppc / (double)N, 1.0 / deff);
    if (!(h > 0) || !std::isfinite(h)) h = 1.0;
    Grid& g = B.g;
    int64_t ncells = 1;
    for (int a = 0; a < 3; ++a) {
        g.lo[a] = a < dim ? lo[a] : 0.0;
        g.h[a] = h;
        g.inv_h[a] = 1.0 / h;
        int64_t na = a < dim ? (int64_t)std::ceil(ext[a] / h) : 1;
        if (na < 1) na = 1;
        if (na > (1 << 20)) na = 1 << 20;
        g.n[a] = (int)na;
        ncells *= na;
    }
    while (ncells > 8 * N + 4096 || ncells > (1ll << 30)) {   // degenerate aspect ratios: coarsen
        h *= 1.5;
        ncells = 1;
        for (int a = 0; a < 3; ++a) {
            g.h[a] = h;
            g.inv_h[a] = 1.0 / h;
            int64_t na = a < dim ? (int64_t)std::ceil(ext[a] / h) : 1;
            if (na < 1) na = 1;
            g.n[a] = (int)na;
            ncells *= na;
        }
    }
    g.ncells = (int)ncells;

    DevBuf<int> cid, ident, cid_sorted;
    CUDA_TRY(ctx, cid.alloc(N, st));
    CUDA_TRY(ctx, ident.alloc(N, st));
    CUDA_TRY(ctx, cid_sorted.alloc(N, st));
    CUDA_TRY(ctx, B.perm.alloc(N, st));
    CUDA_TRY(ctx, B.xs.alloc((size_t)N * dim, st));
    CUDA_TRY(ctx, B.cell_start.alloc((size_t)ncells + 1, st));
    B.has_groups = xgroup != nullptr;
    if (xgroup) CUDA_TRY(ctx, B.sgroup.alloc(N, st));
    const int nb = ceil_div_i64(N, 256);
    cell_id_kernel<<<nb, 256, 0, st>>>(X, N, dim, g, cid.p, ident.p);
    KLAUNCH(ctx);
    int bits = 1;
    while ((1ll << bits) < ncells) ++bits;
    size_t tmp_bytes = 0;
    CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, cid.p, cid_sorted.p, ident.p, B.perm.p, (int)N, 0, bits, st));
    DevBuf<unsigned char> tmp;
    CUDA_TRY(ctx, tmp.alloc(tmp_bytes, st));
    CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, cid.p, cid_sorted.p, ident.p, B.perm.p, (int)N, 0, bits, st));
    gather_sorted_kernel<<<nb, 256, 0, st>>>(X, N, dim, B.perm.p, xgroup, B.xs.p, xgroup ? B.sgroup.p : nullptr);
    cell_start_kernel<<<ceil_div_i64(N + 1, 256), 256, 0, st>>>(cid_sorted.p, N, (int)ncells, B.cell_start.p);
    KLAUNCH(ctx); KLAUNCH(ctx);
    CUDA_TRY(ctx, cudaGetLastError());
    return RBFFD_OK;
}

// short_rows_dev != nullptr: the count of queries with fewer than k visible points goes to that (zeroed) device word and the
// caller fetches it later -- several searches then share ONE status round trip instead of stalling the stream after each
int run_knn(rbffd_context* ctx, const Bins& B, const double* Q, int64_t NQ, const int32_t* qgroup, int k,
            int32_t* idx_out, double* d2_out, int* short_rows_dev = nullptr) {
    cudaStream_t st = ctx->stream;
    if (NQ == 0) return RBFFD_OK;
    KnnArgs a;
    a.g = B.g;
    a.xs = B.xs.p;
    a.perm = B.perm.p;
    a.cell_start = B.cell_start.p;
    a.sgroup = B.has_groups ? B.sgroup.p : nullptr;
    a.Q = Q;
    a.qgroup = (B.has_groups && Q) ? qgroup : nullptr;
    a.N = B.N;
    a.NQ = NQ;
    a.k = k;
    a.idx_out = idx_out;
    a.d2_out = d2_out;
    DevBuf<int> flag;
    if (short_rows_dev) a.short_rows = short_rows_dev;
    else {
        CUDA_TRY(ctx, flag.alloc(1, st));
        CUDA_TRY(ctx, cudaMemsetAsync(flag.p, 0, sizeof(int), st));
        a.short_rows = flag.p;
    }
    size_t smem = (size_t)(KNN_BS + 1) * k * (sizeof(float) + sizeof(int));
    if ((int64_t)smem > ctx->max_smem_optin)
        RBFFD_FAIL(ctx, RBFFD_ERR_UNSUPPORTED, "k=%d needs %zu B of shared memory per block (max %d)", k, smem, ctx->max_smem_optin);
    const int nb = ceil_div_i64(NQ, KNN_BS);
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<nb, KNN_BS, smem, st>>>(a);
        KLAUNCH(ctx);
        return cudaGetLastError();
    };
    if (B.dim == 1) CUDA_TRY(ctx, launch(knn_kernel<1>));
    else if (B.dim == 2) CUDA_TRY(ctx, launch(knn_kernel<2>));
    else CUDA_TRY(ctx, launch(knn_kernel<3>));
    if (short_rows_dev) return RBFFD_OK;
    int h_flag = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&h_flag, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (h_flag > 0)
        RBFFD_FAIL(ctx, RBFFD_ERR_K_TOO_LARGE, "%d queries see fewer than k=%d points", h_flag, k);
    return RBFFD_OK;
}

}  // namespace

int rbffd_knn_impl(rbffd_context* ctx, const double* X, int64_t N, int dim, const double* Q, int64_t NQ, int k,
                   const int32_t* xgroup, const int32_t* qgroup, bool q_is_x, int32_t* idx_out, double* d2_out) {
    if (dim < 1 || dim > 3) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "dim must be 1..3 (got %d)", dim);
    if (N < 1 || k < 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "need N >= 1 and k >= 1");
    if (N > 0x7fffffff - 1024 || NQ > 0x7fffffff / std::max(k, 1))
        RBFFD_FAIL(ctx, RBFFD_ERR_UNSUPPORTED, "node count exceeds the int32 device index range per shard");
    if (k > N) RBFFD_FAIL(ctx, RBFFD_ERR_K_TOO_LARGE, "k=%d exceeds the number of points %lld", k, (long long)N);
    Bins B;
    RBFFD_TRY(build_bins(ctx, X, N, dim, k, xgroup, B));
    return run_knn(ctx, B, q_is_x ? nullptr : Q, q_is_x ? N : NQ, qgroup, k, idx_out, d2_out);
}

int rbffd_stencils_impl(rbffd_context* ctx, const double* X, int64_t N, int dim, const double* Y, int64_t M, int n,
                        const int32_t* xgroup, int32_t* stencils, double* d2_x, int32_t* center, double* d2_y) {
    if (dim < 1 || dim > 3) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "dim must be 1..3 (got %d)", dim);
    if (N < 1 || n < 1) RBFFD_FAIL(ctx, RBFFD_ERR_INVALID, "need N >= 1 and n >= 1");
    if (N > 0x7fffffff - 1024 || N > 0x7fffffff / n || M > 0x7fffffff / n)
        RBFFD_FAIL(ctx, RBFFD_ERR_UNSUPPORTED, "node count exceeds the int32 device index range per shard");
    if (n > N) RBFFD_FAIL(ctx, RBFFD_ERR_K_TOO_LARGE, "n=%d exceeds the number of points %lld", n, (long long)N);
    cudaStream_t st = ctx->stream;
    Bins B;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], st));
    RBFFD_TRY(build_bins(ctx, X, N, dim, n, xgroup, B));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], st));
    // both searches report into one pair of status words, fetched once behind the second launch
    DevBuf<int> status;
    CUDA_TRY(ctx, status.alloc(2, st));
    CUDA_TRY(ctx, cudaMemsetAsync(status.p, 0, 2 * sizeof(int), st));
    if (stencils) RBFFD_TRY(run_knn(ctx, B, nullptr, N, nullptr, n, stencils, d2_x, status.p));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[2], st));
    if (center) {
        // every Y query sees all of X (calculateneighbors.jl:90-94): unmasked search
        Bins& Bu = B;
        bool saved = Bu.has_groups;
        Bu.has_groups = false;
        int rc = run_knn(ctx, Bu, Y, M, nullptr, 1, center, d2_y, status.p + 1);
        Bu.has_groups = saved;
        RBFFD_TRY(rc);
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], st));
    int h_status[2] = {0, 0};
    CUDA_TRY(ctx, cudaMemcpyAsync(h_status, status.p, sizeof(h_status), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (h_status[0] > 0) RBFFD_FAIL(ctx, RBFFD_ERR_K_TOO_LARGE, "%d queries see fewer than k=%d points", h_status[0], n);
    if (h_status[1] > 0) RBFFD_FAIL(ctx, RBFFD_ERR_K_TOO_LARGE, "%d queries see fewer than k=%d points", h_status[1], 1);
    float ms;
    for (int i = 0; i < 3; ++i) {
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]));
        ctx->timings[i] = ms;
    }
    return RBFFD_OK;
}
