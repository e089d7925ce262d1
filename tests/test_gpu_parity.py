"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI of librbffd.so
(via the ctypes mirror) and is compared with the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star):
  * neighbour indices / CSR pattern: bit-exact, ties by (d^2, index)
  * weights: |w_gpu - w_ref|_inf <= C * eps * cond_1(A_i) * |w_ref|_inf per row, C = 50   (~1e-10 at cond 1e5)
  * operator application: 1e-12 relative (same weights on both sides)
"""
import numpy as np
import pytest

import rbffd_b200 as rb

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps
C_TOL = 50.0


@pytest.fixture(scope="module")
def ctx():
    return rb.Context(0)


def _check_weights(vals, ref, cond_rows, names=None):
    for o in range(ref.shape[0]):
        err = np.abs(vals[o] - ref[o]).max(axis=1)
        scale = np.abs(ref[o]).max(axis=1)
        tol = C_TOL * EPS * cond_rows * scale + 1e-300
        bad = err > tol
        assert not bad.any(), (names[o] if names else o, int(bad.sum()), float((err / tol).max()))


# ------------------------------------------------------------------------------------------------ kNN
@pytest.mark.parametrize("d,g,k", [(2, 150, 30), (3, 22, 60), (2, 40, 1), (2, 64, 50)])
def test_knn_bit_exact_lattice(ctx, oracle, d, g, k):
    X = rb.nodes.jittered_lattice(d, g, seed=1)
    idx, idy, dx, dy = rb.calculateneighbors(X, None, k, ctx=ctx)
    ref, d2 = oracle.knn(X, X, k)
    assert np.array_equal(idx, ref)
    assert np.array_equal(dx, np.sqrt(d2))
    assert np.array_equal(idy[:, 0], np.arange(len(X)))
    assert np.array_equal(idx[:, 0], np.arange(len(X)))


def test_knn_exact_ties_on_regular_lattice(ctx, oracle):
    g = np.arange(20.0)
    X = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    X = X[np.random.default_rng(0).permutation(len(X))]
    idx, _, dx, _ = rb.calculateneighbors(X, None, 13, ctx=ctx)
    ref, d2 = oracle.knn(X, X, 13, brute=True)
    assert np.array_equal(idx, ref) and np.array_equal(dx, np.sqrt(d2))


def test_knn_tominec_two_sets(ctx, oracle, tominec):
    X, Y = tominec["X"], tominec["Y"]
    idx, idy, dx, dy = rb.calculateneighbors(X, Y, 20, ctx=ctx)
    ref, d2 = oracle.knn(X, X, 20)
    refy, d2y = oracle.knn(X, Y, 1)
    assert np.array_equal(idx, ref) and np.array_equal(idy, refy)
    assert np.array_equal(dy, np.sqrt(d2y))


def test_knn_clustered_and_degenerate(ctx, oracle):
    rng = np.random.default_rng(7)
    X = np.concatenate([rng.random((3000, 2)) ** 3, 0.5 + 1e-4 * rng.standard_normal((500, 2)), [[5.0, 5.0]]])
    idx, _, _, _ = rb.calculateneighbors(X, None, 17, ctx=ctx)
    assert np.array_equal(idx, oracle.knn(X, X, 17)[0])
    Xl = np.stack([np.linspace(0, 1, 400) ** 2, np.zeros(400)], 1)     # collinear: zero extent along y
    idx, _, _, _ = rb.calculateneighbors(Xl, None, 5, ctx=ctx)
    assert np.array_equal(idx, oracle.knn(Xl, Xl, 5)[0])
    Y = rng.random((300, 2)) * 3 - 1                                    # queries outside the bounding box of X
    _, idy, _, dy = rb.calculateneighbors(X, Y, 3, ctx=ctx)
    ry, d2y = oracle.knn(X, Y, 1)
    assert np.array_equal(idy, ry) and np.array_equal(dy, np.sqrt(d2y))


def test_knn_masked_groups(ctx, oracle):
    rng = np.random.default_rng(3)
    N = 4000
    X = rng.random((N, 2))
    idx_in, bc, gh = range(0, 3000), [range(3000, 3200), range(3200, 3500)], [range(3500, 3700), range(3700, 4000)]
    idx, idy, _, _ = rb.calculateneighbors(X, X[:100] + 0.001, 25, idx_in, bc, gh, ctx=ctx)
    ref, cy, _, _ = oracle.calculateneighbors(X, X[:100] + 0.001, 25, (0, 3000), [(3000, 3200), (3200, 3500)],
                                              [(3500, 3700), (3700, 4000)])
    assert np.array_equal(idx, ref) and np.array_equal(idy[:, 0], cy)


def test_knn_errors(ctx):
    X = np.random.default_rng(0).random((10, 2))
    with pytest.raises(rb.RbffdError) as e:
        rb.calculateneighbors(X, None, 11, ctx=ctx)           # ArgumentError from knn in the reference
    assert e.value.code == rb._lib.ERR_K_TOO_LARGE
    with pytest.raises(rb.RbffdError):
        rb.calculateneighbors(np.array([[0.0, np.nan]] * 5), None, 2, ctx=ctx)


# ------------------------------------------------------------------------------------------------ weights
CASES = [
    # d, g, p, deg, n, ops
    (2, 60, 5, 3, 30, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"]),          # config 2 shape, reference operator tuple
    (2, 40, 5, 3, 30, ["Lap"]),                                         # config 2, Laplacian from one RHS
    (2, 36, 5, 5, 42, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy", ("Dk", 0, 4), ("Dk", 1, 4)]),   # adv_diff_test.jl parameters
    (2, 36, 5, 4, 50, ["Dxx", "Dyy", ("Dk", 0, 4), ("Dk", 1, 4)]),      # config 3 shape
    (3, 14, 7, 3, 60, ["Lap", "Dx", "Dy", "Dz"]),                       # config 4 shape
    (3, 12, 7, 3, 60, ["E", "Dx", "Dy", "Dz", "Dxx", "Dyy", "Dzz", "Dxy", "Dxz", "Dyz"]),
    (2, 30, 3, 3, 20, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"]),          # poisson_test.jl parameters
    # hyperviscosity at the order the survey probed against sympy (p = 7, K = 6: hyperviscosity_operator.jl:26 with K < p)
    (2, 36, 7, 5, 42, [("Dk", 0, 6), ("Dk", 1, 6), "Dxx"]),             # q = 21: multi-warp null-space kernel, zero polynomial RHS
    (2, 30, 7, 6, 60, [("Dk", 0, 6), ("Dk", 1, 6), "Lap"]),             # q = 28: degree-6 monomials give the polynomial RHS 6!
    (3, 12, 7, 3, 60, [("Dk", 2, 4), ("Dk", 0, 4), "Lap"]),             # 3-D d^4/dz^4 and d^4/dx^4 (extension of :26 to d = 3)
    (3, 12, 7, 4, 70, [("Dk", 2, 4), ("Dk", 1, 6)]),                    # 3-D, q = 35, degree-4 monomials hit by d^4/dz^4
]


@pytest.mark.parametrize("d,g,p,deg,n,ops", CASES)
def test_weights_vs_oracle(ctx, oracle, d, g, p, deg, n, ops):
    X = rb.nodes.jittered_lattice(d, g, seed=2)
    colind, vals = rb.generate_raw(X, None, p, n, deg, ops, ctx=ctx)
    rcol, rvals, cond = oracle.generate_operator(X, X, p, n, deg, ops=ops, mode=0, want_cond=True)
    assert np.array_equal(colind, rcol)                                 # pattern bit-exact (stencil order)
    _check_weights(vals, rvals, cond, ops)


@pytest.mark.parametrize("d,g,p,deg,n,ops", [
    (2, 50, 5, 3, 30, ["Lap"]), (2, 50, 5, 3, 30, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy", ("Dk", 0, 4), ("Dk", 1, 4)]),
    (2, 40, 3, 3, 20, ["E", "Dx", "Dy"]), (2, 40, 7, 2, 33, ["Lap", "Dxy"]), (2, 40, 5, 4, 31, ["Dxx"]),
    (3, 12, 5, 2, 36, ["Lap", "Dx", "Dy", "Dz"]), (3, 12, 3, 1, 18, ["Dzz", "Dxz"]),
    # multi-warp kernel (weights_mw.cu): 48 < m <= 96
    (2, 30, 5, 3, 46, ["Lap", "Dx"]),                                                   # m = 56: MT 7, 2 warps
    (2, 36, 5, 5, 42, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy", ("Dk", 0, 4), ("Dk", 1, 4)]),   # m = 63: MT 8 (padded), 2 warps
    (2, 36, 5, 4, 50, ["Dxx", "Dyy", ("Dk", 0, 4), ("Dk", 1, 4)]),                     # m = 65: MT 9, 4 warps (config 3)
    (3, 14, 7, 3, 60, ["Lap", "Dx", "Dy", "Dz"]),                                       # m = 80: MT 10, 4 warps (config 4)
    (2, 30, 7, 5, 66, ["Lap"]),                                                         # m = 87: MT 11
    (3, 11, 5, 4, 58, ["Lap", "Dz"])])                                                  # m = 93: MT 12
def test_dmma_kernel_vs_generic_and_oracle(ctx, oracle, d, g, p, deg, n, ops):
    """kernel=2 forces the register/DMMA Gauss-Jordan kernel, kernel=1 the shared-memory LU kernel."""
    X = rb.nodes.jittered_lattice(d, g, seed=8)
    c2, v2 = rb.generate_raw(X, None, p, n, deg, ops, ctx=ctx, kernel=2)
    c1, v1 = rb.generate_raw(X, None, p, n, deg, ops, ctx=ctx, kernel=1)
    rcol, rvals, cond = oracle.generate_operator(X, X, p, n, deg, ops=ops, mode=0, want_cond=True)
    assert np.array_equal(c2, rcol) and np.array_equal(c1, rcol)
    _check_weights(v2, rvals, cond, ops)
    _check_weights(v1, rvals, cond, ops)


@pytest.mark.parametrize("d,g,p,deg,n,ops", [
    (2, 60, 5, 3, 30, ["Lap"]),                                            # config 2
    (2, 50, 5, 3, 30, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy", ("Dk", 0, 4), ("Dk", 1, 4)]),
    (2, 40, 3, 3, 20, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"]),             # poisson_test.jl parameters (collocated)
    (2, 40, 7, 3, 32, ["Lap", "Dxy"]), (2, 40, 5, 2, 18, ["Dxx"]), (2, 40, 3, 1, 9, ["Lap", "Dy"]),
    (3, 12, 5, 2, 32, ["Lap", "Dx", "Dy", "Dz"]), (3, 12, 3, 1, 12, ["Dzz", "Dxz"])])
def test_nullspace_kernel_vs_oracle(ctx, oracle, d, g, p, deg, n, ops):
    """kernel=3 forces the null-space kernel (weights_ns.cu): column reduction of P, DMMA-formed Z'Phi Z, elimination
    without pivoting.  Same tolerance as every other weight kernel."""
    X = rb.nodes.jittered_lattice(d, g, seed=8)
    c3, v3 = rb.generate_raw(X, None, p, n, deg, ops, ctx=ctx, kernel=3)
    rcol, rvals, cond = oracle.generate_operator(X, X, p, n, deg, ops=ops, mode=0, want_cond=True)
    assert np.array_equal(c3, rcol)
    _check_weights(v3, rvals, cond, ops)


@pytest.mark.parametrize("d,g,p,deg,n,ops", [
    (2, 36, 5, 5, 42, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy", ("Dk", 0, 4), ("Dk", 1, 4)]),   # config 1 (adv_diff_test.jl): q=21, nb=21
    (2, 36, 5, 4, 50, ["Dxx", "Dyy", ("Dk", 0, 4), ("Dk", 1, 4)]),                     # config 3: q=15, nb=35
    (3, 14, 7, 3, 60, ["Lap", "Dx", "Dy", "Dz"]),                                       # configs 4/5: q=20, nb=40
    (2, 30, 5, 4, 62, ["Lap", "Dx"]),                                                   # nb=47: right-hand sides in a 7th tile column
    (2, 30, 5, 3, 46, ["Lap", "Dx"]), (3, 12, 5, 2, 40, ["Lap", "Dz"]),                 # q=10 with n > 32
    (3, 12, 7, 3, 33, ["E", "Dx", "Dy", "Dz", "Dxx", "Dyy", "Dzz", "Dxy"])])            # small null space (nb=13), 8 operators
def test_multiwarp_nullspace_kernel_vs_oracle(ctx, oracle, d, g, p, deg, n, ops):
    """kernel=3 with n > 32 runs the multi-warp null-space kernel (weights_nsw.cu).  Same tolerance as every other kernel."""
    X = rb.nodes.jittered_lattice(d, g, seed=8)
    c3, v3 = rb.generate_raw(X, None, p, n, deg, ops, ctx=ctx, kernel=3)
    rcol, rvals, cond = oracle.generate_operator(X, X, p, n, deg, ops=ops, mode=0, want_cond=True)
    assert np.array_equal(c3, rcol)
    _check_weights(v3, rvals, cond, ops)


@pytest.mark.parametrize("d,N,p,deg,n,ops", [
    (2, 1500, 5, 4, 50, ["E", "Dx", "Dy", "Dxy"]),                          # n = 50 / four operators: the instance compiled for config 3
    (2, 1500, 3, 4, 50, ["Lap", ("Dk", 0, 2), "Dyy", "E"]),                 # same instance, r^3
    (3, 2500, 7, 3, 60, ["E", "Dzz", "Dxz", "Dy"]),                         # n = 60 / four operators: the instance compiled for configs 4/5
    (3, 2500, 5, 3, 60, ["Lap", "Dxy", "Dz", "Dyy"])])                      # same instance, r^5
def test_specialised_two_stage_instances_other_operators_scattered_nodes(ctx, oracle, d, N, p, deg, n, ops):
    """The two-stage null-space kernels compiled for the BASELINE shapes (n = 50 or 60, four operators) are selected by (n, number of
    operators) alone: other operator sets and PHS powers, on SCATTERED (uniformly random, locally clustered) nodes instead of a
    jittered lattice -- in particular the symmetric form of S = Z' Phi Z (rows of the basic nodes of Y only, S = V + V')."""
    rng = np.random.default_rng(17)
    X = rng.random((N, d))
    X[: N // 5] = 0.5 + 0.05 * rng.standard_normal((N // 5, d))             # a cluster: strongly varying stencil diameters
    colind, vals = rb.generate_raw(X, None, p, n, deg, ops, ctx=ctx)
    rcol, rvals, cond = oracle.generate_operator(X, X, p, n, deg, ops=ops, mode=0, want_cond=True)
    assert np.array_equal(colind, rcol)
    _check_weights(vals, rvals, cond, ops)
    # row sums: E reproduces constants, derivatives annihilate them (polynomial reproduction of degree 0)
    for o, nm in enumerate(ops):
        rs = vals[o].sum(axis=1)
        scale = np.abs(vals[o]).max(axis=1) * cond * EPS * C_TOL
        assert np.all(np.abs(rs - (1.0 if nm == "E" else 0.0)) <= scale + 1e-12)


@pytest.mark.parametrize("d,g,p,deg,n,ops,over", [
    (2, 40, 3, 3, 20, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"], 3),          # poisson_test.jl:56 (M ~ 3N), single-warp kernel
    (2, 30, 5, 5, 42, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"], 2),          # mesh_import_test.jl parameters, multi-warp kernel
    (3, 12, 7, 3, 60, ["Lap", "Dx", "Dz", "E"], 2)])                        # 3-D oversampled rows (n + nops <= 64)
def test_oversampled_rows_through_nullspace_kernels(ctx, oracle, d, g, p, deg, n, ops, over):
    """Y != X (generate_operator.jl:89-167) with kernel=3: one work item per Y row, reading the stencil of its nearest X
    node and evaluating every right-hand side at eta = (Y_k - X_c) s != 0.  Checked against the oracle (same tolerance as
    the collocated rows) and against the generic kernel that shares one factorisation per centre (kernel=1)."""
    X = rb.nodes.jittered_lattice(d, g, seed=8)
    rng = np.random.default_rng(3)
    Y = rng.uniform(0.02, 0.98, size=(over * len(X), d))
    Y[::7] = X[rng.integers(0, len(X), size=len(Y[::7]))]                   # some rows sit exactly on a node (eta == 0)
    c3, v3 = rb.generate_raw(X, Y, p, n, deg, ops, ctx=ctx, kernel=3)
    c1, v1 = rb.generate_raw(X, Y, p, n, deg, ops, ctx=ctx, kernel=1)
    c0, v0 = rb.generate_raw(X, Y, p, n, deg, ops, ctx=ctx)
    rcol, rvals, cond = oracle.generate_operator(X, Y, p, n, deg, ops=ops, mode=0, want_cond=True)
    center = oracle.knn(X, Y, 1)[0][:, 0]
    assert np.array_equal(c3, rcol) and np.array_equal(c1, rcol) and np.array_equal(c0, rcol)
    _check_weights(v3, rvals, cond[center], ops)
    _check_weights(v1, rvals, cond[center], ops)
    assert np.array_equal(v0, v3)                                           # the automatic dispatch takes the null-space path


@pytest.mark.parametrize("d,g,p,deg,n,ops,over", [
    (2, 40, 5, 3, 30, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"], 3),          # config-2 stencil, reference tuple: 5 tile columns
    (2, 40, 3, 3, 20, ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"], 3),          # poisson_test.jl:52-56: 4 tile columns
    (2, 40, 5, 3, 30, ["Lap"], 4),                                           # one operator: 3 tile columns, centres with > 3 rows
    (2, 30, 3, 2, 14, ["Lap", "Dx"], 2),                                     # degree 2 (q = 6)
    (3, 10, 5, 2, 30, ["Lap", "Dx", "Dy", "Dz"], 2)])                        # 3-D, degree 2 (q = 10)
def test_segmented_rows_share_the_elimination(ctx, oracle, d, g, p, deg, n, ops, over):
    """Y != X on the single-warp kernel: by default the rows of one centre ride as right-hand-side columns of ONE elimination
    (generate_operator.jl:89-95,158: they share inv(A)); kernel=4 solves row by row.  Same pattern, weights equal to rounding
    of the solve, and both within the oracle tolerance.  Clustered rows give centres with many rows, others get none."""
    X = rb.nodes.jittered_lattice(d, g, seed=5)
    rng = np.random.default_rng(11)
    Y = rng.uniform(0.02, 0.98, size=(over * len(X), d))
    Y[: len(X) // 2] = 0.3 + 0.1 * rng.uniform(size=(len(X) // 2, d))      # a cluster: up to dozens of rows per centre
    Y[::9] = X[rng.integers(0, len(X), size=len(Y[::9]))]                   # rows exactly on a node (eta == 0)
    cs, vs = rb.generate_raw(X, Y, p, n, deg, ops, ctx=ctx, kernel=3)
    cr, vr = rb.generate_raw(X, Y, p, n, deg, ops, ctx=ctx, kernel=4)
    rcol, rvals, cond = oracle.generate_operator(X, Y, p, n, deg, ops=ops, mode=0, want_cond=True)
    center = oracle.knn(X, Y, 1)[0][:, 0]
    assert np.array_equal(cs, rcol) and np.array_equal(cr, rcol)
    _check_weights(vs, rvals, cond[center], ops)
    _check_weights(vr, rvals, cond[center], ops)
    _check_weights(vs, vr, cond[center], ops)                               # segmented vs row by row, row-wise tolerance


def test_nullspace_kernel_falls_back_when_not_definite(ctx, oracle):
    """polydeg < (p-1)/2: Z'Phi Z is not definite, kernel=3 refuses and the automatic dispatch uses the pivoted kernels."""
    X = rb.nodes.jittered_lattice(2, 30, seed=8)
    with pytest.raises(rb.RbffdError) as e:
        rb.generate_raw(X, None, 7, 24, 2, ["Lap"], ctx=ctx, kernel=3)
    assert e.value.code == rb._lib.ERR_UNSUPPORTED
    c0, v0 = rb.generate_raw(X, None, 7, 24, 2, ["Lap"], ctx=ctx)
    rcol, rvals, cond = oracle.generate_operator(X, X, 7, 24, 2, ops=["Lap"], mode=0, want_cond=True)
    _check_weights(v0, rvals, cond)


def test_dmma_kernel_scope(ctx):
    X = rb.nodes.jittered_lattice(2, 30, seed=1)
    with pytest.raises(rb.RbffdError) as e:
        rb.generate_raw(X, None, 7, 80, 5, ["Lap"], ctx=ctx, kernel=2)      # m = 101 > 96: generic kernel only
    assert e.value.code == rb._lib.ERR_UNSUPPORTED
    Xd = X.copy()
    Xd[5] = Xd[6]
    with pytest.raises(rb.RbffdError) as e:
        rb.generate_raw(Xd, None, 3, 20, 3, ["Lap"], ctx=ctx, kernel=2)      # duplicate node: rows are no longer collocated
    assert e.value.code in (rb._lib.ERR_SINGULAR, rb._lib.ERR_UNSUPPORTED)
    with pytest.raises(rb.RbffdError) as e:
        rb.generate_raw(Xd, None, 3, 20, 3, ["Lap"], ctx=ctx)
    assert e.value.code == rb._lib.ERR_SINGULAR


def test_sorted_pattern_bit_exact(ctx, oracle):
    X = rb.nodes.jittered_lattice(2, 50, seed=9)
    colind, vals = rb.generate_raw(X, None, 5, 30, 3, ["Dxx", "Dyy"], ctx=ctx, sort_columns=True)
    rcol, rvals, cond = oracle.generate_operator(X, X, 5, 30, 3, ops=["Dxx", "Dyy"], want_cond=True)
    scol, svals = oracle.sort_rows(rcol, list(rvals))
    assert np.array_equal(colind, scol)
    _check_weights(vals, np.stack(svals), cond)


def test_sorted_one_based_pattern_is_the_reference_sparse_matrix(ctx, oracle, tominec):
    """What the Julia shim hands to SparseMatrixCSC: sort_columns = 1 and index_base = 1 (julia/RBFFDB200.jl).  The fixed-row
    CSR (rowptr = 1 + k n, 1-based sorted column ids, values) read as the CSC of the TRANSPOSE must be, entry for entry, the
    matrix `sparse(vec(idx_rows), vec(idx_columns), vec(W))` builds (generate_operator.jl:171-182): same size (max I, max J),
    same stored entries (zeros kept), strictly increasing row ids inside every column of the transpose."""
    import scipy.sparse as sp
    X, Y = tominec["X"], tominec["Y"]
    p, n, deg = 3, 30, 4
    ops = ["E", "Dx", "Dyy"]
    c1, v1 = rb.generate_raw(X, Y, p, n, deg, ops, ctx=ctx, sort_columns=True, index_base=1)
    c0, v0 = rb.generate_raw(X, Y, p, n, deg, ops, ctx=ctx)                      # stencil order, 0-based
    M = len(Y)
    assert c1.min() >= 1 and c1.max() <= len(X)
    assert np.all(np.diff(c1, axis=1) > 0)                                         # sorted, no duplicates (kNN ids are distinct)
    rowptr1 = 1 + n * np.arange(M + 1)                                             # what the shim passes as colptr
    rcol, rvals = oracle.generate_operator(X, Y, p, n, deg, ops=ops)
    assert np.array_equal(c0, rcol)
    for o in range(len(ops)):
        # Julia: SparseMatrixCSC(N, M, colptr, rowval, nzval) is D' ; here: the same three arrays, 1-based -> 0-based
        Dt = sp.csc_matrix((v1[o].ravel(), c1.ravel() - 1, rowptr1 - 1), shape=(int(c1.max()), M))
        assert Dt.has_sorted_indices
        # the reference's COO -> CSC assembly from the unsorted stencil-order triplets of the GPU path
        rows = np.repeat(np.arange(M), n)
        ref = sp.coo_matrix((v0[o].ravel(), (rows, c0.ravel())), shape=(M, int(c0.max()) + 1)).tocsc()
        D = Dt.T.tocsc()
        assert D.shape == ref.shape and D.nnz == ref.nnz == M * n
        assert np.array_equal(D.indptr, ref.indptr) and np.array_equal(D.indices, ref.indices)
        assert np.array_equal(D.data, ref.data)                                    # sorting moves values, never changes them


def test_two_set_oversampled_poisson(ctx, oracle, tominec):
    """Y != X (generate_operator.jl:89-167): rows share the factorisation of their nearest X node."""
    from poisson_helper import poisson_error as _poisson_error
    X, Y = tominec["X"], tominec["Y"]
    colind, vals = rb.generate_raw(X, Y, 3, 20, 3, ctx=ctx)
    rcol, rvals, cond = oracle.generate_operator(X, Y, 3, 20, 3, want_cond=True)
    assert np.array_equal(colind, rcol)
    center = oracle.knn(X, Y, 1)[0][:, 0]
    _check_weights(vals, rvals, cond[center], rb.REFERENCE_OPS)
    # the reference's own assertion (test/poisson_test.jl:132) through the GPU path
    err = _poisson_error(tominec, lambda X, Y, p, n, deg: rb.generate_raw(X, Y, p, n, deg, ctx=ctx))
    assert err < float(tominec["poisson_threshold"])
    assert abs(err - 0.0026579) < 2e-6


def test_poisson_reference_test_on_device(tominec):
    """test/poisson_test.jl end to end on the device: operators in HBM, row-block collocation matrix, the least-squares solve
    `u = D \\ f` (:121, sparse QR in the reference) by preconditioned CGLS over SpMV / transposed SpMV, `E*u` by SpMV.  Same
    known answer as the host replay: rel. l2 error 0.0026579 < 0.0027 (:132)."""
    import torch
    from poisson_helper import poisson_error_device
    c = rb.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    err, iters, rel = poisson_error_device(tominec, c)
    assert rel <= 1e-11 and iters < 50000
    assert err < float(tominec["poisson_threshold"])
    assert abs(err - 0.0026579) < 2e-6


def test_mesh_import_reference_test_on_device(tominec):
    """test/mesh_import_test.jl on the device (n = 42, polydeg = 5: the oversampled rows go through the multi-warp null-space
    kernel, the least-squares solve through CGLS): err < 0.001 (:158), the host replay gives 4.70e-4."""
    import os
    import torch
    from poisson_helper import mesh_import_error_device
    c = rb.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    path = os.path.join(os.path.dirname(__file__), "golden", "tominec_Y.cgns")
    err, iters, rel = mesh_import_error_device(path, tominec["X"], c)
    assert rel <= 1e-11 and iters < 100000
    assert err < 0.001 and abs(err - 4.70e-4) < 2e-6


def test_mesh_import_reference_test(ctx, tominec):
    """test/mesh_import_test.jl:158 (err < 0.001) through the GPU path: CGNS mesh -> processmesh (exact GPU 1-NN for the
    ghost offset and the normal orientation) -> generate_operator on the device -> the reference's least-squares solve."""
    import os
    from poisson_helper import mesh_import_error
    path = os.path.join(os.path.dirname(__file__), "golden", "tominec_Y.cgns")
    err = mesh_import_error(path, tominec["X"], lambda X, Y, p, n, deg: rb.generate_raw(X, Y, p, n, deg, ctx=ctx), ctx=ctx)
    assert err < 0.001


def test_hyperviscosity_reference_test(ctx, tominec):
    """test/hyperviscosity_test.jl:13-33 through the mirrored API (isapprox on sparse = Frobenius, rtol sqrt(eps))."""
    import scipy.sparse.linalg as spl
    X = tominec["X"]
    E, Dx, Dy, Dxx, Dyy, Dxy = rb.generate_operator(X, X, 5, 42, 5, ctx=ctx)
    Dxk, Dyk = rb.hyperviscosity_operator(2, X, X, 5, 42, 5, ctx=ctx)
    rtol = float(tominec["hyperviscosity_rtol"])
    for a, b in ((Dxk, Dxx), (Dyk, Dyy)):
        assert a.shape == b.shape == (2000, 2000)
        assert spl.norm(a - b) <= rtol * max(spl.norm(a), spl.norm(b))
        assert spl.norm(a - b) <= 1e-10 * spl.norm(b)


def test_boundary_aware_generate(ctx, oracle):
    rng = np.random.default_rng(5)
    X = rb.nodes.jittered_lattice(2, 40, seed=4)
    N = len(X)
    perm = rng.permutation(N)
    X = X[perm]
    sets = (range(0, 1200), [range(1200, 1300), range(1300, 1400)], [range(1400, 1500), range(1500, 1600)])
    ops = rb.generate_operator(X, X, 3, 20, 3, *sets, None, None, None, ctx=ctx, shape="full")
    rcol, rvals = oracle.generate_operator(X, X, 3, 20, 3, groups=((0, 1200), [(1200, 1300), (1300, 1400)],
                                                                   [(1400, 1500), (1500, 1600)]))
    import scipy.sparse as sp
    for o in range(6):
        ref = sp.csr_matrix((rvals[o].ravel(), rcol.ravel(), np.arange(0, N * 20 + 1, 20)), shape=(N, N))
        diff = abs(ops[o] - ref.tocsc())
        assert diff.max() <= 1e-7 * abs(ref).max()
        assert (ops[o] != 0).nnz <= ref.nnz


def test_errors(ctx):
    X = rb.nodes.jittered_lattice(2, 12)
    with pytest.raises(rb.RbffdError) as e:
        rb.generate_raw(X, None, 4, 20, 3, ctx=ctx)                      # even PHS power
    assert e.value.code == rb._lib.ERR_UNSUPPORTED
    with pytest.raises(rb.RbffdError) as e:
        rb.generate_raw(X, None, 3, 8, 3, ctx=ctx)                       # n < number of polynomial terms
    assert e.value.code == rb._lib.ERR_SINGULAR
    Xd = X.copy()
    Xd[5] = Xd[6]                                                        # duplicate node -> singular A (SingularException)
    with pytest.raises(rb.RbffdError) as e:
        rb.generate_raw(Xd, None, 3, 20, 3, ctx=ctx)
    assert e.value.code == rb._lib.ERR_SINGULAR


# ------------------------------------------------------------------------------------------------ application
def test_spmv_and_rhs(ctx, oracle):
    X = rb.nodes.jittered_lattice(2, 70, seed=6)
    N = len(X)
    names = ["E", "Dx", "Dy", "Dxx", "Dyy", ("Dk", 0, 4), ("Dk", 1, 4)]
    colind, vals = rb.generate_raw(X, None, 5, 42, 5, names, ctx=ctx)
    op = rb.Operator.from_host(ctx, colind, vals, N)
    rng = np.random.default_rng(0)
    u = rng.standard_normal(N)
    for w in range(7):
        y = op.spmv(w, u)
        ref = oracle.spmv(colind, vals[w], u)
        assert np.max(np.abs(y - ref)) <= 1e-12 * np.max(np.abs(ref))
    y0 = rng.standard_normal(N)
    y = op.spmv(3, u, alpha=0.5, beta=-2.0, y=y0.copy())
    ref = oracle.spmv(colind, vals[3], u, 0.5, -2.0, y0.copy())
    assert np.max(np.abs(y - ref)) <= 1e-12 * np.max(np.abs(ref))
    yt = op.spmv_t(1, u)
    reft = oracle.spmv_t(colind, vals[1], u, N)
    assert np.max(np.abs(yt - reft)) <= 1e-12 * np.max(np.abs(reft))
    prm = rb.AdvDiffParams(iE=0, iDx=1, iDy=2, iDxx=3, iDyy=4, iDxk=5, iDyk=6, alpha=1.0, ux=0.5, uy=-0.25, gamma=100 * (1 / 70) ** 4)
    du = op.rhs_advdiff(u, prm)
    ref = oracle.rhs_advdiff(colind, *vals, 1.0, 0.5, -0.25, prm.gamma, u)
    assert np.max(np.abs(du - ref)) <= 1e-12 * np.max(np.abs(ref))
    prm0 = rb.AdvDiffParams(iE=0, iDx=1, iDy=2, iDxx=3, iDyy=4, iDxk=5, iDyk=6, alpha=1.0, ux=0.0, uy=0.0, gamma=prm.gamma)
    du0 = op.rhs_advdiff(u, prm0)                                       # the shipped example: u_x = u_y = 0 (adv_diff_test.jl:89)
    ref0 = oracle.rhs_advdiff(colind, *vals, 1.0, 0.0, 0.0, prm.gamma, u)
    assert np.max(np.abs(du0 - ref0)) <= 1e-12 * np.max(np.abs(ref0))
    with pytest.raises(ValueError):
        op.spmv(0, u[:-1])


def test_fused_cons_sys_and_stage_kernels(ctx, oracle):
    """The collocated form of cons_sys (E = I checked on the device: all six operators in ONE pass over the shared pattern)
    against the oracle's three-product form (adv_diff_test.jl:151-152): the two differ by the rounding noise of E's weights
    (|E - I| <= eps * cond, here < 1e-9), the general path agrees to 1e-12; and the SSP-RK stage entry points
    (rbffd_rhs_advdiff_stage_device, rbffd_spmv_stage_device, rbffd_stage_update_device) against NumPy on the same du."""
    import torch
    X = rb.nodes.jittered_lattice(2, 70, seed=6)
    N = len(X)
    names = ["E", "Dx", "Dy", "Dxx", "Dyy", ("Dk", 0, 4), ("Dk", 1, 4)]
    colind, vals = rb.generate_raw(X, None, 5, 42, 5, names, ctx=ctx)
    op = rb.Operator.from_host(ctx, colind, vals, N)
    rng = np.random.default_rng(3)
    x, u = rng.standard_normal(N), rng.standard_normal(N)
    gamma = 100 * (1 / 70) ** 4
    ref = oracle.rhs_advdiff(colind, *vals, 1.0, 0.5, -0.25, gamma, x)
    scale = np.max(np.abs(ref))
    general = rb.AdvDiffParams(iE=0, iDx=1, iDy=2, iDxx=3, iDyy=4, iDxk=5, iDyk=6, alpha=1.0, ux=0.5, uy=-0.25, gamma=gamma)
    fused = rb.AdvDiffParams(iE=0, iDx=1, iDy=2, iDxx=3, iDyy=4, iDxk=5, iDyk=6, alpha=1.0, ux=0.5, uy=-0.25, gamma=gamma, flags=rb.ADVDIFF_COLLOCATED)
    l0 = ctx.launch_count()
    du_f = op.rhs_advdiff(x, fused)
    first = ctx.launch_count() - l0                       # identity check (1 launch, once) + ONE product
    l0 = ctx.launch_count()
    du_f2 = op.rhs_advdiff(x, fused)
    assert ctx.launch_count() - l0 == 1 and first == 2
    assert np.array_equal(du_f, du_f2)
    E_defect = np.abs(vals[0] - (colind == np.arange(N)[:, None])).max()
    assert E_defect < 1e-9
    assert np.max(np.abs(du_f - ref)) <= 4 * E_defect * scale
    # ... and it IS the plain sum of the six products (no E anywhere)
    six = sum(c * oracle.spmv(colind, vals[w], x) for w, c in zip([3, 4, 1, 2, 5, 6], [1.0, 1.0, -0.5, 0.25, -gamma, -gamma]))
    assert np.max(np.abs(du_f - six)) <= 1e-12 * scale
    du_g = op.rhs_advdiff(x, general)
    assert np.max(np.abs(du_g - ref)) <= 1e-12 * scale
    # an operator whose "E" is not the identity must keep the three-product form even when the flag is set
    wrong = rb.AdvDiffParams(iE=1, iDx=1, iDy=2, iDxx=3, iDyy=4, iDxk=5, iDyk=6, alpha=1.0, ux=0.5, uy=-0.25, gamma=gamma, flags=rb.ADVDIFF_COLLOCATED)
    vv = list(vals)
    vv[0] = vals[1]
    ref_w = oracle.rhs_advdiff(colind, *vv, 1.0, 0.5, -0.25, gamma, x)
    op2 = rb.Operator.from_host(ctx, colind, vals, N)
    assert np.max(np.abs(op2.rhs_advdiff(x, wrong) - ref_w)) <= 1e-12 * np.max(np.abs(ref_w))
    # stage entry points
    xd, ud = torch.from_numpy(x).cuda(), torch.from_numpy(u).cuda()
    out = torch.empty_like(xd)
    a, b, dt = 0.75, 0.25, 3e-4
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    op.rhs_advdiff_stage_device(xd.data_ptr(), a, ud.data_ptr(), b, dt, out.data_ptr(), fused)
    torch.cuda.synchronize()
    want = a * u + b * (x + dt * du_f)
    assert np.max(np.abs(out.cpu().numpy() - want)) <= 4 * EPS * np.max(np.abs(want))
    op.rhs_advdiff_stage_device(xd.data_ptr(), a, ud.data_ptr(), b, dt, out.data_ptr(), general)
    torch.cuda.synchronize()
    want = a * u + b * (x + dt * du_g)
    assert np.max(np.abs(out.cpu().numpy() - want)) <= 4 * EPS * np.max(np.abs(want))
    which, coef = [3, 4, 1, 2, 5, 6], [1.0, 1.0, -0.5, 0.25, -gamma, -gamma]
    op.spmv_stage_device(which, coef, xd.data_ptr(), a, ud.data_ptr(), b, dt, out.data_ptr())
    yd = torch.empty_like(xd)
    op.spmv_multi_device(which, coef, xd.data_ptr(), yd.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(yd.cpu().numpy(), du_f)         # six terms in one launch = the collocated cons_sys line
    want = a * u + b * (x + dt * yd.cpu().numpy())
    assert np.max(np.abs(out.cpu().numpy() - want)) <= 4 * EPS * np.max(np.abs(want))
    dud = torch.from_numpy(du_f).cuda()
    ctx.stage_update_device(N, a, ud.data_ptr(), b, xd.data_ptr(), dt, dud.data_ptr(), ud.data_ptr())       # in place on u
    torch.cuda.synchronize()
    assert np.max(np.abs(ud.cpu().numpy() - (a * u + b * (x + dt * du_f)))) <= 4 * EPS * np.max(np.abs(want))
    with pytest.raises(rb.RbffdError):
        op.spmv_stage_device(which, coef, xd.data_ptr(), a, ud.data_ptr(), b, dt, xd.data_ptr())            # out aliases x
    ctx.reset_stream()


def test_device_resident_pipeline_and_lattice(ctx, oracle):
    import torch
    dev = torch.device("cuda:0")
    g = 300
    N = g * g
    Xd = torch.empty((N, 2), dtype=torch.float64, device=dev)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.jittered_lattice_device(2, g, 0, 0, N, Xd.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(Xd.cpu().numpy(), rb.nodes.jittered_lattice(2, g, 0))
    opts = rb.make_options(2, 5, 30, 3, ["Lap"])
    op = ctx.operator_generate(opts, Xd.data_ptr(), N)
    x = torch.randn(N, dtype=torch.float64, device=dev)
    y = torch.empty(N, dtype=torch.float64, device=dev)
    op.spmv_device(0, x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    colind, vals = op.to_host()
    ref = oracle.spmv(colind, vals[0], x.cpu().numpy())
    assert np.max(np.abs(y.cpu().numpy() - ref)) <= 1e-12 * np.max(np.abs(ref))
    # size-independent properties at a size the oracle does not replay in full: Laplacian of a quadratic is exact
    Xh = Xd.cpu().numpy()
    f = torch.from_numpy(Xh[:, 0] ** 2 + 3 * Xh[:, 0] * Xh[:, 1] - Xh[:, 1] ** 2 + 2).to(dev)
    op.spmv_device(0, f.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    assert float(y.abs().max()) < 1e-6        # Lap = 2 - 2 = 0, cond*eps*|w| ~ 1e5*1e-16*1e6
    sub = np.random.default_rng(0).choice(N, 2000, replace=False)
    rcol = oracle.knn(Xh, Xh[sub], 30)[0]
    assert np.array_equal(colind[sub], rcol)


def test_ghost_node_boundary_updates(ctx):
    """K7: the ghost updates of cons_sys (examples/adv_diff_test.jl:118-141, 162-176) against a NumPy restatement
    of those lines (sparse slicing + dense inv), on a synthetic rectangle with boundary and ghost node sets."""
    import scipy.sparse as sp
    g = 40
    h = 1.0 / g
    X_in = rb.nodes.jittered_lattice(2, g, seed=12)
    t = (np.arange(g) + 0.5) * h
    sides = [(np.stack([np.zeros(g), t], 1), (-1, 0)), (np.stack([np.ones(g), t], 1), (1, 0)),
             (np.stack([t, np.ones(g)], 1), (0, 1)), (np.stack([t, np.zeros(g)], 1), (0, -1))]      # left, right, top, bottom
    bc_pts = [s[0] for s in sides]
    gh_pts = [s[0] + 0.7 * h * np.array(s[1]) for s in sides]
    X = np.concatenate([X_in] + bc_pts + gh_pts)
    N, n_in = len(X), len(X_in)
    idx_bc = [range(n_in + b * g, n_in + (b + 1) * g) for b in range(4)]
    idx_g = [range(n_in + 4 * g + b * g, n_in + 4 * g + (b + 1) * g) for b in range(4)]
    E, Dx, Dy, Dxx, Dyy, Dxy = rb.generate_operator(X, X, 3, 20, 3, range(0, n_in), idx_bc, idx_g, None, None, None, ctx=ctx, shape="full")
    colind, vals = rb.generate_raw(X, None, 3, 20, 3, ["Dx", "Dy"], groups=rb.groups_from_index_sets(N, range(0, n_in), idx_bc, idx_g), ctx=ctx)
    op = rb.Operator.from_host(ctx, colind, vals, N)
    # application order of the reference: right (Dx), left (Dirichlet 1.0), top (Dy), bottom (Dy)
    order = [(1, 0), (0, None), (2, 1), (3, 1)]
    bcs = rb.BoundaryConditions(op, [{"bc": list(idx_bc[b]), "ghost": list(idx_g[b]), **({"matrix": m} if m is not None else {"value": 1.0})}
                                     for b, m in order])
    rng = np.random.default_rng(3)
    u0 = rng.standard_normal(N)
    got = bcs.apply(u0.copy())
    # NumPy restatement of adv_diff_test.jl:118-141,162-176
    ref = u0.copy()
    D = {0: sp.csr_matrix(Dx), 1: sp.csr_matrix(Dy)}
    for b, m in order:
        bc, gh = np.array(idx_bc[b]), np.array(idx_g[b])
        if m is None:
            ref[bc] = 1.0
            ref[gh] = 1.0
            continue
        w_g = D[m][bc][:, gh].toarray()
        rest = np.setdiff1d(np.arange(N), gh)
        w_int = D[m][bc][:, rest]
        ref[gh] = -np.linalg.inv(w_g) @ (w_int @ ref[rest])
    assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref))
    # the constraint the ghost values enforce: (D_b u)[bc_b] = 0 on the last-applied boundary
    assert np.max(np.abs((D[1] @ got)[np.array(idx_bc[3])])) <= 1e-8 * np.abs(D[1]).max()


@pytest.mark.parametrize("mesh", [None, "rect_0_10.cgns"])
def test_device_resident_time_stepping_example(oracle, mesh):
    """examples/adv_diff_b200.py (generate -> RHS -> ghost updates -> SSP-RK3, all on the device) against the same
    scheme driven by the CPU oracle's operators and a NumPy restatement of the ghost updates.  mesh = rect_0_10.cgns is
    BASELINE config 1 (examples/adv_diff_test.jl on the coarsest mesh, node set through rb.mesh.processmesh)."""
    import importlib.util, os
    import scipy.sparse as sp
    spec = importlib.util.spec_from_file_location("adv_diff_b200", os.path.join(os.path.dirname(__file__), "..", "examples", "adv_diff_b200.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gy, steps = 16, 12
    if mesh:
        path = os.path.join(os.path.dirname(__file__), "golden", mesh)
        X, u, (idx_in, idx_bc, idx_g) = mod.run(steps=steps, verbose=False, mesh=path)
        h = mod.mesh_nodes(path)[4]
        assert len(X) == 1812 and len(idx_in) == 1572                     # 1572 centroids + 120 BC + 120 ghosts
    else:
        X, u, (idx_in, idx_bc, idx_g) = mod.run(gy=gy, steps=steps, verbose=False)
        Xg, ug, _ = mod.run(gy=gy, steps=steps, verbose=False, collocated=False)      # E' kept as a product (the reference's form)
        assert np.max(np.abs(u - ug)) <= 1e-10 * np.max(np.abs(ug))
        h = 1.0 / gy
    N, n = len(X), 42
    groups = ((idx_in.start, idx_in.stop), [(r.start, r.stop) for r in idx_bc], [(r.start, r.stop) for r in idx_g])
    names = ["E", "Dx", "Dy", "Dxx", "Dyy", ("Dk", 0, 4), ("Dk", 1, 4)]
    colind, vals = oracle.generate_operator(X, X, 5, n, 5, groups=groups, ops=names, mode=0)
    D = [sp.csr_matrix((v.ravel(), colind.ravel(), np.arange(0, N * n + 1, n)), shape=(N, N)) for v in vals]
    gamma, dt = 100.0 * h ** 4, 0.0025 * h * h

    def cons_sys(uu):                                  # adv_diff_test.jl:144-188 (du from u, then the ghost updates mutate u)
        du = oracle.rhs_advdiff(colind, *vals, 1.0, 0.0, 0.0, gamma, uu)
        for b, m in [(1, 1), (0, None), (2, 2), (3, 2)]:
            bc, gh = np.array(idx_bc[b]), np.array(idx_g[b])
            if m is None:
                uu[bc] = 1.0
                uu[gh] = 1.0
                continue
            rest = np.setdiff1d(np.arange(N), gh)
            uu[gh] = -np.linalg.inv(D[m][bc][:, gh].toarray()) @ (D[m][bc][:, rest] @ uu[rest])
        return du

    ur = np.where((X[:, 0] - 0.5) ** 2 + (X[:, 1] - 0.5) ** 2 <= 0.04, 10.0, 1.0)
    for _ in range(steps):
        du = cons_sys(ur)
        u1 = ur + dt * du
        du = cons_sys(u1)
        u2 = 0.75 * ur + 0.25 * (u1 + dt * du)
        du = cons_sys(u2)
        ur = ur / 3.0 + (2.0 / 3.0) * (u2 + dt * du)
    assert np.all(np.isfinite(u))
    assert np.max(np.abs(u - ur)) <= 1e-8 * np.max(np.abs(ur))
    # Dirichlet nodes are re-imposed at the start of every RHS call (adv_diff_test.jl:166-167), so they drift by O(dt) in between
    assert np.allclose(u[np.array(idx_bc[0])], 1.0, atol=1e-2)


def test_time_stepping_cuda_graph_replay():
    """The SSP-RK3 step of the config 1 example captured once into a CUDA graph (the context launches on the capture stream)
    and replayed: bit-identical fields to the eager loop."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("adv_diff_b200", os.path.join(os.path.dirname(__file__), "..", "examples", "adv_diff_b200.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    path = os.path.join(os.path.dirname(__file__), "golden", "rect_0_10.cgns")
    _, u_eager, _ = mod.run(steps=15, verbose=False, mesh=path)
    _, u_graph, _ = mod.run(steps=15, verbose=False, mesh=path, graph=True)
    assert np.array_equal(u_eager, u_graph)


def test_chunked_host_path_with_pinned_buffers(ctx):
    """rbffd_generate_operator_host overlaps the D2H of finished row chunks with the next chunk's solve when the caller's
    output buffers are pinned; the result must be bit-identical to the one-shot path (pageable buffers)."""
    import ctypes
    import torch
    X = rb.nodes.jittered_lattice(2, 120, seed=21)
    N, n, ops = len(X), 30, ["Lap", "Dx"]
    col_ref, val_ref = rb.generate_raw(X, None, 5, n, 3, ops, ctx=ctx)
    opts = rb.make_options(2, 5, n, 3, ops)
    Xh = torch.from_numpy(X).pin_memory()
    ch = torch.empty((N, n), dtype=torch.int64).pin_memory()
    vh = torch.empty((2, N, n), dtype=torch.float64).pin_memory()
    ctx._check(ctx._L.rbffd_generate_operator_host(ctx._h, ctypes.byref(opts), ctypes.c_void_p(Xh.data_ptr()), N, None, N, None,
                                                   ctypes.c_void_p(ch.data_ptr()), ctypes.c_void_p(vh.data_ptr())))
    assert np.array_equal(ch.numpy(), col_ref)
    assert np.array_equal(vh.numpy(), val_ref)
    # both ways of producing the caller's int64 pattern (int32 over PCIe + host threads / int64 widened on the device), 1-based
    import os
    opts1 = rb.make_options(2, 5, n, 3, ops, 1)
    for mode in ("1", "0"):
        os.environ["RBFFD_HOST_WIDEN"] = mode
        try:
            ch.zero_(); vh.zero_()
            ctx._check(ctx._L.rbffd_generate_operator_host(ctx._h, ctypes.byref(opts1), ctypes.c_void_p(Xh.data_ptr()), N, None, N, None,
                                                           ctypes.c_void_p(ch.data_ptr()), ctypes.c_void_p(vh.data_ptr())))
        finally:
            del os.environ["RBFFD_HOST_WIDEN"]
        assert np.array_equal(ch.numpy(), col_ref + 1)
        assert np.array_equal(vh.numpy(), val_ref)


def test_chunked_host_path_deferred_status(ctx, oracle):
    """The chunked host path queues every row chunk without a host synchronisation and inspects the status words once at the
    end: a chunk whose stencils the null-space kernel refuses is redone by the pivoted kernels (same weights as the
    synchronous path), and a singular stencil still surfaces as RBFFD_ERR_SINGULAR with its node index."""
    import ctypes
    import torch
    X = rb.nodes.jittered_lattice(2, 60, seed=22)
    N, n = len(X), 20
    Xd = X.copy()
    Xd[N - 7] = Xd[N - 8]                                   # duplicate node in the LAST chunk: singular interpolation matrix
    opts = rb.make_options(2, 3, n, 3, ["Lap"])
    Xh = torch.from_numpy(Xd).pin_memory()
    ch = torch.empty((N, n), dtype=torch.int64).pin_memory()
    vh = torch.empty((1, N, n), dtype=torch.float64).pin_memory()
    rc = ctx._L.rbffd_generate_operator_host(ctx._h, ctypes.byref(opts), ctypes.c_void_p(Xh.data_ptr()), N, None, N, None,
                                             ctypes.c_void_p(ch.data_ptr()), ctypes.c_void_p(vh.data_ptr()))
    assert rc == rb._lib.ERR_SINGULAR
    # nearly coincident nodes: the stencil is solvable but may fail the null-space kernel's checks; whatever path the chunk
    # takes, the pinned and pageable entry points must agree bit for bit and stay within tolerance of the oracle
    Xn = X.copy()
    Xn[N // 2] = Xn[N // 2 + 1] + 1e-9
    Xh.copy_(torch.from_numpy(Xn))
    ctx._check(ctx._L.rbffd_generate_operator_host(ctx._h, ctypes.byref(opts), ctypes.c_void_p(Xh.data_ptr()), N, None, N, None,
                                                   ctypes.c_void_p(ch.data_ptr()), ctypes.c_void_p(vh.data_ptr())))
    col_ref, val_ref = rb.generate_raw(Xn, None, 3, n, 3, ["Lap"], ctx=ctx)
    assert np.array_equal(ch.numpy(), col_ref)
    assert np.array_equal(vh.numpy(), val_ref)


def test_peer_memory_halo_exchange_two_gpus():
    """NVLink peer-memory halo exchange + overlapped sharded SpMV (tests/mgpu_halo_check.py) when >= 2 GPUs are visible."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", os.path.join(here, "mgpu_halo_check.py")], capture_output=True, text=True, timeout=300,
                       env={**os.environ, "HALO_G": "200"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout


def test_operator_combine(ctx, oracle):
    """rbffd_operator_combine_device: alpha*Lap - ux*Dx - uy*Dy as ONE value array over the shared pattern (the sparse sum the
    reference forms inside every cons_sys call, adv_diff_test.jl:151-152); applying it equals the fused multi-operator SpMV."""
    import torch
    X = rb.nodes.jittered_lattice(2, 45, seed=12)
    N, n = len(X), 30
    colind, vals = rb.generate_raw(X, None, 5, n, 3, ["Lap", "Dx", "Dy"], ctx=ctx)
    op = rb.Operator.from_host(ctx, colind, vals, N)
    coef = [0.7, -1.3, 0.4]
    vc = torch.empty((1, N, n), dtype=torch.float64, device="cuda")
    op.combine_device([0, 1, 2], coef, vc.data_ptr())
    ctx.synchronize()
    ref = coef[0] * vals[0] + coef[1] * vals[1] + coef[2] * vals[2]
    got = vc.cpu().numpy()[0]
    assert np.max(np.abs(got - ref)) <= 4 * EPS * np.max(np.abs(coef[0] * vals[0]) + np.abs(coef[1] * vals[1]) + np.abs(coef[2] * vals[2]))
    ci, _ = op.pointers(0)
    opc = ctx.operator_from_device(N, N, n, 1, ci, vc.data_ptr())
    u = torch.from_numpy(np.random.default_rng(1).standard_normal(N)).cuda()
    y1 = torch.empty(N, dtype=torch.float64, device="cuda")
    y2 = torch.empty(N, dtype=torch.float64, device="cuda")
    opc.spmv_device(0, u.data_ptr(), y1.data_ptr())
    op.spmv_multi_device([0, 1, 2], coef, u.data_ptr(), y2.data_ptr())
    ctx.synchronize()
    bound = oracle.spmv(colind, np.abs(ref), np.abs(u.cpu().numpy()))
    assert np.all(np.abs((y1 - y2).cpu().numpy()) <= 1e-13 * bound + 1e-300)
    with pytest.raises(rb.RbffdError):
        op.combine_device([0, 5], [1.0, 1.0], vc.data_ptr())


def _run_example_3d(nproc, g, steps, extra=()):
    import json
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    script = os.path.join(os.path.dirname(here), "examples", "adv_diff3d_sharded.py")
    cmd = [sys.executable, script] if nproc == 1 else [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                                                        "--master-addr", "127.0.0.1", "--master-port", "29583", script]
    r = subprocess.run(cmd + ["--g", str(g), "--steps", str(steps), *extra], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


def test_sharded_3d_time_stepping_example_one_gpu():
    """BASELINE configs[4] at reduced size on one GPU (examples/adv_diff3d_sharded.py): 3-D operators (n = 60, PHS r^7, degree 3)
    generated on the device, SSP-RK3 over the fused multi-operator SpMV; the advected, spreading Gaussian pulse is the known
    answer."""
    out = _run_example_3d(1, 30, 12)
    assert out["global_nodes"] == 27000 and out["n"] == 60
    assert out["rel_l2_error_vs_exact"] < 2e-2, out
    # the same run with the SSP-RK3 step replayed from a CUDA graph, and with the four operators applied separately
    rep = _run_example_3d(1, 30, 12, ("--graph",))
    assert rep["cuda_graph"] and rep["checksum"] == out["checksum"] and rep["rel_l2_error_vs_exact"] == out["rel_l2_error_vs_exact"]
    sep = _run_example_3d(1, 30, 12, ("--no-combine",))
    assert abs(sep["rel_l2_error_vs_exact"] - out["rel_l2_error_vs_exact"]) <= 1e-9 * out["rel_l2_error_vs_exact"]


def test_sharded_3d_time_stepping_example_two_gpus():
    """The same run on two GPUs (spatial-block shards, zero-communication generation, halo exchange fused into the SpMV launch,
    CUDA-graph replay) must reproduce the one-GPU result: identical stencils and weights, hence identical fields."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    one = _run_example_3d(1, 30, 12)
    two = _run_example_3d(2, 24, 12, ("--graph",))          # G = round(24 * 2^(1/3)) = 30: the same global lattice
    assert two["global_nodes"] == one["global_nodes"] and two["n_gpus"] == 2 and two["cuda_graph"]
    assert abs(two["rel_l2_error_vs_exact"] - one["rel_l2_error_vs_exact"]) <= 1e-9 * one["rel_l2_error_vs_exact"]
    for a, b in zip(one["checksum"], two["checksum"]):
        assert abs(a - b) <= 1e-11 * abs(a)


@pytest.mark.gpu
def test_legacy_collocated_methods(ctx, oracle):
    """generate_operator(X, p, n, polydeg) / hyperviscosity_operator(K, X, p, n, polydeg) (generate_operator.jl:354,
    hyperviscosity_operator.jl:314): unscaled stencils, centre at (eps, eps), RBF rows at X_j - x_c.  variant = 1."""
    X = rb.nodes.jittered_lattice(2, 40, seed=5)
    ops = ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy", ("Dk", 0, 2), ("Dk", 1, 2)]
    colind, vals = rb.generate_raw(X, None, 3, 20, 3, ops, ctx=ctx, variant=1)
    rcol, rvals, cond = oracle.generate_operator_collocated(X, 3, 20, 3, ops=ops, want_cond=True)
    assert np.array_equal(colind, rcol)
    _check_weights(vals, rvals, cond, ops)
    E, Dx, Dy, Dxx, Dyy, Dxy = rb.generate_operator_collocated(X, 3, 20, 3, ctx=ctx)
    Dxk, Dyk = rb.hyperviscosity_operator_collocated(2, X, 3, 20, 3, ctx=ctx)
    assert abs(Dxk - Dxx).max() <= 1e-9 * abs(Dxx).max() and abs(Dyk - Dyy).max() <= 1e-9 * abs(Dyy).max()
    assert np.allclose(np.asarray(E.sum(1)).ravel(), 1.0, atol=1e-8)
    with pytest.raises(rb.RbffdError):
        rb.generate_raw(X, X + 1e-3, 3, 20, 3, ["E"], ctx=ctx, variant=1)        # the legacy methods take X only
    with pytest.raises(rb.RbffdError):
        rb.generate_raw(X, None, 3, 20, 3, ["E"], ctx=ctx, variant=1, kernel=3)


def test_host_language_device_path(ctx, oracle):
    """What julia/RBFFDB200.jl does for the device-resident time loop, call for call through ctypes: operators generated from
    HOST node sets straight into HBM (rbffd_operator_generate_host), field vectors in library buffers (rbffd_device_*), one
    cons_sys + stage combination on them; and the int32 host index option (SparseMatrixCSC{Float64,Int32})."""
    import ctypes as C
    X = rb.nodes.jittered_lattice(2, 50, seed=8)
    N, n = len(X), 42
    names = ["E", "Dx", "Dy", "Dxx", "Dyy", ("Dk", 0, 4), ("Dk", 1, 4)]
    opts = rb.make_options(2, 5, n, 5, names)
    L = ctx._L
    h = C.c_void_p()
    ctx._check(L.rbffd_operator_generate_host(ctx._h, C.byref(opts), X.ctypes.data, N, None, N, None, C.byref(h)))
    op = rb.Operator(ctx, h)
    colind, vals = rb.generate_raw(X, None, 5, n, 5, names, ctx=ctx)
    lc, lv = op.to_host()
    assert np.array_equal(lc, colind) and np.array_equal(lv, vals)
    u = np.random.default_rng(4).standard_normal(N)
    bufs = []
    for _ in range(3):
        p = C.c_void_p()
        ctx._check(L.rbffd_device_malloc(ctx._h, 8 * N, C.byref(p)))
        bufs.append(p)
    du_d, u_d, out_d = bufs
    ctx._check(L.rbffd_device_upload(ctx._h, u_d, u.ctypes.data, 8 * N))
    gamma = 100 * (1 / 50) ** 4
    prm = rb.AdvDiffParams(iE=0, iDx=1, iDy=2, iDxx=3, iDyy=4, iDxk=5, iDyk=6, alpha=1.0, ux=0.3, uy=-0.2, gamma=gamma)
    op.rhs_advdiff_device(u_d, du_d, prm)
    ctx.stage_update_device(N, 0.75, u_d, 0.25, u_d, 1e-4, du_d, out_d)
    got = np.empty(N)
    ctx._check(L.rbffd_device_download(ctx._h, got.ctypes.data, out_d, 8 * N))
    du = oracle.rhs_advdiff(colind, *vals, 1.0, 0.3, -0.2, gamma, u)
    want = 0.75 * u + 0.25 * (u + 1e-4 * du)
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    for p in bufs:
        ctx._check(L.rbffd_device_free(ctx._h, p))
    op.close()
    # int32 indices straight into the caller's buffer, one-shot and pipelined (pinned) paths, 1-based as Julia asks
    c32, v32 = rb.generate_raw(X, None, 5, n, 5, names[:3], ctx=ctx, index_base=1, index_width=32)
    assert c32.dtype == np.int32 and np.array_equal(c32.astype(np.int64), colind + 1) and np.array_equal(v32, vals[:3])
    import torch
    o32 = rb.make_options(2, 5, n, 5, names[:3], index_base=1, index_width=32)
    Xh = torch.from_numpy(X).pin_memory()
    ch = torch.empty((N, n), dtype=torch.int32).pin_memory()
    vh = torch.empty((3, N, n), dtype=torch.float64).pin_memory()
    ctx._check(L.rbffd_generate_operator_host(ctx._h, C.byref(o32), Xh.data_ptr(), N, None, N, None, ch.data_ptr(), vh.data_ptr()))
    assert np.array_equal(ch.numpy().astype(np.int64), colind + 1) and np.array_equal(vh.numpy(), vals[:3])
