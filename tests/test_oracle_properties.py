"""Reference-free checks of the CPU oracle (SURVEY.md §9 'independent checks') and of the host-side helpers."""
import numpy as np
import pytest

import rbffd_b200 as rb


def test_kdtree_equals_bruteforce(oracle):
    rng = np.random.default_rng(1)
    for d, N, k in ((2, 3000, 30), (3, 2500, 60), (2, 500, 1)):
        X = rng.random((N, d))
        Q = rng.random((400, d))
        i1, d1 = oracle.knn(X, Q, k)
        i2, d2 = oracle.knn(X, Q, k, brute=True)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
        assert np.all(np.diff(d1, axis=1) >= 0)


def test_knn_ties_broken_by_index(oracle):
    g = np.arange(8.0)
    X = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)      # exact lattice: many exact distance ties
    idx, d2 = oracle.knn(X, X, 9)
    idb, d2b = oracle.knn(X, X, 9, brute=True)
    assert np.array_equal(idx, idb)
    for r in range(len(X)):
        key = list(zip(d2[r], idx[r]))
        assert key == sorted(key) and idx[r, 0] == r


def test_masked_rule_of_calculateneighbors(oracle):
    """calculateneighbors.jl:16-42: boundary/ghost nodes of boundary b see interior + all boundary + ghosts of b only."""
    rng = np.random.default_rng(2)
    N = 600
    X = rng.random((N, 2))
    idx_in = (0, 400)
    bc = [(400, 450), (450, 500)]
    gh = [(500, 550), (550, 600)]
    idx, cy, _, _ = oracle.calculateneighbors(X, X, 12, idx_in, bc, gh)
    kind, bnd = oracle.groups_from_ranges(N, idx_in, bc, gh)
    for i in range(N):
        if kind[i] == 0:
            continue
        nb = idx[i]
        ghosts = nb[kind[nb] == 2]
        assert np.all(bnd[ghosts] == bnd[i])
    free, _ = oracle.knn(X, X, 12)
    assert np.array_equal(idx[:400], free[:400])        # interior queries see everything (:83-87)
    assert np.array_equal(cy, np.arange(N))


@pytest.mark.parametrize("p,alpha", [(5, (4, 0)), (7, (6, 0)), (3, (2, 0)), (5, (1, 1)), (7, (0, 2, 1))])
def test_rbf_derivative_tables_match_sympy(oracle, p, alpha):
    import sympy as sp
    d = len(alpha)
    xs = sp.symbols("x y z")[:d]
    r = sp.sqrt(sum(v**2 for v in xs))
    expr = r**p
    for v, a in zip(xs, alpha):
        expr = sp.diff(expr, v, a)
    tab = oracle.rbf_derivative_table(p, d, alpha)
    rng = np.random.default_rng(0)
    for _ in range(5):
        pt = rng.standard_normal(d)
        rr = np.linalg.norm(pt)
        mine = sum(c * np.prod(pt ** np.array([e0, e1, e2][:d])) * rr**q for c, e0, e1, e2, q in tab)
        ref = float(expr.subs(dict(zip(xs, pt))))
        assert abs(mine - ref) <= 1e-11 * max(1.0, abs(ref))


def test_known_closed_forms(oracle):
    # SURVEY.md §8a probe: p=5,K=4 -> 45 r + 90 x^2/r - 15 x^4/r^3
    tab = {(int(e0), int(q)): c for c, e0, _, _, q in oracle.rbf_derivative_table(5, 2, (4, 0))}
    assert tab == {(0, 1): 45.0, (2, -1): 90.0, (4, -3): -15.0}


@pytest.mark.parametrize("d,p,deg,n", [(2, 5, 3, 30), (2, 3, 3, 20), (3, 7, 3, 60), (2, 5, 4, 50)])
def test_polynomial_reproduction_and_row_sums(oracle, d, p, deg, n):
    g = 24 if d == 2 else 9
    X = rb.nodes.jittered_lattice(d, g, seed=3)
    names = ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy", "Lap"] + (["Dz", "Dzz"] if d == 3 else [])
    colind, vals, cond = oracle.generate_operator(X, X, p, n, deg, ops=names, want_cond=True)
    ex = oracle.monomial_exponents(d, deg)[:, :d]
    tol = 200 * np.finfo(float).eps * cond.max()
    alphas = {"E": (0, 0, 0), "Dx": (1, 0, 0), "Dy": (0, 1, 0), "Dxx": (2, 0, 0), "Dyy": (0, 2, 0), "Dxy": (1, 1, 0),
              "Dz": (0, 0, 1), "Dzz": (0, 0, 2)}

    def mono_deriv(e, al, P):
        out = np.ones(len(P))
        for a in range(d):
            if e[a] < al[a]:
                return np.zeros(len(P))
            c = np.prod([e[a] - t for t in range(al[a])]) if al[a] else 1.0
            out = out * c * P[:, a] ** (e[a] - al[a])
        return out

    for oi, nm in enumerate(names):
        W = vals[oi]
        for e in ex:
            f = np.prod(X ** e, axis=1)
            got = np.einsum("kj,kj->k", W, f[colind])
            if nm == "Lap":
                want = sum(mono_deriv(e, tuple(2 if b == a else 0 for b in range(3)), X) for a in range(d))
            else:
                want = mono_deriv(e, alphas[nm], X)
            scale = np.abs(W).sum(1).max()
            assert np.max(np.abs(got - want)) <= tol * scale, (nm, e)
    E = vals[0]
    assert np.max(np.abs(E.sum(1) - 1)) < 1e-9
    # Y_k == X_c  =>  E row = unit vector
    assert np.allclose(E[:, 0], 1.0, atol=1e-7) and np.max(np.abs(E[:, 1:])) < 1e-7
    # Lap == Dxx + Dyy (+ Dzz) to rounding
    lap = vals[names.index("Lap")]
    s2 = vals[names.index("Dxx")] + vals[names.index("Dyy")] + (vals[names.index("Dzz")] if d == 3 else 0)
    assert np.max(np.abs(lap - s2)) <= tol * np.abs(s2).max()


def test_inverse_mode_vs_lu_vs_long_double(oracle):
    X = rb.nodes.jittered_lattice(2, 20, seed=5)
    idx, _ = oracle.knn(X, X, 30)
    ops = oracle.op_table(2, ["Dxx", "Dy"])
    c = np.arange(len(X))
    w0, cond = oracle.weights(X, X, idx, c, 5, 30, 3, ops, mode=0, want_cond=True)
    w1 = oracle.weights(X, X, idx, c, 5, 30, 3, ops, mode=1)
    w2 = oracle.weights(X, X, idx, c, 5, 30, 3, ops, mode=2)
    eps = np.finfo(float).eps
    for w in (w0, w1):
        err = np.abs(w - w2).max(axis=2) / np.abs(w2).max(axis=2)
        assert np.all(err <= 50 * eps * cond[None, :])


def test_lattice_generator_is_deterministic_and_sliceable():
    A = rb.nodes.jittered_lattice(3, 7, seed=11)
    B = rb.nodes.jittered_lattice(3, 7, seed=11, first=100, count=50)
    assert np.array_equal(A[100:150], B)
    assert A.min() > 0 and A.max() < 1
    g = 7
    cell = np.floor(A * g).astype(int)
    lin = cell[:, 0] + g * cell[:, 1] + g * g * cell[:, 2]
    assert np.array_equal(lin, np.arange(g**3))           # one node per lattice cell


def test_spmv_and_rhs_oracle_against_scipy(oracle):
    import scipy.sparse as sp
    rng = np.random.default_rng(4)
    M = N = 500
    n = 9
    colind = np.stack([rng.choice(N, n, replace=False) for _ in range(M)])
    mats = rng.standard_normal((7, M, n))
    u = rng.standard_normal(N)
    S = [sp.csr_matrix((m.ravel(), colind.ravel(), np.arange(0, M * n + 1, n)), shape=(M, N)) for m in mats]
    E, Dx, Dy, Dxx, Dyy, Dxk, Dyk = S
    assert np.allclose(oracle.spmv(colind, mats[1], u, 2.0), 2.0 * (Dx @ u), rtol=1e-13, atol=1e-13)
    assert np.allclose(oracle.spmv_t(colind, mats[0], u, N), E.T @ u, rtol=1e-13, atol=1e-13)
    a, ux, uy, gam = 1.0, 0.3, -0.2, 1e-3
    want = E.T @ (a * (Dxx @ u) + a * (Dyy @ u) - ux * (Dx @ u) - uy * (Dy @ u)) - gam * ((Dxk + Dyk) @ u)
    got = oracle.rhs_advdiff(colind, *mats, a, ux, uy, gam, u)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


def _legacy_numpy(X, idx, p, polydeg):
    """Line-by-line numpy replay of generate_operator(X, p, n, polydeg) (src/generate_operator.jl:388-468) for one
    stencil at a time: X_shift with the centre at (eps, eps), A = [Phi P; P' 0] unscaled, inv(A) * RHS with the
    Symbolics closed forms written out (comments at :139-153)."""
    eps = np.finfo(float).eps
    ex = [(a, g - a) for g in range(polydeg + 1) for a in range(g, -1, -1)]
    out = np.zeros((6, len(X), idx.shape[1]))
    for i in range(len(X)):
        S = X[idx[i]] - X[idx[i][0]]
        S[0] = eps
        n = len(S)
        d = S[:, None, :] - S[None, :, :]
        Phi = np.sqrt((d ** 2).sum(-1)) ** p
        P = np.stack([S[:, 0] ** a * S[:, 1] ** b for a, b in ex], 1)
        q = P.shape[1]
        A = np.block([[Phi, P], [P.T, np.zeros((q, q))]])
        x, y = S[:, 0], S[:, 1]
        r = np.hypot(x, y)
        b = r ** p
        bx, by = p * x * r ** (p - 2), p * y * r ** (p - 2)
        bxx = p * r ** (p - 2) + p * (p - 2) * x ** 2 * r ** (p - 4)
        byy = p * r ** (p - 2) + p * (p - 2) * y ** 2 * r ** (p - 4)
        bxy = p * (p - 2) * x * y * r ** (p - 4)

        def mono(a, b_, da, db):
            if a < da or b_ < db:
                return 0.0
            ca = np.prod([a - t for t in range(da)]) if da else 1.0
            cb = np.prod([b_ - t for t in range(db)]) if db else 1.0
            return ca * cb * eps ** (a - da) * eps ** (b_ - db)
        rows = []
        for (da, db), rb_ in (((0, 0), b), ((1, 0), bx), ((0, 1), by), ((2, 0), bxx), ((0, 2), byy), ((1, 1), bxy)):
            rows.append(np.concatenate([rb_, [mono(a, b_, da, db) for a, b_ in ex]]))
        W = np.linalg.inv(A) @ np.stack(rows, 1)
        out[:, i, :] = W[:n].T
    return out


def test_legacy_collocated_method(oracle):
    """generate_operator(X, p, n, polydeg) (generate_operator.jl:354-491): the oracle's variant 1 against a literal numpy
    replay, polynomial reproduction of every operator (the constraint rows hold whatever the RBF rows are), and the sign
    quirk of the odd-derivative RBF rows relative to the two-set method."""
    X = rb.nodes.jittered_lattice(2, 14, seed=4)
    p, n, deg = 3, 16, 2
    idx, vals, cond = oracle.generate_operator_collocated(X, p, n, deg, want_cond=True)
    ref = _legacy_numpy(X, idx, p, deg)
    tol = 200 * np.finfo(float).eps * cond[:, None]
    for o in range(6):
        assert np.all(np.abs(vals[o] - ref[o]) <= tol * np.abs(ref[o]).max(1, keepdims=True))
    # polynomial reproduction at the centre: sum_j w_j mono(X_j - x_c) = (L mono)(0) up to the (eps, eps) shift
    names = ["E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"]
    want = {"E": (0, 0), "Dx": (1, 0), "Dy": (0, 1), "Dxx": (2, 0), "Dyy": (0, 2), "Dxy": (1, 1)}
    S = X[idx] - X[idx[:, :1]]
    h = np.abs(S).max()
    for o, nm in enumerate(names):
        da, db = want[nm]
        for a in range(deg + 1):
            for b_ in range(deg + 1 - a):
                lhs = (vals[o] * S[:, :, 0] ** a * S[:, :, 1] ** b_).sum(1)
                exact = float(np.prod([a - t for t in range(da)]) * np.prod([b_ - t for t in range(db)])) if (a, b_) == (da, db) else 0.0
                scale = np.abs(vals[o]).sum(1) * h ** (a + b_)
                assert np.all(np.abs(lhs - exact) <= 1e3 * np.finfo(float).eps * cond * np.maximum(scale, 1.0))
    # two-set method on the same nodes: even-order operators agree to truncation level, Dx/Dy differ (flipped RBF rows)
    _, v2 = oracle.generate_operator(X, X, p, n, deg)
    interior = np.all((X > 0.3) & (X < 0.7), axis=1)
    rel = lambda a_, b_: np.abs(a_ - b_).max() / np.abs(b_).max()
    assert rel(vals[3][interior], v2[3][interior]) < 0.2          # Dxx: same operator up to the anisotropic-scaling effect
    assert rel(vals[1][interior], v2[1][interior]) > 1e-3         # Dx: the sign quirk is really there


def _twoset_numpy(X, Y, idx, center, rows, p, polydeg):
    """Line-by-line numpy replay of the two-set generate_operator(X, Y, p, n, polydeg) (src/generate_operator.jl:29-190)
    for the Y rows `rows`, with LAPACK doing what LinearAlgebra does there: scalestencil (src/scalestencil.jl:10-20),
    A = [Phi P; P' 0] (src/interpolationmatrix.jl:5), numpy.linalg.inv = getrf + getri (:8), inv(A) * RHS (:158), chain-rule
    factors (:161-166).  Returns [6, len(rows), n] in the order E, Dx, Dy, Dxx, Dyy, Dxy, and cond_1(A) per row."""
    eps = np.finfo(float).eps
    ex = [(a, g - a) for g in range(polydeg + 1) for a in range(g, -1, -1)]
    n = idx.shape[1]
    out = np.zeros((6, len(rows), n))
    cond = np.zeros(len(rows))
    for kk, k in enumerate(rows):
        c = center[k]
        st = idx[c]
        Xs = X[st] - X[st[0]]
        sx, sy = 1.0 / np.abs(Xs[:, 0]).max(), 1.0 / np.abs(Xs[:, 1]).max()
        S = Xs * np.array([sx, sy])
        d = S[:, None, :] - S[None, :, :]
        Phi = np.sqrt((d ** 2).sum(-1)) ** p
        P = np.stack([S[:, 0] ** a * S[:, 1] ** b for a, b in ex], 1)
        q = P.shape[1]
        A = np.block([[Phi, P], [P.T, np.zeros((q, q))]])
        Ainv = np.linalg.inv(A)
        cond[kk] = np.abs(A).sum(0).max() * np.abs(Ainv).sum(0).max()
        ys = (Y[k] - X[st[0]]) * np.array([sx, sy])
        dx, dy = ys[0] - S[:, 0], ys[1] - S[:, 1]
        dx[dx == 0] = eps
        dy[dy == 0] = eps
        r = np.hypot(dx, dy)
        b = r ** p
        bx, by = p * dx * r ** (p - 2), p * dy * r ** (p - 2)
        bxx = p * r ** (p - 2) + p * (p - 2) * dx ** 2 * r ** (p - 4)
        byy = p * r ** (p - 2) + p * (p - 2) * dy ** 2 * r ** (p - 4)
        bxy = p * (p - 2) * dx * dy * r ** (p - 4)

        def mono(a, b_, da, db):          # d^(da,db) of x^a y^b_ at the scaled evaluation point (polylinearoperator.jl:36-44)
            if a < da or b_ < db:
                return 0.0
            ca = float(np.prod([a - t for t in range(da)])) if da else 1.0
            cb = float(np.prod([b_ - t for t in range(db)])) if db else 1.0
            return ca * cb * ys[0] ** (a - da) * ys[1] ** (b_ - db)
        cols = []
        for (da, db), rb_ in (((0, 0), b), ((1, 0), bx), ((0, 1), by), ((2, 0), bxx), ((0, 2), byy), ((1, 1), bxy)):
            cols.append(np.concatenate([rb_, [mono(a, b_, da, db) for a, b_ in ex]]))
        W = Ainv @ np.stack(cols, 1)
        f = [1.0, sx, sy, sx * sx, sy * sy, sx * sy]
        for o in range(6):
            out[o, kk] = f[o] * W[:n, o]
    return out, cond


def test_twoset_weights_against_lapack_inverse(oracle, tominec):
    """Pins the oracle's weight-level arithmetic (mode 0: LU + explicit inverse + inv(A)*RHS) to a real LAPACK getrf/getri on
    the reference's own two-set fixture (test/data/x_nodes_fitted.csv, y_nodes_fitted.csv; Y != X, eta != 0 rows): the
    oracle and the numpy replay of generate_operator.jl:29-190 must agree to 10 eps cond_1(A) per row, for all six
    operators, at the parameters of test/poisson_test.jl:56-60."""
    X, Y = tominec["X"], tominec["Y"]
    p, n, deg = 3, 2 * 15, 4            # poisson_test.jl: rbfdeg = 3, polydeg = 4, n = 2 * binomial(polydeg + 2, 2)
    colind, vals, cond_x = oracle.generate_operator(X, Y, p, n, deg, want_cond=True)
    idx, _ = oracle.knn(X, X, n)
    center = oracle.knn(X, Y, 1)[0][:, 0]
    assert np.array_equal(colind, idx[center])
    rows = np.random.default_rng(5).choice(len(Y), 400, replace=False)
    ref, cond = _twoset_numpy(X, Y, idx, center, rows, p, deg)
    assert np.allclose(cond, cond_x[center[rows]], rtol=1e-6)        # same matrices on both sides
    eps = np.finfo(float).eps
    worst = 0.0
    for o in range(6):
        err = np.abs(vals[o][rows] - ref[o]).max(1)
        tol = 10 * eps * cond * np.abs(ref[o]).max(1)
        worst = max(worst, (err / tol).max())
    assert worst <= 1.0, f"oracle vs LAPACK inverse: {worst:.2f} x (10 eps cond)"
