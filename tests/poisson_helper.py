"""test/poisson_test.jl:19-132 replayed with a pluggable generate_operator (shared by the oracle and GPU tests)."""
import numpy as np
import scipy.sparse as sp


def csr(colind, vals, ncols):
    M, n = colind.shape
    return sp.csr_matrix((vals.ravel(), colind.ravel(), np.arange(0, M * n + 1, n)), shape=(M, ncols))


def poisson_error(tominec, gen):
    """test/poisson_test.jl:19-132 with `gen` standing in for generate_operator (0-based indices)."""
    X = tominec["X"].copy()
    Y = tominec["Y"].copy()
    N, M = len(X), len(Y)
    iin = tominec["Y_idx_in"] - 1
    idi = tominec["Y_idx_dirichlet"] - 1
    ine = tominec["Y_idx_neumann"] - 1
    xn, yn = tominec["x_normals"][:, 2], tominec["y_normals"][:, 2]
    # :34-45 overwrite the Y node nearest to each X node by that X node
    from scipy.spatial import cKDTree
    nearest = cKDTree(Y).query(X, 1)[1]
    for i in range(N):
        Y[nearest[i]] = X[i]
    p, polydeg = 3, 3
    n = 2 * 10
    colind, vals = gen(X, Y, p, n, polydeg)
    E, Dx, Dy, Dxx, Dyy, Dxy = (csr(colind, v, N) for v in vals)
    u_exact = lambda x, y: np.sin(2 * np.pi * x * y)
    f2 = lambda x, y: -4.0 * x**2 * np.pi**2 * np.sin(2 * np.pi * x * y) - 4.0 * y**2 * np.pi**2 * np.sin(2 * np.pi * x * y)
    f1 = lambda n1, n2, x, y: n2 * x * np.pi * np.cos(2 * np.pi * x * y) * 2.0 + n1 * y * np.pi * np.cos(2 * np.pi * x * y) * 2.0
    D = np.zeros((M, N))
    D[iin] = (Dxx + Dyy)[iin].toarray()
    D[ine] = xn[:, None] * Dx[ine].toarray() + yn[:, None] * Dy[ine].toarray()
    D[idi] = E[idi].toarray()
    f = np.zeros(M)
    f[iin] = f2(Y[iin, 0], Y[iin, 1])
    f[ine] = f1(xn, yn, Y[ine, 0], Y[ine, 1])
    f[idi] = u_exact(Y[idi, 0], Y[idi, 1])
    h = np.mean(cKDTree(X).query(X, 2)[0][:, 1])
    M0, M1, M2 = len(idi), len(ine), len(iin)
    D[iin] *= 1 / np.sqrt(M2); f[iin] *= 1 / np.sqrt(M2)
    D[ine] *= 1 / np.sqrt(M1); f[ine] *= 1 / np.sqrt(M1)
    D[idi] *= 1 / h / np.sqrt(M0); f[idi] *= 1 / h / np.sqrt(M0)
    u = np.linalg.lstsq(D, f, rcond=None)[0]
    uY = E @ u
    ue = u_exact(Y[:, 0], Y[:, 1])
    return np.linalg.norm(uY - ue) / np.linalg.norm(ue)




def _collocation_error_device(ctx, X, Y, iin, idi, ine, xn, yn, p, n, polydeg, tol=1e-11):
    """Everything after the node sets on the device: operators generated into HBM, the collocation matrix assembled from their
    row blocks (poisson_test.jl:76-79 with the scalings of :110-118), `u = D \\ f` (:121) by CGLS over the library's SpMV /
    transposed SpMV, the evaluation `E*u` (:124) by SpMV.  Returns (rel. l2 error, CGLS iterations, relative residual)."""
    import torch
    import rbffd_b200 as rb
    from scipy.spatial import cKDTree
    N, M = len(X), len(Y)
    dev = torch.device("cuda", ctx.device)
    Xd, Yd = torch.from_numpy(np.ascontiguousarray(X)).to(dev), torch.from_numpy(np.ascontiguousarray(Y)).to(dev)
    op = ctx.operator_generate(rb.make_options(2, p, n, polydeg, rb.REFERENCE_OPS), Xd.data_ptr(), N, Yd.data_ptr(), M)   # E Dx Dy Dxx Dyy Dxy
    u_exact = lambda x, y: np.sin(2 * np.pi * x * y)
    f2 = lambda x, y: -4.0 * x**2 * np.pi**2 * np.sin(2 * np.pi * x * y) - 4.0 * y**2 * np.pi**2 * np.sin(2 * np.pi * x * y)
    f1 = lambda n1, n2, x, y: n2 * x * np.pi * np.cos(2 * np.pi * x * y) * 2.0 + n1 * y * np.pi * np.cos(2 * np.pi * x * y) * 2.0
    h = np.mean(cKDTree(X).query(X, 2)[0][:, 1])
    s0, s1, s2 = 1 / h / np.sqrt(len(idi)), 1 / np.sqrt(len(ine)), 1 / np.sqrt(len(iin))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    D, keep = rb.lsq.mixed_row_operator(ctx, op, [(t(iin), [(3, float(s2)), (4, float(s2))]),
                                                  (t(ine), [(1, t(xn * s1)), (2, t(yn * s1))]),
                                                  (t(idi), [(0, float(s0))])], M)
    f = np.zeros(M)
    f[iin] = f2(Y[iin, 0], Y[iin, 1]) * s2
    f[ine] = f1(xn, yn, Y[ine, 0], Y[ine, 1]) * s1
    f[idi] = u_exact(Y[idi, 0], Y[idi, 1]) * s0
    u, iters, rel = rb.lsq.cgls(D, 0, t(f), tol=tol)
    uY = torch.empty(M, dtype=torch.float64, device=dev)
    op.spmv_device(0, u.data_ptr(), uY.data_ptr())
    ctx.synchronize()
    ue = u_exact(Y[:, 0], Y[:, 1])
    return float(np.linalg.norm(uY.cpu().numpy() - ue) / np.linalg.norm(ue)), iters, rel


def poisson_error_device(tominec, ctx):
    """test/poisson_test.jl:19-132 on the device (p = 3, n = 20, polydeg = 3)."""
    from scipy.spatial import cKDTree
    X = tominec["X"].copy()
    Y = tominec["Y"].copy()
    iin, idi, ine = tominec["Y_idx_in"] - 1, tominec["Y_idx_dirichlet"] - 1, tominec["Y_idx_neumann"] - 1
    xn, yn = tominec["x_normals"][:, 2], tominec["y_normals"][:, 2]
    nearest = cKDTree(Y).query(X, 1)[1]
    for i in range(len(X)):
        Y[nearest[i]] = X[i]
    return _collocation_error_device(ctx, X, Y, iin, idi, ine, xn, yn, 3, 20, 3)


def mesh_import_error_device(cgns_path, X, ctx):
    """test/mesh_import_test.jl:19-158 on the device (p = 3, n = 42, polydeg = 5; Y from processmesh on the CGNS mesh)."""
    import rbffd_b200 as rb
    from scipy.spatial import cKDTree
    Y_, _, iin, ibc, _, _, normals, _ = rb.mesh.processmesh(cgns_path, ["dirichlet", "neumann"], ctx=ctx)
    iin, idi, ine = (np.arange(r.start, r.stop) for r in (iin, ibc[0], ibc[1]))
    Y = np.concatenate([Y_[iin], Y_[idi], Y_[ine]])
    xn, yn = normals[1][:, 0], normals[1][:, 1]
    X = X.copy()
    nearest = cKDTree(Y).query(X, 1)[1]
    for i in range(len(X)):
        Y[nearest[i]] = X[i]
    return _collocation_error_device(ctx, X, Y, iin, idi, ine, xn, yn, 3, 42, 5)


def mesh_import_error(cgns_path, X, gen, ctx=None):
    """test/mesh_import_test.jl:19-158 with `gen` standing in for generate_operator: Y comes from processmesh on the CGNS
    mesh (ghost nodes dropped, :52), Neumann normals from the mesh (:38-40), p = 3, polydeg = 5, n = 42."""
    import rbffd_b200 as rb
    from scipy.spatial import cKDTree
    Y_, _, iin, ibc, _, _, normals, _ = rb.mesh.processmesh(cgns_path, ["dirichlet", "neumann"], ctx=ctx)
    iin, idi, ine = (np.arange(r.start, r.stop) for r in (iin, ibc[0], ibc[1]))
    Y = np.concatenate([Y_[iin], Y_[idi], Y_[ine]])
    assert iin[0] == 0 and idi[0] == len(iin) and ine[0] == len(iin) + len(idi)      # :52-54 keep the index sets valid
    xn, yn = normals[1][:, 0], normals[1][:, 1]
    X = X.copy()
    N, M = len(X), len(Y)
    nearest = cKDTree(Y).query(X, 1)[1]                                              # :57-68 (HNSW there; exact here)
    for i in range(N):
        Y[nearest[i]] = X[i]
    p, polydeg = 3, 5
    n = 2 * 21
    colind, vals = gen(X, Y, p, n, polydeg)
    E, Dx, Dy, Dxx, Dyy, Dxy = (csr(colind, v, N) for v in vals)
    u_exact = lambda x, y: np.sin(2 * np.pi * x * y)
    f2 = lambda x, y: -4.0 * x**2 * np.pi**2 * np.sin(2 * np.pi * x * y) - 4.0 * y**2 * np.pi**2 * np.sin(2 * np.pi * x * y)
    f1 = lambda n1, n2, x, y: n2 * x * np.pi * np.cos(2 * np.pi * x * y) * 2.0 + n1 * y * np.pi * np.cos(2 * np.pi * x * y) * 2.0
    D = np.zeros((M, N))
    D[iin] = (Dxx + Dyy)[iin].toarray()
    D[ine] = xn[:, None] * Dx[ine].toarray() + yn[:, None] * Dy[ine].toarray()
    D[idi] = E[idi].toarray()
    f = np.zeros(M)
    f[iin] = f2(Y[iin, 0], Y[iin, 1])
    f[ine] = f1(xn, yn, Y[ine, 0], Y[ine, 1])
    f[idi] = u_exact(Y[idi, 0], Y[idi, 1])
    h = np.mean(cKDTree(X).query(X, 2)[0][:, 1])
    M0, M1, M2 = len(idi), len(ine), len(iin)
    D[iin] *= 1 / np.sqrt(M2); f[iin] *= 1 / np.sqrt(M2)
    D[ine] *= 1 / np.sqrt(M1); f[ine] *= 1 / np.sqrt(M1)
    D[idi] *= 1 / h / np.sqrt(M0); f[idi] *= 1 / h / np.sqrt(M0)
    u = np.linalg.lstsq(D, f, rcond=None)[0]
    uY = E @ u
    ue = u_exact(Y[:, 0], Y[:, 1])
    return np.linalg.norm(uY - ue) / np.linalg.norm(ue)
