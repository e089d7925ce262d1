"""ctypes binding of the CPU oracle (oracle/rbffd_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package never does (and fails loudly when its CUDA library is missing).

Function names mirror the reference (0-based indices here, the Julia API is 1-based):
  knn                      NearestNeighbors.knn as called at src/generate_operator.jl:43-47
  calculateneighbors       src/calculateneighbors.jl:1-97
  generate_operator        src/generate_operator.jl:29-190  (+ :192-352 via `groups`)
  hyperviscosity_operator  src/hyperviscosity_operator.jl:26-175
  rhs_advdiff              examples/adv_diff_test.jl:144-152
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OP_DERIV = 0
OP_LAPLACE = 1


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "rbffd_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_num_threads.restype = C.c_int
        _LIB.orc_num_monomials.restype = C.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(nthreads: int) -> int:
    """OpenMP threads of the oracle's parallel loops (overrides OMP_NUM_THREADS); returns the value now in effect."""
    lib().orc_set_num_threads(int(nthreads))
    return num_threads()


def num_monomials(d: int, deg: int) -> int:
    return lib().orc_num_monomials(d, deg)


def monomial_exponents(d: int, deg: int) -> np.ndarray:
    q = num_monomials(d, deg)
    ex = np.zeros((q, 3), np.int32)
    lib().orc_monomial_exponents(d, deg, _p(ex, C.c_int32))
    return ex


def rbf_derivative_table(p: int, d: int, alpha) -> np.ndarray:
    al = np.zeros(3, np.int32)
    al[: len(alpha)] = alpha
    out = np.zeros((64, 5))
    n = lib().orc_rbf_derivative_table(p, d, _p(al, C.c_int32), _p(out, C.c_double))
    return out[:n]


def knn(X, Q, k, xkind=None, xbnd=None, qkind=None, qbnd=None, brute=False):
    X = np.ascontiguousarray(X, np.float64)
    Q = np.ascontiguousarray(Q, np.float64)
    N, d = X.shape
    NQ = Q.shape[0]
    idx = np.empty((NQ, k), np.int64)
    d2 = np.empty((NQ, k), np.float64)
    conv = lambda a: None if a is None else np.ascontiguousarray(a, np.int32)
    xkind, xbnd, qkind, qbnd = conv(xkind), conv(xbnd), conv(qkind), conv(qbnd)
    rc = lib().orc_knn(_p(X, C.c_double), C.c_int64(N), d, _p(Q, C.c_double), C.c_int64(NQ), k,
                       _p(xkind, C.c_int32), _p(xbnd, C.c_int32), _p(qkind, C.c_int32), _p(qbnd, C.c_int32),
                       int(brute), _p(idx, C.c_int64), _p(d2, C.c_double))
    if rc:
        raise ValueError("orc_knn: invalid arguments")
    return idx, d2


def groups_from_ranges(N, idx_in, idx_bc, idx_bc_g):
    """kind/bnd arrays from the reference's index sets (0-based, half-open (lo, hi) ranges or index arrays)."""
    kind = np.zeros(N, np.int32)
    bnd = np.zeros(N, np.int32)
    for b, r in enumerate(idx_bc):
        ii = np.arange(r[0], r[1]) if isinstance(r, tuple) else np.asarray(r)
        kind[ii] = 1
        bnd[ii] = b
    for b, r in enumerate(idx_bc_g):
        ii = np.arange(r[0], r[1]) if isinstance(r, tuple) else np.asarray(r)
        kind[ii] = 2
        bnd[ii] = b
    return kind, bnd


def calculateneighbors(X, Y, n, X_idx_in, X_idx_bc, X_idx_bc_g, brute=False):
    """src/calculateneighbors.jl:1-97 -> (idxs_x [N,n], idxs_y_x [M], d2_x, d2_y_x); squared distances."""
    X = np.ascontiguousarray(X, np.float64)
    kind, bnd = groups_from_ranges(X.shape[0], X_idx_in, X_idx_bc, X_idx_bc_g)
    idx, d2 = knn(X, X, n, kind, bnd, kind, bnd, brute)
    cy, d2y = knn(X, Y, 1, None, None, None, None, brute)
    return idx, cy[:, 0], d2, d2y[:, 0]


def weights(X, Y, idx, center, p, n, polydeg, ops, mode=0, variant=0, want_cond=False):
    X = np.ascontiguousarray(X, np.float64)
    Y = np.ascontiguousarray(Y, np.float64)
    idx = np.ascontiguousarray(idx, np.int64)
    center = np.ascontiguousarray(center, np.int64)
    ops = np.ascontiguousarray(ops, np.int32).reshape(-1, 4)
    N, d = X.shape
    M = Y.shape[0]
    vals = np.zeros((ops.shape[0], M, n))
    cond = np.zeros(N) if want_cond else None
    rc = lib().orc_weights(_p(X, C.c_double), C.c_int64(N), d, _p(Y, C.c_double), C.c_int64(M),
                           _p(idx, C.c_int64), _p(center, C.c_int64), p, n, polydeg, ops.shape[0],
                           _p(ops, C.c_int32), mode, variant, _p(vals, C.c_double), _p(cond, C.c_double))
    if rc != 0:
        raise np.linalg.LinAlgError(f"orc_weights: singular stencil at node {rc - 1}" if rc > 0 else "bad args")
    return (vals, cond) if want_cond else vals


def op_table(d, names):
    """operator names -> [kind, a0, a1, a2] rows.  E, Dx, Dy, Dz, Dxx, Dyy, Dzz, Dxy, Dxz, Dyz, Lap, ('Dk', axis, K)"""
    tab = {"E": (0, 0, 0), "Dx": (1, 0, 0), "Dy": (0, 1, 0), "Dz": (0, 0, 1), "Dxx": (2, 0, 0), "Dyy": (0, 2, 0),
           "Dzz": (0, 0, 2), "Dxy": (1, 1, 0), "Dxz": (1, 0, 1), "Dyz": (0, 1, 1)}
    out = []
    for nm in names:
        if nm == "Lap":
            out.append([OP_LAPLACE, 0, 0, 0])
        elif isinstance(nm, tuple):
            a = [0, 0, 0]
            a[nm[1]] = nm[2]
            out.append([OP_DERIV] + a)
        else:
            out.append([OP_DERIV] + list(tab[nm]))
    return np.asarray(out, np.int32)


def sort_rows(colind, vals_list):
    """column-sorted rows: the bit-comparable CSR form (reference sorts into CSC, generate_operator.jl:177)."""
    order = np.argsort(colind, axis=1, kind="stable")
    ci = np.take_along_axis(colind, order, 1)
    return ci, [np.take_along_axis(v, order, 1) for v in vals_list]


def generate_operator(X, Y, p, n, polydeg, groups=None, ops=("E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"), mode=0,
                      brute=False, want_cond=False):
    """src/generate_operator.jl:29-190 (groups=None) / :192-352 (groups=(idx_in, idx_bc, idx_bc_g)).
    Returns (colind [M,n] int64 in stencil order, vals [nops,M,n]) (+ cond1 [N])."""
    X = np.ascontiguousarray(X, np.float64)
    Y = np.ascontiguousarray(Y, np.float64)
    if groups is None:
        idx, _ = knn(X, X, n, brute=brute)
        center = knn(X, Y, 1, brute=brute)[0][:, 0]
    else:
        idx, center, _, _ = calculateneighbors(X, Y, n, *groups, brute=brute)
    res = weights(X, Y, idx, center, p, n, polydeg, op_table(X.shape[1], ops), mode, 0, want_cond)
    colind = idx[center]
    if want_cond:
        return colind, res[0], res[1]
    return colind, res


def generate_operator_collocated(X, p, n, polydeg, ops=("E", "Dx", "Dy", "Dxx", "Dyy", "Dxy"), mode=0, want_cond=False):
    """src/generate_operator.jl:354-491 / src/hyperviscosity_operator.jl:314-440: the legacy 4-argument methods."""
    X = np.ascontiguousarray(X, np.float64)
    idx, _ = knn(X, X, n)
    center = np.arange(len(X), dtype=np.int64)
    res = weights(X, X, idx, center, p, n, polydeg, op_table(X.shape[1], ops), mode, 1, want_cond)
    if want_cond:
        return idx, res[0], res[1]
    return idx, res


def hyperviscosity_operator(K, X, Y, p, n, polydeg, groups=None, mode=0, brute=False):
    """src/hyperviscosity_operator.jl:26-175: (Dxk, Dyk[, Dzk]) = d^K/dx_a^K per axis."""
    d = np.asarray(X).shape[1]
    return generate_operator(X, Y, p, n, polydeg, groups, tuple(("Dk", a, K) for a in range(d)), mode, brute)


def spmv(colind, vals, x, alpha=1.0, beta=0.0, y=None):
    colind = np.ascontiguousarray(colind, np.int64)
    vals = np.ascontiguousarray(vals, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    M, n = colind.shape
    y = np.zeros(M) if y is None else np.ascontiguousarray(y, np.float64)
    lib().orc_spmv(C.c_int64(M), n, _p(colind, C.c_int64), _p(vals, C.c_double), _p(x, C.c_double),
                   C.c_double(alpha), C.c_double(beta), _p(y, C.c_double))
    return y


def spmv_t(colind, vals, v, N, alpha=1.0, beta=0.0, y=None):
    colind = np.ascontiguousarray(colind, np.int64)
    vals = np.ascontiguousarray(vals, np.float64)
    v = np.ascontiguousarray(v, np.float64)
    M, n = colind.shape
    y = np.zeros(N) if y is None else np.ascontiguousarray(y, np.float64)
    lib().orc_spmv_t(C.c_int64(M), C.c_int64(N), n, _p(colind, C.c_int64), _p(vals, C.c_double), _p(v, C.c_double),
                     C.c_double(alpha), C.c_double(beta), _p(y, C.c_double))
    return y


def rhs_advdiff(colind, E, Dx, Dy, Dxx, Dyy, Dxk, Dyk, alpha, ux, uy, gamma, u, N=None):
    """du of examples/adv_diff_test.jl:151-152 (before the ghost-node updates)."""
    colind = np.ascontiguousarray(colind, np.int64)
    M, n = colind.shape
    N = M if N is None else N
    arrs = [np.ascontiguousarray(a, np.float64) for a in (E, Dx, Dy, Dxx, Dyy, Dxk, Dyk)]
    u = np.ascontiguousarray(u, np.float64)
    du = np.zeros(N)
    work = np.zeros(M)
    lib().orc_rhs_advdiff(C.c_int64(M), C.c_int64(N), n, _p(colind, C.c_int64), *[_p(a, C.c_double) for a in arrs],
                          C.c_double(alpha), C.c_double(ux), C.c_double(uy), C.c_double(gamma), _p(u, C.c_double),
                          _p(du, C.c_double), _p(work, C.c_double))
    return du
