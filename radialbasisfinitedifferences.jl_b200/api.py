"""Host-side mirror of the reference's exported Julia API for the hot path, bound to librbffd.so.

Names, argument order and meaning follow src/RadialBasisFiniteDifferences.jl:25-75:

    generate_operator(X, Y, p, n, polydeg)                         src/generate_operator.jl:29
    generate_operator(X, Y, p, n, polydeg, X_idx_in, X_idx_bc, X_idx_bc_g, Y_idx_in, Y_idx_bc, Y_idx_bc_g)   :192
    hyperviscosity_operator(k_deriv, X, Y, p, n, polydeg[, six index sets])   src/hyperviscosity_operator.jl:26,177
    calculateneighbors(X, Y, n, X_idx_in, X_idx_bc, X_idx_bc_g, Y_idx_in, Y_idx_bc, Y_idx_bc_g)   src/calculateneighbors.jl:1

Differences forced by the host language: indices are 0-based, `X`/`Y` are (N, d) float64 arrays (the memory
layout of Vector{SVector{d,Float64}}), index sets are `range`s / integer arrays, and sparse results are
scipy.sparse.csc_matrix (the SparseMatrixCSC of the reference).  The Julia shim (julia/RBFFDB200.jl) keeps
the 1-based originals.  `Operator` is the device-resident handle for the time loop (operators never leave HBM).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import AdvDiffParams, Options, RbffdError

_DERIV_NAMES = {"E": (0, 0, 0), "Dx": (1, 0, 0), "Dy": (0, 1, 0), "Dz": (0, 0, 1), "Dxx": (2, 0, 0), "Dyy": (0, 2, 0),
                "Dzz": (0, 0, 2), "Dxy": (1, 1, 0), "Dxz": (1, 0, 1), "Dyz": (0, 1, 1)}
REFERENCE_OPS = ("E", "Dx", "Dy", "Dxx", "Dyy", "Dxy")     # return order of generate_operator.jl:189


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _coords(A, name):
    A = np.ascontiguousarray(A, dtype=np.float64)
    if A.ndim != 2 or A.shape[1] not in (1, 2, 3):
        raise ValueError(f"{name} must be an (N, d) array with d in 1..3, got shape {A.shape}")
    return A


def make_ops(dim, ops):
    """operator names -> [[kind, a0, a1, a2], ...].  Names: E, Dx, ..., Dyz, 'Lap', ('Dk', axis, K), (a0, a1, a2)."""
    rows = []
    for o in ops:
        if isinstance(o, str) and o == "Lap":
            rows.append((_lib.OP_LAPLACE, 0, 0, 0))
        elif isinstance(o, str):
            rows.append((_lib.OP_DERIV,) + _DERIV_NAMES[o])
        elif len(o) == 3 and o[0] == "Dk":
            al = [0, 0, 0]
            al[int(o[1])] = int(o[2])
            rows.append((_lib.OP_DERIV, *al))
        else:
            al = list(o) + [0] * (3 - len(o))
            rows.append((_lib.OP_DERIV, *[int(v) for v in al]))
    for r in rows:
        if any(r[1 + a] != 0 for a in range(dim, 3)):
            raise ValueError(f"operator {r} differentiates along an axis >= dim={dim}")
    return rows


def make_options(dim, p, n, polydeg, ops, index_base=0, sort_columns=False, kernel=0, variant=0, index_width=64):
    rows = make_ops(dim, ops)
    if not 1 <= len(rows) <= _lib.MAX_OPS:
        raise ValueError(f"between 1 and {_lib.MAX_OPS} operators per call")
    o = Options()
    o.dim, o.p, o.polydeg, o.n, o.nops = int(dim), int(p), int(polydeg), int(n), len(rows)
    for i, r in enumerate(rows):
        for j in range(4):
            o.ops[i][j] = r[j]
    o.index_base, o.sort_columns, o.kernel = int(index_base), int(bool(sort_columns)), int(kernel)
    o.variant = int(variant)
    o.index_width = int(index_width)
    return o


def groups_from_index_sets(N, X_idx_in, X_idx_bc, X_idx_bc_g):
    """index sets of src/processmesh.jl:87,174,183 -> per-node group code (0 interior, 1+2b boundary b, 2+2b ghost b)."""
    g = np.zeros(N, np.int32)
    for b, r in enumerate(X_idx_bc):
        g[np.asarray(r, dtype=np.int64)] = 1 + 2 * b
    for b, r in enumerate(X_idx_bc_g):
        g[np.asarray(r, dtype=np.int64)] = 2 + 2 * b
    return g


class Context:
    """One per (device, stream).  Mirrors nothing in the reference (it has no device state)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._L = _lib.lib()
        h = C.c_void_p()
        rc = self._L.rbffd_create(int(device), C.byref(h))
        if rc != 0:
            raise RbffdError(rc, self._L.rbffd_last_error(None).decode())
        self._h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def _check(self, rc):
        if rc != 0:
            raise RbffdError(rc, self._L.rbffd_last_error(self._h).decode())

    def set_stream(self, stream: int):
        self._check(self._L.rbffd_set_stream(self._h, C.c_void_p(int(stream))))

    @property
    def stream(self) -> int:
        """the cudaStream_t (as an int) the context launches on"""
        s = C.c_void_p()
        self._check(self._L.rbffd_get_stream(self._h, C.byref(s)))
        return s.value or 0

    def reset_stream(self):
        """back to the context's own non-blocking stream"""
        self._check(self._L.rbffd_reset_stream(self._h))

    def synchronize(self):
        self._check(self._L.rbffd_synchronize(self._h))

    def timings(self):
        t = (C.c_double * 8)()
        self._check(self._L.rbffd_timings(self._h, t, 8))
        return dict(zip(("binning", "knn", "nearest", "weights", "sort"), list(t)[:5]))

    def close(self):
        if getattr(self, "_h", None):
            self._L.rbffd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device-pointer entry points (pointers are ints, e.g. torch.Tensor.data_ptr()) ----
    def knn_device(self, X_ptr, N, dim, k, idx_out_ptr, Q_ptr=None, NQ=0, xgroup_ptr=None, qgroup_ptr=None, d2_out_ptr=None):
        self._check(self._L.rbffd_knn_device(self._h, X_ptr, N, dim, Q_ptr, NQ, k, xgroup_ptr, qgroup_ptr, idx_out_ptr, d2_out_ptr))

    def weights_device(self, opts, X_ptr, N, stencils_ptr, colind_ptr, vals_ptr, Y_ptr=None, M=None, center_ptr=None, NS=0):
        self._check(self._L.rbffd_weights_device(self._h, C.byref(opts), X_ptr, N, Y_ptr, N if M is None else M,
                                                 stencils_ptr, NS, center_ptr, colind_ptr, vals_ptr))

    def stencils_device(self, X_ptr, N, dim, n, stencils_ptr, center_ptr=None, Y_ptr=None, M=None, xgroup_ptr=None):
        self._check(self._L.rbffd_stencils_device(self._h, X_ptr, N, dim, Y_ptr, N if M is None else M, n, xgroup_ptr,
                                                  stencils_ptr, center_ptr))

    def operator_from_device(self, M, N, n, nmat, colind_ptr, vals_ptr):
        h = C.c_void_p()
        self._check(self._L.rbffd_operator_from_device(self._h, M, N, n, nmat, colind_ptr, vals_ptr, C.byref(h)))
        return Operator(self, h)

    def launch_count(self) -> int:
        return int(self._L.rbffd_launch_count(self._h))

    def measure_fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self._check(self._L.rbffd_measure_fp64_peak(self._h, C.byref(a), C.byref(b)))
        return {"dfma_tflops": a.value, "dmma_tflops": b.value}

    def operator_generate(self, opts, X_ptr, N, Y_ptr=None, M=None, xgroup_ptr=None):
        h = C.c_void_p()
        self._check(self._L.rbffd_operator_generate(self._h, C.byref(opts), X_ptr, N, Y_ptr, N if M is None else M,
                                                    xgroup_ptr, C.byref(h)))
        return Operator(self, h)

    def jittered_lattice_device(self, dim, g, seed, first, count, out_ptr):
        self._check(self._L.rbffd_jittered_lattice_device(self._h, dim, g, seed, first, count, out_ptr))

    def gather_device(self, src_ptr, index_ptr, count, dst_ptr):
        self._check(self._L.rbffd_gather_device(self._h, src_ptr, index_ptr, count, dst_ptr))

    def scatter_add_device(self, src_ptr, index_ptr, count, dst_ptr):
        self._check(self._L.rbffd_scatter_add_device(self._h, src_ptr, index_ptr, count, dst_ptr))

    def stage_update_device(self, N, a, u_ptr, b, x_ptr, dt, du_ptr, out_ptr):
        """out = a*u + b*(x + dt*du): the Shu-Osher stage combination of an SSP-RK scheme (element-wise; out may alias u or x)"""
        self._check(self._L.rbffd_stage_update_device(self._h, N, a, u_ptr, b, x_ptr, dt, du_ptr, out_ptr))


class Operator:
    """Device-resident operator set (shared sparsity pattern, fixed row length)."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle
        M, N, n, nm = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
        ctx._check(ctx._L.rbffd_operator_info(handle, C.byref(M), C.byref(N), C.byref(n), C.byref(nm)))
        self.M, self.N, self.n, self.nmat = M.value, N.value, n.value, nm.value

    @classmethod
    def from_host(cls, ctx: Context, colind, vals, N, index_base=0):
        colind = np.ascontiguousarray(colind, np.int64)
        vals = np.ascontiguousarray(vals, np.float64)
        M, n = colind.shape
        vals = vals.reshape(-1, M, n)
        h = C.c_void_p()
        ctx._check(ctx._L.rbffd_operator_from_host(ctx._h, M, N, n, vals.shape[0], _ptr(colind), index_base, _ptr(vals), C.byref(h)))
        return cls(ctx, h)

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._L.rbffd_operator_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pointers(self, which=0):
        ci, va = C.c_void_p(), C.c_void_p()
        self.ctx._check(self.ctx._L.rbffd_operator_pointers(self._h, which, C.byref(ci), C.byref(va)))
        return ci.value, va.value

    def to_host(self, index_base=0):
        colind = np.empty((self.M, self.n), np.int64)
        vals = np.empty((self.nmat, self.M, self.n), np.float64)
        self.ctx._check(self.ctx._L.rbffd_operator_to_host(self._h, index_base, _ptr(colind), _ptr(vals)))
        return colind, vals

    # host vectors
    def spmv(self, which, x, alpha=1.0, beta=0.0, y=None):
        x = np.ascontiguousarray(x, np.float64)
        if x.shape != (self.N,):
            raise ValueError(f"DimensionMismatch: x has shape {x.shape}, operator has {self.N} columns")
        y = np.zeros(self.M) if y is None else np.ascontiguousarray(y, np.float64)
        self.ctx._check(self.ctx._L.rbffd_spmv_host(self._h, which, alpha, _ptr(x), beta, _ptr(y)))
        return y

    def spmv_t(self, which, v, alpha=1.0, beta=0.0, y=None):
        v = np.ascontiguousarray(v, np.float64)
        if v.shape != (self.M,):
            raise ValueError(f"DimensionMismatch: v has shape {v.shape}, operator has {self.M} rows")
        y = np.zeros(self.N) if y is None else np.ascontiguousarray(y, np.float64)
        self.ctx._check(self.ctx._L.rbffd_spmv_t_host(self._h, which, alpha, _ptr(v), beta, _ptr(y)))
        return y

    def rhs_advdiff(self, u, params: AdvDiffParams):
        u = np.ascontiguousarray(u, np.float64)
        du = np.empty(self.N)
        self.ctx._check(self.ctx._L.rbffd_rhs_advdiff_host(self._h, C.byref(params), _ptr(u), _ptr(du)))
        return du

    # device pointers
    def spmv_device(self, which, x_ptr, y_ptr, alpha=1.0, beta=0.0):
        self.ctx._check(self.ctx._L.rbffd_spmv_device(self._h, which, alpha, x_ptr, beta, y_ptr))

    def spmv_t_device(self, which, v_ptr, y_ptr, alpha=1.0, beta=0.0):
        self.ctx._check(self.ctx._L.rbffd_spmv_t_device(self._h, which, alpha, v_ptr, beta, y_ptr))

    def spmv_multi_device(self, which, coef, x_ptr, y_ptr):
        w = (C.c_int32 * len(which))(*which)
        c = (C.c_double * len(coef))(*coef)
        self.ctx._check(self.ctx._L.rbffd_spmv_multi_device(self._h, len(which), w, c, x_ptr, y_ptr))

    def combine_device(self, which, coef, vals_out_ptr):
        """vals_out = sum_i coef[i] * D[which[i]] over the shared pattern (constant-coefficient alpha*Dxx + ... done once)"""
        w = (C.c_int32 * len(which))(*which)
        c = (C.c_double * len(coef))(*coef)
        self.ctx._check(self.ctx._L.rbffd_operator_combine_device(self._h, len(which), w, c, vals_out_ptr))

    def rhs_advdiff_device(self, u_ptr, du_ptr, params: AdvDiffParams):
        self.ctx._check(self.ctx._L.rbffd_rhs_advdiff_device(self._h, C.byref(params), u_ptr, du_ptr))

    def rhs_advdiff_stage_device(self, x_ptr, a, u_ptr, b, dt, out_ptr, params: AdvDiffParams):
        """out = a*u + b*(x + dt*cons_sys(x)) (ONE launch on the collocated path)"""
        self.ctx._check(self.ctx._L.rbffd_rhs_advdiff_stage_device(self._h, C.byref(params), x_ptr, a, u_ptr, b, dt, out_ptr))

    def spmv_stage_device(self, which, coef, x_ptr, a, u_ptr, b, dt, out_ptr):
        """out = a*u + b*(x + dt * sum_i coef[i] D[which[i]] x): one SSP-RK stage as ONE launch (rows = nodes)"""
        w = (C.c_int32 * len(which))(*which)
        c = (C.c_double * len(coef))(*coef)
        self.ctx._check(self.ctx._L.rbffd_spmv_stage_device(self._h, len(which), w, c, x_ptr, a, u_ptr, b, dt, out_ptr))


class BoundaryConditions:
    """Ghost-node updates of cons_sys (examples/adv_diff_test.jl:118-141,162-176), device resident.

    boundaries: list of dicts in application order, each {"bc": indices, "ghost": indices, and either
    "matrix": index of D_b in the operator (ghost values solve (D_b u)[bc] = 0) or "value": Dirichlet value}."""

    def __init__(self, op: Operator, boundaries, index_base=0):
        self.op = op
        nb = len(boundaries)
        kind = np.array([1 if "matrix" in b else 0 for b in boundaries], np.int32)
        which = np.array([b.get("matrix", 0) for b in boundaries], np.int32)
        value = np.array([b.get("value", 0.0) for b in boundaries], np.float64)
        bcs = [np.asarray(b["bc"], np.int64).ravel() for b in boundaries]
        ghs = [np.asarray(b["ghost"], np.int64).ravel() for b in boundaries]
        for a, g in zip(bcs, ghs):
            if a.shape != g.shape:
                raise ValueError("DimensionMismatch: every boundary needs as many ghost nodes as boundary nodes")
        ptr = np.concatenate([[0], np.cumsum([len(a) for a in bcs])]).astype(np.int64)
        bc_idx = np.ascontiguousarray(np.concatenate(bcs)) if nb else np.zeros(0, np.int64)
        gh_idx = np.ascontiguousarray(np.concatenate(ghs)) if nb else np.zeros(0, np.int64)
        h = C.c_void_p()
        op.ctx._check(op.ctx._L.rbffd_bc_create(op._h, nb, _ptr(kind), _ptr(which), _ptr(value), _ptr(ptr), _ptr(bc_idx),
                                                _ptr(gh_idx), index_base, C.byref(h)))
        self._h = h

    def apply(self, u):
        u = np.ascontiguousarray(u, np.float64)
        self.op.ctx._check(self.op.ctx._L.rbffd_bc_apply_host(self._h, _ptr(u)))
        return u

    def apply_device(self, u_ptr):
        self.op.ctx._check(self.op.ctx._L.rbffd_bc_apply_device(self._h, u_ptr))

    def close(self):
        if getattr(self, "_h", None):
            self.op.ctx._L.rbffd_bc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def bind_to_gpu_numa(device: int = 0) -> int:
    """Restrict this process to the CPU cores next to GPU `device` (NVML's CPU affinity of the device, intersected with the
    current mask): pinned staging buffers are then first-touched on that NUMA node and the widening threads of
    rbffd_generate_operator_host stay next to the PCIe root they drain.  One process per GPU: call it before creating the
    Context.  Returns the number of hardware threads left to the process (0: NVML unavailable, nothing changed)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(visible.split(",")[device]) if visible and all(t.strip().isdigit() for t in visible.split(",")) else device
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return 0
        # one process per GPU: the ranks that share this NUMA node split its cores between them
        lw = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        lr = int(os.environ.get("LOCAL_RANK", "0"))
        if lw > 1:
            same = []
            for r in range(lw):
                try:
                    hr = pynvml.nvmlDeviceGetHandleByIndex(int(visible.split(",")[r]) if visible and all(t.strip().isdigit() for t in visible.split(",")) else r)
                    wr = pynvml.nvmlDeviceGetCpuAffinity(hr, (ncpu + 63) // 64)
                    if list(wr) == list(words):
                        same.append(r)
                except Exception:
                    pass
            if lr in same and len(same) > 1:
                order = sorted(cpus)
                share = [c for i, c in enumerate(order) if i % len(same) == same.index(lr)]
                if share:
                    cpus = set(share)
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _to_csc(colind, vals, shape_mode, N):
    import scipy.sparse as sp
    M, n = colind.shape
    ncols = N if shape_mode == "full" else (int(colind.max()) + 1 if colind.size else 0)   # sparse(I,J,V): (max I, max J)
    indptr = np.arange(0, M * n + 1, n)
    return [sp.csr_matrix((v.ravel(), colind.ravel(), indptr), shape=(M, ncols)).tocsc() for v in vals]


def generate_raw(X, Y, p, n, polydeg, ops=REFERENCE_OPS, groups=None, ctx=None, sort_columns=False, kernel=0, variant=0,
                 index_base=0, index_width=64):
    """colind [M, n] int64 (int32 with index_width=32; stencil order unless sort_columns; index_base 0, or 1 as the Julia shim
    asks for), vals [nops, M, n]: the fixed-row CSR the kernels write."""
    ctx = ctx or default_context()
    X = _coords(X, "X")
    Y = X if Y is None else _coords(Y, "Y")
    if Y.shape[1] != X.shape[1]:
        raise ValueError("DimensionMismatch: X and Y have different dimensions")
    N, dim = X.shape
    M = Y.shape[0]
    opts = make_options(dim, p, n, polydeg, ops, index_base, sort_columns, kernel, variant, index_width)
    colind = np.empty((M, n), np.int32 if index_width == 32 else np.int64)
    vals = np.empty((opts.nops, M, n), np.float64)
    g = None if groups is None else np.ascontiguousarray(groups, np.int32)
    if g is not None and g.shape != (N,):
        raise ValueError("DimensionMismatch: one group code per X node")
    ctx._check(ctx._L.rbffd_generate_operator_host(ctx._h, C.byref(opts), _ptr(X), N, None if Y is X else _ptr(Y), M,
                                                   _ptr(g), _ptr(colind), _ptr(vals)))
    return colind, vals


def generate_operator(X, Y, p, n, polydeg, X_idx_in=None, X_idx_bc=None, X_idx_bc_g=None, Y_idx_in=None, Y_idx_bc=None,
                      Y_idx_bc_g=None, *, ctx=None, shape="reference"):
    """-> (E, Dx, Dy, Dxx, Dyy, Dxy) as csc_matrix.  generate_operator.jl:29 (5 args) / :192 (11 args)."""
    X = _coords(X, "X")
    groups = None
    if X_idx_bc is not None:
        groups = groups_from_index_sets(X.shape[0], X_idx_in, X_idx_bc, X_idx_bc_g)   # Y_idx_* unused, as in the reference
    if X.shape[1] != 2:
        raise ValueError("the reference API is 2-D; use generate_raw(..., ops=...) for 3-D operator sets")
    colind, vals = generate_raw(X, Y, p, n, polydeg, REFERENCE_OPS, groups, ctx)
    return tuple(_to_csc(colind, vals, shape, X.shape[0]))


def generate_operator_collocated(X, p, n, polydeg, *, ctx=None, shape="reference"):
    """-> (E, Dx, Dy, Dxx, Dyy, Dxy): the legacy 4-argument method generate_operator(X, p, n, polydeg)
    (generate_operator.jl:354-491): unscaled stencils, centre at (eps, eps), RBF rows at X_j - x_c."""
    X = _coords(X, "X")
    colind, vals = generate_raw(X, None, p, n, polydeg, REFERENCE_OPS, None, ctx, variant=1)
    return tuple(_to_csc(colind, vals, shape, X.shape[0]))


def hyperviscosity_operator_collocated(k_deriv, X, p, n, polydeg, *, ctx=None, shape="reference"):
    """-> (Dxk, Dyk): the legacy method hyperviscosity_operator(K, X, p, n, polydeg) (hyperviscosity_operator.jl:314-440)."""
    X = _coords(X, "X")
    ops = [("Dk", a, int(k_deriv)) for a in range(X.shape[1])]
    colind, vals = generate_raw(X, None, p, n, polydeg, ops, None, ctx, variant=1)
    return tuple(_to_csc(colind, vals, shape, X.shape[0]))


def hyperviscosity_operator(k_deriv, X, Y, p, n, polydeg, X_idx_in=None, X_idx_bc=None, X_idx_bc_g=None, Y_idx_in=None,
                            Y_idx_bc=None, Y_idx_bc_g=None, *, ctx=None, shape="reference"):
    """-> (Dxk, Dyk[, Dzk]): d^K/dx_a^K per axis.  hyperviscosity_operator.jl:26 / :177."""
    X = _coords(X, "X")
    groups = None
    if X_idx_bc is not None:
        groups = groups_from_index_sets(X.shape[0], X_idx_in, X_idx_bc, X_idx_bc_g)
    ops = [("Dk", a, int(k_deriv)) for a in range(X.shape[1])]
    colind, vals = generate_raw(X, Y, p, n, polydeg, ops, groups, ctx)
    return tuple(_to_csc(colind, vals, shape, X.shape[0]))


def calculateneighbors(X, Y, n, X_idx_in=None, X_idx_bc=None, X_idx_bc_g=None, Y_idx_in=None, Y_idx_bc=None,
                       Y_idx_bc_g=None, *, ctx=None):
    """-> (idxs_x [N, n], idxs_y_x [M, 1], dists_x [N, n], dists_y_x [M, 1]).  calculateneighbors.jl:1-97.
    With X_idx_bc=None this is the unmasked inline search of generate_operator.jl:43-47."""
    ctx = ctx or default_context()
    X = _coords(X, "X")
    Y = X if Y is None else _coords(Y, "Y")
    N, dim = X.shape
    M = Y.shape[0]
    g = None
    if X_idx_bc is not None:
        g = groups_from_index_sets(N, X_idx_in, X_idx_bc, X_idx_bc_g)
    idx = np.empty((N, n), np.int64)
    idy = np.empty((M, 1), np.int64)
    dx = np.empty((N, n))
    dy = np.empty((M, 1))
    ctx._check(ctx._L.rbffd_calculateneighbors_host(ctx._h, _ptr(X), N, None if Y is X else _ptr(Y), M, dim, int(n), _ptr(g), 0,
                                                    _ptr(idx), _ptr(idy), _ptr(dx), _ptr(dy)))
    return idx, idy, dx, dy
