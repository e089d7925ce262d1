"""Least-squares solve  min_u ||D u - f||_2  on the device, for the oversampled (M > N) collocation systems the reference's
tests assemble and solve with a sparse QR (`u = D \\ f`, test/poisson_test.jl:121, test/mesh_import_test.jl:147).

CGLS (conjugate gradients on the normal equations, never forming D'D) with a column-norm (Jacobi) preconditioner: every
iteration is one `D*p` and one `D'*r` through the library's SpMV kernels; the vectors are torch tensors on the same stream.
The scalars alpha, beta stay on the device; the host looks at the residual of the normal equations every `check` iterations.
"""
from __future__ import annotations


def mixed_row_operator(ctx, op, row_terms, M):
    """One matrix over the operator's shared pattern whose row blocks are different combinations of the operator's matrices:
    row_terms = [(rows (torch int64 index tensor), [(which, coefficient (float or per-row tensor)), ...]), ...].  This is the
    `D[Y_idx_in,:] = Dxx + Dyy; D[Y_idx_neumann,:] = nx.*Dx + ny.*Dy; D[Y_idx_dirichlet,:] = E` assembly of
    test/poisson_test.jl:76-79 with the row scalings of :110-118 folded into the coefficients.
    Returns (new operator with one matrix, its value tensor [1, M, n] — keep it alive as long as the operator)."""
    import torch
    n, nmat = op.n, op.nmat
    ci_ptr, v0_ptr = op.pointers(0)

    class _Raw:
        __cuda_array_interface__ = {"shape": (nmat, M, n), "typestr": "<f8", "data": (v0_ptr, False), "version": 3, "strides": None}
    vals = torch.as_tensor(_Raw(), device=torch.device("cuda", ctx.device))
    out = torch.zeros((1, M, n), dtype=torch.float64, device=vals.device)
    for rows, terms in row_terms:
        acc = torch.zeros((len(rows), n), dtype=torch.float64, device=vals.device)
        for which, c in terms:
            c = c if isinstance(c, float) or isinstance(c, int) else c.reshape(-1, 1)
            acc += c * vals[which, rows]
        out[0, rows] = acc
    return ctx.operator_from_device(M, op.N, n, 1, ci_ptr, out.data_ptr()), out


def cgls(op, which, f, tol=1e-12, maxit=100000, check=200, precondition=True):
    """Solve min ||D u - f|| for D = matrix `which` of `op` (M x N, M >= N).  f: torch float64 [M] on the operator's device.
    Stops when ||S D'(f - D u)|| <= tol * ||S D' f|| (S = inverse column norms).  Returns (u [N], iterations, relative residual)."""
    import torch
    M, N, n = op.M, op.N, op.n
    dev = f.device
    # the torch arithmetic below and the library's SpMVs must run on ONE stream: a Context created without stream= owns
    # a non-blocking stream of its own, on which the products would race with the vector updates
    ctx_stream = op.ctx.stream
    op.ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    try:
        return _cgls(op, which, f, tol, maxit, check, precondition)
    finally:
        op.ctx.synchronize()
        op.ctx.set_stream(ctx_stream)


def _cgls(op, which, f, tol, maxit, check, precondition):
    import torch
    M, N, n = op.M, op.N, op.n
    dev = f.device
    if precondition:                                    # column norms of D: D.^2' * 1 through the transposed product
        ci_ptr, v_ptr = op.pointers(which)
        class _Raw:
            __cuda_array_interface__ = {"shape": (M * n,), "typestr": "<f8", "data": (v_ptr, False), "version": 3, "strides": None}
        v = torch.as_tensor(_Raw(), device=dev)
        sq = (v * v).reshape(1, M, n).contiguous()
        op2 = op.ctx.operator_from_device(M, N, n, 1, ci_ptr, sq.data_ptr())
        cn = torch.empty(N, dtype=torch.float64, device=dev)
        ones = torch.ones(M, dtype=torch.float64, device=dev)
        op2.spmv_t_device(0, ones.data_ptr(), cn.data_ptr())
        op.ctx.synchronize()
        op2.close()
        S = torch.where(cn > 0, cn.rsqrt(), torch.ones_like(cn))
    else:
        S = torch.ones(N, dtype=torch.float64, device=dev)
    x = torch.zeros(N, dtype=torch.float64, device=dev)
    r = f.clone()
    s = torch.empty(N, dtype=torch.float64, device=dev)
    q = torch.empty(M, dtype=torch.float64, device=dev)
    sp = torch.empty(N, dtype=torch.float64, device=dev)
    op.spmv_t_device(which, r.data_ptr(), s.data_ptr())
    s *= S
    p = s.clone()
    gamma = torch.dot(s, s)
    g0 = float(gamma)
    if g0 == 0.0:
        return x, 0, 0.0
    it, rel = 0, 1.0
    while it < maxit:
        for _ in range(check):
            torch.mul(S, p, out=sp)
            op.spmv_device(which, sp.data_ptr(), q.data_ptr())
            alpha = gamma / torch.dot(q, q)
            x.addcmul_(p, alpha)
            r.addcmul_(q, -alpha)
            op.spmv_t_device(which, r.data_ptr(), s.data_ptr())
            s *= S
            gn = torch.dot(s, s)
            p.mul_(gn / gamma).add_(s)
            gamma = gn
        it += check
        rel = (float(gamma) / g0) ** 0.5                # the only host synchronisation
        if not rel > tol:
            break
    return S * x, it, rel
