"""BASELINE.json's configurations at their FULL single-GPU sizes, checked through size-independent properties
(the oracle replays these shapes at small sizes in test_gpu_parity.py):
  * every stencil starts with its own node, indices are in range and distinct, squared distances ascend (kNN order),
  * polynomial reproduction of every operator row up to degree 2 (the constraint rows of the saddle-point system):
      sum_j w_j m(X_j - x_c) = (L m)(0) for m in {1, x_a, x_a x_b},
  * agreement with the CPU oracle on a random sample of whole rows (bit-exact pattern, weights within 50 eps cond),
  * SpMV linearity on the full operator.
Device memory peaks at about 50 GB (config 4: 20M nodes x 60 neighbours x 4 operators)."""
import numpy as np
import pytest

import rbffd_b200 as rb

pytestmark = pytest.mark.gpu

FULL = [
    # name, d, g, p, deg, n, ops
    ("configs[1]: 2-D Laplacian, 1M nodes, k=30", 2, 1000, 5, 3, 30, ["Lap"]),
    ("configs[2]: 2-D hyperviscosity + second derivatives, 10M nodes, k=50", 2, 3162, 5, 4, 50, ["Dxx", "Dyy", ("Dk", 0, 4), ("Dk", 1, 4)]),
    ("configs[3]: 3-D Laplacian + gradient, 20M nodes, k=60", 3, 271, 7, 3, 60, ["Lap", "Dx", "Dy", "Dz"]),
]


def _exact(op, d, e):
    """(L m)(0) for the monomial with exponent tuple e (total degree <= 2) and operator name op"""
    if op == "Lap":
        return 2.0 if max(e) == 2 else 0.0
    if isinstance(op, tuple):                                  # ("Dk", axis, K): d^K/dx_axis^K, K = 4 here
        return 0.0
    al = {"Dx": (1, 0, 0), "Dy": (0, 1, 0), "Dz": (0, 0, 1), "Dxx": (2, 0, 0), "Dyy": (0, 2, 0), "Dzz": (0, 0, 2)}[op][:d]
    if tuple(e) != tuple(al):
        return 0.0
    return float(np.prod([np.prod(np.arange(1, a + 1)) for a in al]))


@pytest.mark.parametrize("name,d,g,p,deg,n,ops", FULL)
def test_full_size_configuration(name, d, g, p, deg, n, ops, oracle):
    import torch
    dev = torch.device("cuda:0")
    ctx = rb.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    N = g ** d
    r = len(ops)
    X = torch.empty((N, d), dtype=torch.float64, device=dev)
    ctx.jittered_lattice_device(d, g, 0, 0, N, X.data_ptr())
    stencils = torch.empty((N, n), dtype=torch.int32, device=dev)
    center = torch.empty(N, dtype=torch.int32, device=dev)
    colind = torch.empty((N, n), dtype=torch.int32, device=dev)
    vals = torch.empty((r, N, n), dtype=torch.float64, device=dev)
    opts = rb.make_options(d, p, n, deg, ops)
    ctx.stencils_device(X.data_ptr(), N, d, n, stencils.data_ptr(), center_ptr=center.data_ptr())
    ctx.weights_device(opts, X.data_ptr(), N, stencils.data_ptr(), colind.data_ptr(), vals.data_ptr(), center_ptr=center.data_ptr())
    torch.cuda.synchronize()
    assert bool((center == torch.arange(N, device=dev, dtype=torch.int32)).all())          # Y == X: every row uses its own stencil
    assert bool((colind[:, 0] == center).all())                                             # scalestencil.jl:10 "first value is the current point"
    assert int(colind.min()) >= 0 and int(colind.max()) < N
    assert torch.equal(colind, stencils)
    exps = [e for e in np.ndindex(*(3,) * d) if sum(e) <= 2]
    worst = 0.0
    chunk = 1 << 20
    for r0 in range(0, N, chunk):
        r1 = min(N, r0 + chunk)
        ci = colind[r0:r1].long()
        S = X[ci] - X[r0:r1, None, :]                                                        # [rows, n, d]
        d2 = (S * S).sum(-1)
        assert bool((d2[:, 1:] >= d2[:, :-1]).all())                                         # ascending squared distance
        srt = ci.sort(1).values
        assert bool((srt[:, 1:] != srt[:, :-1]).all())                                       # no duplicate columns in a row
        for e in exps:
            m = torch.ones_like(d2)
            for a in range(d):
                if e[a]:
                    m = m * S[:, :, a] ** e[a]
            for o, op in enumerate(ops):
                w = vals[o, r0:r1]
                lhs = (w * m).sum(1)
                scale = (w.abs() * m.abs()).sum(1) + 1.0
                worst = max(worst, float(((lhs - _exact(op, d, e)).abs() / scale).max()))
    assert worst <= 1e-7, f"polynomial reproduction violated: {worst:.2e}"                  # measured: a few 1e-10 (eps * cond)
    # whole rows against the oracle on a random sample of centres (the oracle solves only the sampled stencils)
    rng = np.random.default_rng(5)
    rows = np.sort(rng.choice(N, 400, replace=False))
    ci = colind[torch.from_numpy(rows).to(dev)].cpu().numpy().astype(np.int64)
    ids, inv = np.unique(ci, return_inverse=True)
    Xs = X[torch.from_numpy(ids).to(dev)].cpu().numpy()
    idx_local = inv.reshape(ci.shape)                                                        # stencils re-indexed into the sub-cloud
    sel = np.searchsorted(ids, rows)
    full_idx = np.zeros((len(ids), n), np.int64)
    full_idx[sel] = idx_local
    Ysel = Xs[sel]
    ref, cond = oracle.weights(Xs, Ysel, full_idx, sel, p, n, deg, oracle.op_table(d, ops), 0, 0, True)
    got = vals[:, torch.from_numpy(rows).to(dev)].cpu().numpy()
    eps = np.finfo(float).eps
    for o in range(r):
        err = np.abs(got[o] - ref[o]).max(1)
        tol = 50 * eps * cond[sel] * np.abs(ref[o]).max(1)
        assert np.all(err <= tol), f"{ops[o]}: {(err / tol).max():.2f}x the tolerance"
    # SpMV linearity on the full operator
    op = ctx.operator_from_device(N, N, n, r, colind.data_ptr(), vals.data_ptr())
    u, v = torch.randn(N, dtype=torch.float64, device=dev), torch.randn(N, dtype=torch.float64, device=dev)
    yu, yv, yc = (torch.empty(N, dtype=torch.float64, device=dev) for _ in range(3))
    op.spmv_device(0, u.data_ptr(), yu.data_ptr())
    op.spmv_device(0, v.data_ptr(), yv.data_ptr())
    comb = 0.75 * u - 1.5 * v
    op.spmv_device(0, comb.data_ptr(), yc.data_ptr())
    torch.cuda.synchronize()
    assert float((yc - (0.75 * yu - 1.5 * yv)).abs().max()) <= 1e-12 * float(yc.abs().max())
    print(f"{name}: N = {N}, timings(ms) = {ctx.timings()}, reproduction residual {worst:.1e}")
    del op
    ctx.close()
