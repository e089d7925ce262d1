"""Pins the CPU oracle against the known-answer tests the reference itself holds.

The reference has no golden index/weight vectors; what it asserts is
  * test/poisson_test.jl:132        relative l2 error of the oversampled Poisson solve < 0.0027
  * test/hyperviscosity_test.jl:32  hyperviscosity_operator(2, ...) ~ Dxx, Dyy (Frobenius isapprox, rtol sqrt(eps))
Both are replayed here on the reference's own node sets (tests/golden/tominec_fitted.npz).
"""
import numpy as np


from poisson_helper import poisson_error as _poisson_error


def test_poisson_known_answer(tominec, oracle):
    err = _poisson_error(tominec, lambda X, Y, p, n, deg: oracle.generate_operator(X, Y, p, n, deg, mode=0))
    assert err < float(tominec["poisson_threshold"])          # the reference's own assertion
    assert abs(err - 0.0026579) < 2e-6                        # survey probe value


def test_poisson_known_answer_lu_mode(tominec, oracle):
    err = _poisson_error(tominec, lambda X, Y, p, n, deg: oracle.generate_operator(X, Y, p, n, deg, mode=1))
    assert err < float(tominec["poisson_threshold"])


def test_hyperviscosity_k2_equals_second_derivatives(tominec, oracle):
    """test/hyperviscosity_test.jl:13-33 (p=5, polydeg=5, n=42, X == Y)."""
    X = tominec["X"]
    p, polydeg, n = 5, 5, 42
    colind, vals = oracle.generate_operator(X, X, p, n, polydeg)
    c2, vk = oracle.hyperviscosity_operator(2, X, X, p, n, polydeg)
    assert np.array_equal(colind, c2)
    rtol = float(tominec["hyperviscosity_rtol"])
    for got, ref in ((vk[0], vals[3]), (vk[1], vals[4])):
        assert np.linalg.norm(got - ref) <= rtol * max(np.linalg.norm(got), np.linalg.norm(ref))
        assert np.linalg.norm(got - ref) <= 1e-11 * np.linalg.norm(ref)    # probe: 3e-14
