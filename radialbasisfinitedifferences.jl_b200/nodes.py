"""Synthetic node sets of SURVEY.md §8d (host side, NumPy): the same closed forms the device generator
(rbffd_jittered_lattice_device) uses, so a node set can be produced on either side bit-identically."""
from __future__ import annotations

import numpy as np


def _splitmix_uniform(seed: int, lin: np.ndarray, axis: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (lin.astype(np.uint64) * np.uint64(3) + np.uint64(axis + 1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def jittered_lattice(dim: int, g: int, seed: int = 0, first: int = 0, count: int | None = None) -> np.ndarray:
    """Nodes [first, first+count) of the g^dim jittered lattice in [0,1]^dim (x index fastest)."""
    total = g ** dim
    count = total - first if count is None else count
    lin = np.arange(first, first + count, dtype=np.int64)
    if dim == 2:
        c = [lin % g, lin // g]
    else:
        c = [lin % g, (lin // g) % g, lin // (g * g)]
    X = np.empty((count, dim))
    for a in range(dim):
        u = _splitmix_uniform(seed, lin, a)
        X[:, a] = (c[a].astype(np.float64) + 0.5 + 0.5 * (u - 0.5)) / float(g)
    return X


def halton(dim: int, N: int, skip: int = 1) -> np.ndarray:
    """Halton points in [0,1]^dim (bases 2,3,5), the second synthetic distribution of SURVEY.md §8d."""
    bases = (2, 3, 5)[:dim]
    X = np.zeros((N, dim))
    for a, b in enumerate(bases):
        i = np.arange(skip, skip + N, dtype=np.int64)
        f = 1.0
        r = np.zeros(N)
        while np.any(i > 0):
            f /= b
            r += f * (i % b)
            i //= b
        X[:, a] = r
    return X
