"""CPU checks of the spatial-block planner (rbffd_shard_plan_host: host only, no device): balanced blocks over arbitrary node
sets, explicit block grids including 2x2x2, argument errors.  The shard construction itself needs the device (tests/test_gpu_shard.py)."""
import numpy as np
import pytest

import rbffd_b200 as rb
from rbffd_b200 import sharding


def test_plan_balanced_quantile_blocks(tominec):
    X = tominec["X"]
    for nparts in (1, 2, 3, 4, 8):
        part = sharding.plan(X, nparts)
        cnt = np.bincount(part, minlength=nparts)
        assert cnt.sum() == len(X) and cnt.max() - cnt.min() <= nparts, (nparts, cnt)
    # 2 x 2: every block is a coordinate box (cut along x first, then y inside every x slab)
    part = sharding.plan(X, 4, blocks=(2, 2))
    for b0 in range(2):
        slab = (part // 2) == b0
        other = (part // 2) == 1 - b0
        assert (X[slab, 0].max() <= X[other, 0].min()) or (X[other, 0].max() <= X[slab, 0].min())
        lo, hi = X[slab & (part % 2 == 0), 1], X[slab & (part % 2 == 1), 1]
        assert lo.max() <= hi.min()


def test_plan_3d_2x2x2_and_automatic_grid():
    X = rb.nodes.jittered_lattice(3, 16, seed=3)
    part = sharding.plan(X, 8, blocks=(2, 2, 2))
    assert np.array_equal(np.bincount(part), np.full(8, 512))
    auto = sharding.plan(X, 8)                     # a cube: 2 x 2 x 2 has the smallest cut surface
    assert np.array_equal(np.sort(np.bincount(auto)), np.full(8, 512))
    for b in range(8):
        sel = X[auto == b]
        assert np.all(sel.max(0) - sel.min(0) < 0.55)
    flat = np.column_stack([X[:, 0], X[:, 1], 0.01 * X[:, 2]])          # a plate: cuts avoid the thin axis
    p2 = sharding.plan(flat, 4)
    for b in range(4):
        sel = flat[p2 == b]
        assert sel[:, 2].max() - sel[:, 2].min() > 0.009


def test_plan_argument_errors():
    X = np.random.default_rng(0).random((50, 2))
    with pytest.raises(rb.RbffdError):
        sharding.plan(X, 4, blocks=(3, 1))
    with pytest.raises(rb.RbffdError):
        sharding.plan(X, 0)
    with pytest.raises(rb.RbffdError):
        sharding.plan(X, 17)


def test_numa_binding_is_harmless_without_a_gpu():
    """rb.bind_to_gpu_numa restricts the process to the cores next to its GPU (NVML); without NVML / a device it must change nothing"""
    import os
    before = os.sched_getaffinity(0)
    n = rb.bind_to_gpu_numa(0)
    after = os.sched_getaffinity(0)
    assert isinstance(n, int) and n >= 0
    assert after <= before and len(after) >= 1
    if n == 0:
        assert after == before
    os.sched_setaffinity(0, before)
