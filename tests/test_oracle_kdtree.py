"""The oracle's cKDTree restatement (third-party arithmetic: scipy is not under
/root/reference) checked black-box against the installed scipy: identical `tree.indices`
(libstdc++ nth_element permutation) and identical nearest index on tie-heavy integer data,
which is what decides the block a lifted hit joins (synteny.py:455-473)."""
import numpy as np
import pytest

import oracle

scipy_spatial = pytest.importorskip("scipy.spatial")


def _cases():
    rng = np.random.default_rng(5)
    for n in (1, 2, 15, 16, 17, 33, 100, 1000, 5000, 20000):
        # diagonal-ish anchors with duplicates in one coordinate (ties in nth_element)
        x = rng.integers(0, max(4, n // 2), size=n)
        y = x + rng.integers(-3, 4, size=n)
        a = np.unique(np.stack([x, y], 1), axis=0)
        a = a[rng.permutation(len(a))]
        yield a, rng
    for n in (50, 400):
        # lattice: massive equidistance
        g = np.stack(np.meshgrid(np.arange(0, n, 3), np.arange(0, n, 3)), -1).reshape(-1, 2)
        yield g[rng.permutation(len(g))], rng


def test_kdtree_matches_scipy():
    total = 0
    for a, rng in _cases():
        tree = scipy_spatial.cKDTree(a, leafsize=16)
        lo, hi = a.min() - 12, a.max() + 12
        q = rng.integers(lo, hi + 1, size=(3000, 2))
        for bound in (5, 10, 20):
            d, i = tree.query(q, p=1, distance_upper_bound=bound)
            idx, near, dist = oracle.kdtree(a, q, bound)
            assert np.array_equal(idx, tree.indices), "tree.indices differ (n=%d)" % len(a)
            assert np.array_equal(near, i), "nearest index differs (n=%d, bound=%d)" % (len(a), bound)
            assert np.array_equal(dist, d)
            total += len(q)
    assert total > 0
