"""Independent pin of the projected-Newton box QP (mpc/pnqp.py:37-201): on small instances the exact minimiser is found
by enumerating all 3^m active sets (free / at lower / at upper), solving the free block and checking the KKT signs - no
code shared with the path.  PNQP returns x *before* a step shorter than 1e-4 (pnqp.py:139-144, Q4), so it agrees with
the exact solution to that order, and on non-degenerate instances its free set is the exact one.  Oracle only (CPU);
the CUDA kernel is compared bit-for-bit on decisions with the same oracle in the -m gpu tests."""
import itertools

import numpy as np
import pytest

from oracle import pnqp as opnqp


def exact_box_qp(H, q, lo, hi):
    m = len(q)
    best = None
    for assign in itertools.product((0, 1, 2), repeat=m):           # 0 free, 1 at lower, 2 at upper
        a = np.array(assign)
        free = a == 0
        x = np.where(a == 1, lo, np.where(a == 2, hi, 0.0))
        if free.any():
            x[free] = np.linalg.solve(H[np.ix_(free, free)], -(q[free] + H[np.ix_(free, ~free)] @ x[~free]))
        g = H @ x + q
        if np.all(x >= lo - 1e-12) and np.all(x <= hi + 1e-12) and np.all(g[a == 1] >= 0) and np.all(g[a == 2] <= 0):
            # distance of the free controls from the box and of the multipliers from zero: how well defined the set is
            margin = np.min(np.minimum(x[free] - lo[free], hi[free] - x[free])) if free.any() else np.inf
            mult = np.min(np.abs(g[~free])) if (~free).any() else np.inf
            best = (x, free, min(margin, mult))
            break                                                   # strictly convex: the KKT point is unique
    assert best is not None
    return best


@pytest.mark.parametrize("m", [1, 2, 3, 4])
@pytest.mark.parametrize("coupling", ["batch", "element"])
def test_pnqp_matches_exhaustive_active_set_enumeration(m, coupling):
    rs = np.random.RandomState(100 + m)
    B = 24
    L = rs.randn(B, m, m)
    H = L @ np.transpose(L, (0, 2, 1)) + 0.5 * np.eye(m)
    q = 2.0 * rs.randn(B, m)
    lo = -np.abs(rs.randn(B, m)) * 0.7 - 0.05
    hi = np.abs(rs.randn(B, m)) * 0.7 + 0.05
    x, _, free, it = opnqp.pnqp(H, q, lo, hi, None, 20, coupling=coupling)[:4]
    x = np.asarray(x); free = np.asarray(free)
    n_clamped = 0
    for b in range(B):
        xe, fe, margin = exact_box_qp(H[b], q[b], lo[b], hi[b])
        cond = np.linalg.cond(H[b])
        assert np.max(np.abs(x[b] - xe)) < 2e-4 * max(1.0, cond), (b, x[b], xe)
        if margin > 1e-3:                                           # non-degenerate: the active set is defined
            assert np.array_equal(free[b].astype(bool), fe), (b, free[b], fe)
        n_clamped += int((~fe).sum())
    assert n_clamped > 0.15 * B * m                                 # the bounds are really active
