"""CPU: the summation order that boxddp_norm_better_kernel (csrc/boxddp_kernels.cuh) implements is numpy's pairwise
summation of a contiguous float64 row (numpy/core/src/umath/loops_utils.h.src).  The device-resident BoxDDP loop is
bit-identical to the host loop only while this holds, so the restatement is pinned against the installed numpy."""
import numpy as np
import pytest


def pairwise(a, lo, n):
    """Statement-by-statement mirror of np_pairwise_sum in boxddp_kernels.cuh."""
    if n < 8:
        r = 0.0
        for i in range(n):
            r += a[lo + i]
        return r
    if n <= 128:
        r = [a[lo + j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] += a[lo + i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res += a[lo + i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise(a, lo, n2) + pairwise(a, lo + n2, n - n2)


@pytest.mark.parametrize("n", [1, 5, 7, 8, 9, 20, 50, 100, 127, 128, 129, 160, 200, 255, 256, 400, 800, 1000])
def test_row_sum_of_squares_matches_numpy_bit_for_bit(n):
    rs = np.random.RandomState(n)
    d = rs.randn(7, n) * np.exp(3 * rs.randn(7, n))          # wide dynamic range: order matters
    sq = d ** 2
    ref = np.sum(sq, axis=1)
    mine = np.array([pairwise(sq[r], 0, n) for r in range(sq.shape[0])])
    assert np.array_equal(ref, mine)


def test_scrambled_reshape_indexing():
    """Row r of transpose(du,(0,2,1)).reshape(B, T*m) (reference mpc_step.py:261-263) holds the flat elements
    q = r*T*m .. (r+1)*T*m - 1 of the [T,m,B]-ordered array: (t, j, b) = (q // (m*B), (q // B) % m, q % B) - the
    index arithmetic of boxddp_norm_better_kernel."""
    T, B, m = 5, 7, 3
    du = np.arange(T * B * m, dtype=np.float64).reshape(T, B, m)
    d = np.transpose(du, (0, 2, 1)).reshape(B, T * m)
    L = T * m
    for r in range(B):
        for q in range(L):
            qa = r * L + q
            b, tj = qa % B, qa // B
            t, j = tj // m, tj % m
            assert d[r, q] == du[t, b, j]
