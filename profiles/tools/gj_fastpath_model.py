"""NumPy model of the compiled-out DMPC_GJ_FASTPATH experiment (lqr_dmma_warp.cuh: gj_step_nopivot): Gauss-Jordan inverse
of an 8x8 matrix without row exchanges, with the acceptance test `|pivot| >= 2^-6 * max_{i>k} |a_ik|` evaluated on the high
32 bits of the doubles exactly as the device code does (abs_hi(p) + (6 << 20) < max abs_hi).  Checks, on the Quu matrices the
Riccati sweep of the bench workload really produces and on adversarial ones, that (a) accepted inverses are as accurate as
LAPACK's pivoted inverse and (b) matrices that need pivoting are flagged for the pivoted fallback.  CPU only."""
import numpy as np

M = 8


def abs_hi(x):
    return (np.float64(x).view(np.uint64) >> np.uint64(32)).astype(np.uint64) & np.uint64(0x7FFFFFFF)


def gj_nopivot(A):
    c = np.concatenate((A.astype(np.float64), np.eye(M)), axis=1)      # columns = lanes 0..15, rows = registers
    bad = False
    for k in range(M):
        pc = c[:, k].copy()
        mx = max([int(abs_hi(pc[i])) for i in range(k + 1, M)] + [0])
        bad |= int(abs_hi(pc[k])) + (6 << 20) < mx
        ck = c[k] / pc[k]
        c[k] = ck
        for i in range(M):
            if i != k:
                c[i] = c[i] - pc[i] * ck
    return c[:, M:], bad


def riccati_quu(rs, n=32, m=8, T=12):
    """Quu_t along a Riccati sweep on the bench's synthetic problem (bench.make_problem_torch in NumPy)."""
    s = n + m
    A = np.eye(n) + 0.2 * rs.randn(n, n)
    A *= min(1.0, 0.95 / np.max(np.abs(np.linalg.eigvals(A))))
    F = np.concatenate((A, rs.randn(n, m)), axis=1)
    L = 0.3 * rs.randn(s, s)
    C = L @ L.T + np.eye(s)
    V = np.zeros((n, n))
    out = []
    for _ in range(T):
        Q = C + F.T @ V @ F
        Quu, Qux = Q[n:, n:], Q[n:, :n]
        out.append(Quu)
        K = -np.linalg.solve(Quu, Qux)
        V = Q[:n, :n] + Q[:n, n:] @ K
    return out


rs = np.random.RandomState(0)
worst, n_ok = 0.0, 0
for trial in range(40):
    for Quu in riccati_quu(rs):
        for pert in (0.0, 0.3):                                          # symmetric, and a non-symmetric C as the tests use
            A = Quu + pert * rs.randn(M, M) * np.sqrt(np.abs(np.diag(Quu)).mean()) * 0.2
            inv, bad = gj_nopivot(A)
            assert not bad, "a workload matrix was sent to the fallback"
            ref = np.linalg.inv(A)
            err = np.max(np.abs(inv - ref)) / np.max(np.abs(ref))
            worst = max(worst, err / (np.linalg.cond(A) * 2.2e-16))
            n_ok += 1
print("accepted %d workload matrices; worst error = %.2f x cond x eps (LAPACK inverse as reference)" % (n_ok, worst))
assert worst < 8.0

flagged = 0
for trial in range(200):
    A = rs.randn(M, M)
    k = rs.randint(M - 1)
    A[k, k] = 1e-9 * rs.randn()                                          # a pivot that partial pivoting would never take
    A[k + 1:, k] += np.sign(A[k + 1:, k]) * 0.5
    A[:k, :] = np.triu(A[:k, :]) if k else A[:k, :]                      # keep the leading steps from disturbing column k
    A[k:, :k] = 0.0
    A[np.arange(k), np.arange(k)] = 1.0 + rs.rand(k)
    _, bad = gj_nopivot(A)
    flagged += int(bad)
print("adversarial matrices flagged for the pivoted fallback: %d / 200" % flagged)
assert flagged == 200
