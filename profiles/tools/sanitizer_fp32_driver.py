"""compute-sanitizer target: the fp32 instantiation of the warp DMMA kernel (n=32, m=8) + the fp32 adjoint, tiny batch."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: F401  (sets up sys.path for the package)
import numpy as np
import differentiable_lqr as dl
T, B, n, m = 5, 6, 32, 8
s = n + m
rs = np.random.RandomState(4)
L = 0.3 * rs.randn(T, B, s, s)
C = L @ np.transpose(L, (0, 1, 3, 2)) + np.eye(s)
F = np.repeat(np.concatenate((0.9 * np.eye(n) + 0.05 * rs.randn(B, n, n), rs.randn(B, n, m)), axis=2)[None], T - 1, axis=0)
node = dl.DiffLqr(T, B, n, m, dtype=np.float32)
x, u = node.apply_numpy(rs.randn(B, n), C, rs.randn(T, B, s), F, 0.1 * rs.randn(T - 1, B, n))
gr = node.backward_numpy(rs.randn(T, B, n), rs.randn(T, B, m))
print("sanitizer fp32 driver ok", x.dtype, float(np.abs(u).max()))
