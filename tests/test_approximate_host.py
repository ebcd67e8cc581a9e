"""CPU: mpc/approximate.py of the package (names and returns of reference mpc/approximate.py:18-54, 77-119) - the
finite-difference Taylor models of Python-callable costs / dynamics against closed forms and the oracle's analytic
pendulum Jacobian.  No GPU: a generic callable never touches the device here."""
import numpy as np

from _helpers import rel_err


def arr(v):
    return np.asarray(getattr(v, "array", v))


def test_approximate_cost_of_a_quadratic_is_the_quadratic():
    from approximate import approximate_cost
    rs = np.random.RandomState(0)
    T, B, n, m = 4, 5, 3, 2
    s = n + m
    L = rs.randn(s, s)
    C = L @ L.T + np.eye(s)
    c = rs.randn(s)
    x, u = rs.randn(T, B, n), rs.randn(T, B, m)
    H, g, cost = approximate_cost(x, u, lambda tau: 0.5 * np.einsum("bi,ij,bj->b", arr(tau), C, arr(tau)) + arr(tau) @ c)
    H, g, cost = arr(H), arr(g), arr(cost)
    assert H.shape == (T, B, s, s) and g.shape == (T, B, s) and cost.shape == (T, B)
    assert np.abs(H - C).max() < 1e-6 * np.abs(C).max()     # second differences, step eps^(1/4): ~1e-7 relative
    # linear term is shifted by -H tau (reference :50): a quadratic gives back its own c, up to the Hessian error
    assert np.abs(g - c).max() < 1e-4
    tau = np.concatenate((x, u), axis=2)
    assert rel_err(cost, 0.5 * np.einsum("tbi,ij,tbj->tb", tau, C, tau) + tau @ c) < 1e-13


def test_approximate_cost_non_quadratic():
    """The reference's own demo cost (approximate.py:71: sqrt(sum tau^2)): gradient tau/|tau|, Hessian (I - nn^T)/|tau|."""
    from approximate import approximate_cost
    rs = np.random.RandomState(1)
    x, u = rs.randn(3, 2, 1) + 2.0, rs.randn(3, 2, 1) + 2.0
    H, g, cost = (arr(v) for v in approximate_cost(x, u, lambda tau: np.sqrt(np.sum(arr(tau) ** 2, axis=1))))
    tau = np.concatenate((x, u), axis=2)
    r = np.linalg.norm(tau, axis=2)
    nrm = tau / r[..., None]
    H_ref = (np.eye(2) - nrm[..., :, None] * nrm[..., None, :]) / r[..., None, None]
    assert np.abs(H - H_ref).max() < 1e-6
    assert np.abs(g - (nrm - np.einsum("tbij,tbj->tbi", H_ref, tau))).max() < 1e-5
    assert rel_err(cost, r) < 1e-14


def test_linearize_dynamics_linear_callable():
    from approximate import linearize_dynamics
    rs = np.random.RandomState(2)
    T, B, n, m = 5, 4, 3, 2
    A = rs.randn(B, n, n + m)
    b = rs.randn(B, n)
    x, u = rs.randn(T, B, n), rs.randn(T, B, m)
    F, f = linearize_dynamics(x, u, lambda xs, us: np.einsum("bij,bj->bi", A, np.concatenate((arr(xs), arr(us)), axis=1)) + b)
    F, f = arr(F), arr(f)
    assert F.shape == (T - 1, B, n, n + m) and f.shape == (T - 1, B, n)
    assert np.abs(F - A[None]).max() < 1e-9 and np.abs(f - b[None]).max() < 1e-8


def test_linearize_dynamics_pendulum_callable_matches_analytic():
    """A numpy pendulum handed over as an opaque callable: re-rolled trajectory (reference :95) + Jacobians agree with
    the closed form the device code implements (oracle/pendulum.py)."""
    from approximate import linearize_dynamics
    from oracle import pendulum as pend
    rs = np.random.RandomState(3)
    T, B = 12, 6
    th = rs.uniform(-np.pi / 2, np.pi / 2, B)
    x0 = np.stack((np.cos(th), np.sin(th), rs.uniform(-1, 1, B)), axis=1)
    u = rs.uniform(-1.5, 1.5, (T, B, 1))
    x = np.zeros((T, B, 3)); x[0] = x0           # only x[0] is read: the callable path re-rolls the trajectory
    F, f = (arr(v) for v in linearize_dynamics(x, u, lambda xs, us: pend.step(arr(xs), arr(us))))
    F_ref, f_ref = pend.linearize(x0, u)
    assert np.abs(F - F_ref).max() < 1e-8 and np.abs(f - f_ref).max() < 1e-8


def test_lindx_passes_through():
    from approximate import linearize_dynamics
    from util import LinDx
    Fm, fm = np.ones((3, 2, 2, 3)), None
    F, f = linearize_dynamics(np.zeros((4, 2, 2)), np.zeros((4, 2, 1)), LinDx(Fm, fm))
    assert F is Fm and f is None


def test_linearize_dynamics_resolves_the_clip_kink_like_the_analytic_jacobian():
    """Controls sitting exactly on the pendulum's torque clip (+-2 = the MPC bounds): the stencil straddles the kink; the
    inclusive slope is taken, as Chainer's F.clip and the analytic Jacobian do (oracle/pendulum.py `inside`)."""
    from approximate import linearize_dynamics
    from oracle import pendulum as pend
    rs = np.random.RandomState(4)
    T, B = 8, 5
    th = rs.uniform(-np.pi / 2, np.pi / 2, B)
    x0 = np.stack((np.cos(th), np.sin(th), rs.uniform(-1, 1, B)), axis=1)
    u = rs.uniform(-1.5, 1.5, (T, B, 1))
    u[1::2, ::2] = 2.0
    u[::3, 1::2] = -2.0
    x = np.zeros((T, B, 3)); x[0] = x0
    F, f = (arr(v) for v in linearize_dynamics(x, u, lambda xs, us: pend.step(arr(xs), arr(us))))
    F_ref, f_ref = pend.linearize(x0, u)
    assert np.abs(F - F_ref).max() < 1e-4 and np.abs(f - f_ref).max() < 1e-4     # one-sided stencils are O(h) accurate
    assert np.abs(F[:, :, 2, 3] - F_ref[:, :, 2, 3]).max() < 1e-6                # d(new dth)/du = 3 dt / (m l^2), not half of it
