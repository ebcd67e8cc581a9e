"""CPU oracle for the LQR / box-DDP hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy (+ torch-CPU LAPACK for batched LU, exactly as the reference does)
restatement of the algorithms in pfnet-research/chainer-differentiable-mpc:

    oracle.lqr       <- lqr/lqr_recursion.py:69-200, lqr/differentiable_lqr.py:78-142
    oracle.pnqp      <- mpc/pnqp.py:26-201
    oracle.mpc       <- mpc/mpc_step.py:70-460, mpc/active_constrained_lqr.py:67-193
    oracle.boxddp    <- mpc/box_ddp.py:93-291
    oracle.pendulum  <- env_dx/pendulum.py:65-102 + analytic Jacobian replacing
                        mpc/approximate.py:77-119 (chainer.grad is not available)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package, and only as the checker / CPU baseline - never on
the product path (which fails loudly when the CUDA library is missing).

Parity pinning: every function here is checked against (a) the golden vectors
stored in the reference's notebooks (Boyd_lqr, one-variable LQR, the PNQP
known-answer test, the LQRnet training trace) and (b) the *unmodified*
reference modules executed under tests/_chainer_stub in the build container;
the generating/validating script is tests/golden/make_golden.py and its
outputs are committed under tests/golden/*.npz.

Switches:
    lu_fp32   True  = literal reference (float32 torch.lu_solve, util.py:522-526)
              False = fp64-clean oracle that the 1e-10 target is defined against
    coupling  'batch'   = literal reference control flow over the whole batch
              'element' = the reference run with n_batch == 1 on every element
"""
