// Device-side bookkeeping of the BoxDDP outer loop (reference mpc/box_ddp.py:121-230), so that an iLQR iteration
// needs no host<->device traffic beyond one 32-byte status record:
//   boxddp_norm_better_kernel : full_du_norm of mpc_step.py:261-263 - the reference's transpose(0,2,1).reshape(B, T*m)
//                           mixes batch elements; that quirk is kept, and the row sums follow
//                           numpy's pairwise summation so the values are the ones numpy would produce - plus the
//                           per-element "better" test (box_ddp.py:195-209) and the reductions the global exit tests
//                           need (any improvement, max full_du_norm, OR of the per-element flags, NaN check)
//   boxddp_post_kernel    : best trajectory update, pendulum linearisation of the next nominal point, exit tests
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "mpc_kernels.cuh"

namespace dmpc {

struct BoxDdpStatus {          // one record per iteration, copied to pinned host memory
  int any_better;
  int flags_or;
  int nonfinite;
  int pad;
  unsigned long long max_du_bits;   // max over b of full_du_norm[b] (non-negative: IEEE order == integer order)
  unsigned long long pad2;
};

// numpy's pairwise summation (numpy/core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum) of f(0..n-1).
// Additions are explicitly rounded (add_rn) so that nvcc cannot contract them with the squares into FMAs.
template <typename R, typename Fn>
__device__ R np_pairwise_sum(const Fn& f, int lo, int n) {
  if (n < 8) {
    R res = R(0);
    for (int i = 0; i < n; ++i) res = add_rn(res, f(lo + i));
    return res;
  }
  if (n <= 128) {
    R r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = f(lo + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = add_rn(r[j], f(lo + i + j));
    }
    R res = add_rn(add_rn(add_rn(r[0], r[1]), add_rn(r[2], r[3])), add_rn(add_rn(r[4], r[5]), add_rn(r[6], r[7])));
    for (; i < n; ++i) res = add_rn(res, f(lo + i));
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return add_rn(np_pairwise_sum<R>(f, lo, n2), np_pairwise_sum<R>(f, lo + n2, n - n2));
}

// Loop control that lives on the device, so that the host does not have to synchronise every iteration: after each
// iteration the last CTA of boxddp_post_kernel applies the reference's exit tests (box_ddp.py:184-230) to the reductions
// of boxddp_norm_better_kernel and raises `done`; every kernel of a later iteration returns at once when it sees it.  The host
// enqueues a few iterations at a time and reads this record once per block.
struct BoxDdpCtl {
  int done;              // 1 = an exit test fired (or a non-finite value was seen)
  int status;            // DMPC_BOXDDP_MAX_ITER / _CONVERGED / _NOT_IMPROVED
  int n_iter;            // iterations executed
  int n_not_improved;    // the reference's shared counter (box_ddp.py:184, 203, 226)
  int flags_or;          // OR of the per-element flags over all iterations
  int nonfinite;
  unsigned int ticket;   // CTAs of boxddp_post_kernel that have finished (the last one applies the exit tests)
  int pad;
};

__device__ __forceinline__ void boxddp_decide(BoxDdpStatus* st, BoxDdpCtl* ctl, int iter, double eps, int not_improved_lim,
                                              int nonfinite) {
  ctl->n_iter = iter + 1;
  ctl->flags_or |= st->flags_or;
  int nn = ctl->n_not_improved + 1;
  if (iter > 0 && st->any_better) nn = 0;
  ctl->n_not_improved = nn;
  const double max_du = __longlong_as_double((long long)st->max_du_bits);
  if (nonfinite) { ctl->nonfinite = 1; ctl->done = 1; }
  else if (max_du < eps) { ctl->status = DMPC_BOXDDP_CONVERGED; ctl->done = 1; }                 // box_ddp.py:223-225
  else if (nn > not_improved_lim) { ctl->status = DMPC_BOXDDP_NOT_IMPROVED; ctl->done = 1; }    // :227-229
  st->any_better = 0; st->flags_or = 0; st->nonfinite = 0; st->max_du_bits = 0ull;               // next iteration's reductions
}

// ---- the two bookkeeping launches of a device-resident iteration ------------------------------------------------------
// boxddp_norm_better_kernel (one thread per element): full_du_norm (scrambled, as above), the reference's per-element
// "better" test and best-cost update (box_ddp.py:195-209) -> mask[b], and the batch-wide reductions of the exit tests.
template <typename R>
__global__ void boxddp_norm_better_kernel(int T, int B, int m, int first, R best_cost_eps, const R* u_nom, const R* u_first,
                                          const R* costs, const int* flags, R* du, R* bcosts, R* bdu, unsigned char* mask,
                                          BoxDdpStatus* st, const int* skip) {
  if (skip && *skip) return;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  int better = 0, fl = 0, bad = 0;
  unsigned long long dub = 0ull;
  if (r < B) {
    const int L = T * m;
    auto sq = [&](int q) -> R {
      const long long qa = (long long)r * L + q;
      const int bb = (int)(qa % B);
      const int tj = (int)(qa / B);
      const int t = tj / m, j = tj - t * m;
      const size_t idx = ((size_t)t * B + bb) * m + j;
      const R d = add_rn(u_nom[idx], -u_first[idx]);
      return mul_rn(d, d);
    };
    const R ssum = np_pairwise_sum<R>(sq, 0, L);
    const R d = sizeof(R) == 8 ? (R)sqrt((double)ssum) : (R)sqrtf((float)ssum);
    du[r] = d;
    const R cb = costs[r];
    better = first ? 1 : (cb <= bcosts[r] + best_cost_eps);
    fl = flags ? flags[r] : 0;
    bad = !(d == d) || !(cb == cb);
    dub = (unsigned long long)__double_as_longlong((double)d);
    if (d < R(0) || !(d == d)) dub = 0x7ff8000000000000ull;    // NaN sorts above every finite norm
    if (better) { bcosts[r] = cb; bdu[r] = d; }
    mask[r] = (unsigned char)better;
  }
  const unsigned full = 0xffffffffu;
  better = __any_sync(full, better);
  bad = __any_sync(full, bad);
  for (int o = 16; o > 0; o >>= 1) {
    fl |= __shfl_xor_sync(full, fl, o);
    const unsigned long long other = __shfl_xor_sync(full, dub, o);
    dub = other > dub ? other : dub;
  }
  if ((threadIdx.x & 31) == 0) {
    if (better && !first) atomicOr(&st->any_better, 1);
    if (fl) atomicOr(&st->flags_or, fl);
    if (bad) atomicOr(&st->nonfinite, 1);
    atomicMax(&st->max_du_bits, dub);
  }
}

// boxddp_post_kernel (one thread per (t, b)): best trajectory <- new trajectory where mask[b]; NaN check of the new
// trajectory; pendulum: linearisation of the NEXT iteration's nominal point (x_new, u_new) - the step's accepted rollout
// already is get_traj(u_new) (same pendulum_step, same inputs), so the per-iteration rollout launch of the reference loop
// (box_ddp.py:123) reduces to this pointwise Jacobian; the last CTA to finish applies the exit tests.
constexpr int kPostThreads = 128;

template <typename R>
__global__ void __launch_bounds__(kPostThreads) boxddp_post_kernel(int T, int B, int n, int m, const R* x, const R* u, const unsigned char* mask, R* bx, R* bu,
                                   const R* dyn_params, R* F_lin, BoxDdpStatus* st, BoxDdpCtl* ctl, int iter, double eps,
                                   int not_improved_lim) {
  if (ctl->done) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0;
  if (i < (size_t)T * B) {
    const int b = (int)(i % B);
    const int t = (int)(i / B);
    const bool better = mask[b] != 0;
    for (int k = 0; k < n; ++k) { const R v = x[i * n + k]; bad |= !(v == v); if (better) bx[i * n + k] = v; }
    for (int k = 0; k < m; ++k) { const R v = u[i * m + k]; bad |= !(v == v); if (better) bu[i * m + k] = v; }
  }
  // pendulum (n = 3, m = 1, checked by the C ABI): Jacobian rows of this CTA's (t, b) range are staged in shared memory and
  // written as one contiguous run (a thread's own 12 doubles are 96 bytes apart from its neighbour's: storing them
  // directly costs four partial-sector writes per sector)
  if (F_lin) {
    __shared__ R sF[kPostThreads * 13];                        // row stride 13: conflict-free staging writes
    const size_t total = (size_t)T * B;
    if (i < total) {
      const int t = (int)(i / B);
      R Fl[12];
      if (t < T - 1) {
        R par[5], tau[4], xn[3];
        for (int k = 0; k < 5; ++k) par[k] = dyn_params[k];
        for (int k = 0; k < 3; ++k) { tau[k] = x[i * 3 + k]; xn[k] = x[(i + B) * 3 + k]; }
        tau[3] = u[i];
        pendulum_jacobian<R>(par, tau, xn, Fl, nullptr);
      } else {
        for (int k = 0; k < 12; ++k) Fl[k] = R(0);
      }
      for (int k = 0; k < 12; ++k) sF[threadIdx.x * 13 + k] = Fl[k];
    }
    __syncthreads();
    const size_t i0 = (size_t)blockIdx.x * blockDim.x;
    const size_t lim = (size_t)(T - 1) * B;                      // F_lin has T-1 time rows
    const size_t i1 = i0 + blockDim.x < lim ? i0 + blockDim.x : lim;
    if (i1 > i0) {
      const int cnt = (int)(i1 - i0) * 12;
      for (int k = threadIdx.x; k < cnt; k += blockDim.x) F_lin[i0 * 12 + k] = sF[(k / 12) * 13 + k % 12];
    }
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(&st->nonfinite, 1);
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctl->ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      ctl->ticket = 0u;
      const int nonfinite = atomicOr(&st->nonfinite, 0);      // through L2: the other CTAs' atomics of this launch
      boxddp_decide(st, ctl, iter, eps, not_improved_lim, nonfinite);
    }
  }
}

// ---- warm-start cache of controls in HBM (env_dx/il_exp.py:215-257 keeps train_warmstart[n_samples][T][m] on the host and
// IL_Env.mpc transposes the gathered rows to [T][B][m], il_env.py:113).  take: u[t][b][:] = cache[idx[b]][t][:];
// put: cache[idx[b]][t][:] = u[t][b][:].  One thread per (t, b, j); indices outside [0, n_samples) are skipped.
template <typename R, bool PUT>
__global__ void warmstart_kernel(int T, int B, int m, int n_samples, R* cache, const int* idx, R* u) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)T * B * m) return;
  const int j = (int)(i % m);
  const int b = (int)((i / m) % B);
  const int t = (int)(i / ((size_t)m * B));
  const int sidx = idx[b];
  if (sidx < 0 || sidx >= n_samples) { if (!PUT) u[i] = R(0); return; }
  R* c = cache + ((size_t)sidx * T + t) * m + j;
  if (PUT) *c = u[i]; else u[i] = *c;
}

}  // namespace dmpc
