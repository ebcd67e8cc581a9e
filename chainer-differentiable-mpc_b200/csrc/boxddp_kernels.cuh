// Device-side bookkeeping of the BoxDDP outer loop (reference mpc/box_ddp.py:121-230), so that an iLQR iteration
// needs no host<->device traffic beyond one 32-byte status record:
//   scrambled_norm_kernel : full_du_norm of mpc_step.py:261-263 - the reference's transpose(0,2,1).reshape(B, T*m)
//                           mixes batch elements; that quirk is kept, and the row sums follow
//                           numpy's pairwise summation so the values are the ones numpy would produce
//   best_update_kernel    : per-element best trajectory (box_ddp.py:195-209) + the reductions the global exit
//                           tests need (any improvement, max full_du_norm, OR of the per-element flags, NaN check)
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

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
// iteration boxddp_decide_kernel applies the reference's exit tests (box_ddp.py:184-230) to the reductions of
// best_update_kernel and raises `done`; every kernel of a later iteration returns at once when it sees it.  The host
// enqueues a few iterations at a time and reads this record once per block.
struct BoxDdpCtl {
  int done;              // 1 = an exit test fired (or a non-finite value was seen)
  int status;            // DMPC_BOXDDP_MAX_ITER / _CONVERGED / _NOT_IMPROVED
  int n_iter;            // iterations executed
  int n_not_improved;    // the reference's shared counter (box_ddp.py:184, 203, 226)
  int flags_or;          // OR of the per-element flags over all iterations
  int nonfinite;
  int pad[2];
};

__global__ void boxddp_decide_kernel(BoxDdpStatus* st, BoxDdpCtl* ctl, int iter, double eps, int not_improved_lim) {
  if (ctl->done) return;
  ctl->n_iter = iter + 1;
  ctl->flags_or |= st->flags_or;
  int nn = ctl->n_not_improved + 1;
  if (iter > 0 && st->any_better) nn = 0;
  ctl->n_not_improved = nn;
  const double max_du = __longlong_as_double((long long)st->max_du_bits);
  if (st->nonfinite) { ctl->nonfinite = 1; ctl->done = 1; }
  else if (max_du < eps) { ctl->status = DMPC_BOXDDP_CONVERGED; ctl->done = 1; }                 // box_ddp.py:223-225
  else if (nn > not_improved_lim) { ctl->status = DMPC_BOXDDP_NOT_IMPROVED; ctl->done = 1; }    // :227-229
  st->any_better = 0; st->flags_or = 0; st->nonfinite = 0; st->max_du_bits = 0ull;               // next iteration's reductions
}

// out[r] = sqrt(sum_q d[q]^2), q in [r L, (r+1) L), L = T m, where d is (a - b)[T,B,m] read in [T,m,B] order
template <typename R>
__global__ void scrambled_norm_kernel(int T, int B, int m, const R* a, const R* b, R* out, const int* skip) {
  if (skip && *skip) return;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  const int L = T * m;
  auto sq = [&](int q) -> R {
    const long long qa = (long long)r * L + q;
    const int bb = (int)(qa % B);
    const int tj = (int)(qa / B);
    const int t = tj / m, j = tj - t * m;
    const size_t idx = ((size_t)t * B + bb) * m + j;
    const R d = add_rn(a[idx], -b[idx]);
    return mul_rn(d, d);
  };
  const R s = np_pairwise_sum<R>(sq, 0, L);
  out[r] = sizeof(R) == 8 ? (R)sqrt((double)s) : (R)sqrtf((float)s);
}

template <typename R>
__global__ void best_update_kernel(int T, int B, int n, int m, int first, R best_cost_eps, const R* x, const R* u,
                                   const R* costs, const R* du, const int* flags, R* bx, R* bu, R* bcosts, R* bdu,
                                   BoxDdpStatus* st, const int* skip) {
  if (skip && *skip) return;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  int better = 0, fl = 0, bad = 0;
  unsigned long long dub = 0ull;
  if (b < B) {
    const R cb = costs[b];
    better = first ? 1 : (cb <= bcosts[b] + best_cost_eps);
    fl = flags ? flags[b] : 0;
    const R d = du[b];
    bad = !(d == d) || !(cb == cb);
    dub = (unsigned long long)__double_as_longlong((double)d);
    if (d < R(0) || !(d == d)) dub = 0x7ff8000000000000ull;    // NaN sorts above every finite norm
    for (int t = 0; t < T; ++t) {
      for (int i = 0; i < n; ++i) { const R v = x[((size_t)t * B + b) * n + i]; bad |= !(v == v); if (better) bx[((size_t)t * B + b) * n + i] = v; }
      for (int j = 0; j < m; ++j) { const R v = u[((size_t)t * B + b) * m + j]; bad |= !(v == v); if (better) bu[((size_t)t * B + b) * m + j] = v; }
    }
    if (better) { bcosts[b] = cb; bdu[b] = d; }
  }
  // warp-level reduction, then one atomic per warp
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

}  // namespace dmpc
