// Common device-side building blocks for the B200 (sm_100a) LQR / box-DDP kernels.
//
// Execution model: a *group* of G cooperating lanes owns one batch
// element for its whole horizon.  G <= 32 -> the group is a slice of a warp and
// synchronises with __syncwarp(mask); G > 32 -> the group is the whole CTA (one
// element per CTA) and synchronises with __syncthreads().  All per-element state
// (Q_t, V_t, F_t tiles, gains) lives in that group's shared-memory region; the
// per-timestep C/c/F/f tiles are streamed with cp.async (LDGSTS) into a
// double-buffered stage straight from the reference's [T][B][...] layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/diffmpc_b200.h"

namespace dmpc {

// status codes / per-element flags come from the public header
enum ElemFlag : int {
  FLAG_QP_NOT_CONVERGED = DMPC_FLAG_QP_NOT_CONVERGED,
  FLAG_NONFINITE = DMPC_FLAG_NONFINITE,
  FLAG_LS_CAPPED = DMPC_FLAG_LS_CAPPED,
  FLAG_BAD_BOUNDS = DMPC_FLAG_BAD_BOUNDS,
};

// ---------------------------------------------------------------- pinned arithmetic
// Explicitly rounded operations: the compiler may neither contract nor re-associate them, so a
// rollout re-evaluated in another kernel reproduces the first one bit for bit (the line search
// relies on cost(alpha -> 0) == cost(nominal), mpc_step.py:196).
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

// ---------------------------------------------------------------- groups
template <int G>
struct Grp {
  int lane;
  unsigned mask;
  __device__ __forceinline__ Grp() {
    if (G <= 32) {
      const int l = threadIdx.x & 31;
      const int base = l & ~(G - 1);
      lane = l - base;
      mask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << base);
    } else {
      lane = threadIdx.x;
      mask = 0xffffffffu;
    }
  }
  __device__ __forceinline__ void sync() const {
    if (G <= 32) __syncwarp(mask); else __syncthreads();
  }
  // lanes of the first warp-slice of the group (used for ballot-based code)
  __device__ __forceinline__ int shift() const { return (G <= 32) ? ((threadIdx.x & 31) & ~(G - 1)) : 0; }
};

// ---------------------------------------------------------------- batch-wide OR (coupling = BATCH)
// The reference's PNQP decides on xp.sum / xp.max over the WHOLE batch (pnqp.py:139-144, 172-187).  One CTA:
// __syncthreads_or.  Batch spread over the CTAs of a thread-block cluster (up to 16): every CTA publishes its OR in its own
// shared memory, a cluster barrier, then everybody reads all ranks' flags through distributed shared memory.  Two flag slots
// used alternately: a slot is rewritten two decisions later, i.e. after a cluster barrier that every reader of its old value
// has passed.  `phase` is per-thread state (identical in all threads of the cluster).
__device__ __forceinline__ unsigned cluster_nctas() { unsigned n; asm("mov.u32 %0, %%cluster_nctarank;" : "=r"(n)); return n; }
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ int batch_or(int pred, unsigned& phase) {
  int v = __syncthreads_or(pred);
  const unsigned nr = cluster_nctas();
  if (nr > 1) {
    __shared__ int batch_flag[2];
    if (threadIdx.x == 0) batch_flag[phase] = v;
    cluster_barrier();
    const unsigned local = (unsigned)__cvta_generic_to_shared(&batch_flag[phase]);
    int r = 0;
    for (unsigned k = 0; k < nr; ++k) {
      unsigned remote; int f;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(k));
      asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(f) : "r"(remote) : "memory");
      r |= f;
    }
    phase ^= 1u;
    v = r;
  }
  return v;
}
// before a CTA of a cluster exits: nobody may still be reading its flags
__device__ __forceinline__ void batch_or_finish() { if (cluster_nctas() > 1) cluster_barrier(); }

// ---------------------------------------------------------------- cp.async
__device__ __forceinline__ void cp_async16(void* s, const void* g) {
  unsigned a = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(a), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void* s, const void* g) {
  unsigned a = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(a), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(void* s, const void* g) {
  unsigned a = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(a), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <typename R> struct Vec { static constexpr int W = 16 / sizeof(R); };

__host__ __device__ __forceinline__ int rup(int x, int a) { return (x + a - 1) / a * a; }

// Cooperative async copy of `count` contiguous reals global -> shared (dst is 16B aligned by layout).
template <int G, typename R>
__device__ __forceinline__ void g_cp_async(const Grp<G>& g, R* dst, const R* src, int count) {
  constexpr int W = Vec<R>::W;
  if ((((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) && (count % W) == 0) {
    for (int i = g.lane * W; i < count; i += G * W) cp_async16(dst + i, src + i);
  } else {
    for (int i = g.lane; i < count; i += G) {
      if (sizeof(R) == 8) cp_async8(dst + i, src + i); else cp_async4(dst + i, src + i);
    }
  }
}

// Cooperative store shared -> global of `count` contiguous reals.
template <int G, typename R>
__device__ __forceinline__ void g_store(const Grp<G>& g, R* dst, const R* src, int count, bool valid) {
  if (!valid) return;
  for (int i = g.lane; i < count; i += G) dst[i] = src[i];
}

// ---------------------------------------------------------------- small dense linear algebra
// out(i,j) = base(i,j) + sum_k A(i,k) * B(k,j)   for i<I, j<J, k<K
// X(i,j) lives at X[i*rs + j*cs]; base may be nullptr.  Lanes split the I*J outputs.
template <int G, typename R>
__device__ __forceinline__ void g_gemm(const Grp<G>& g, int I, int J, int K,
                                       R* out, int ors,
                                       const R* base, int brs,
                                       const R* A, int ars, int acs,
                                       const R* B, int brs2, int bcs2) {
  for (int o = g.lane; o < I * J; o += G) {
    const int i = o / J, j = o - i * J;
    R a0 = base ? base[i * brs + j] : R(0), a1 = R(0);
    int k = 0;
#pragma unroll 4
    for (; k + 1 < K; k += 2) {
      a0 += A[i * ars + k * acs] * B[k * brs2 + j * bcs2];
      a1 += A[i * ars + (k + 1) * acs] * B[(k + 1) * brs2 + j * bcs2];
    }
    if (k < K) a0 += A[i * ars + k * acs] * B[k * brs2 + j * bcs2];
    out[i * ors + j] = a0 + a1;
  }
}

// Row dot product sum_k row[k] * vec[k] started at element `rot` (wrapping): when lane o reads row o
// of a row-major matrix, rot = o turns the access into a diagonal sweep that is shared-memory
// bank-conflict free for any leading dimension; two interleaved accumulators shorten the chain.
template <typename R>
__device__ __forceinline__ R dot_rot(const R* row, const R* vec, int K, int rot, R init) {
  R a0 = init, a1 = R(0);
  int k = rot % K;
  int i = 0;
  for (; i + 1 < K; i += 2) {
    int k1 = k + 1; if (k1 == K) k1 = 0;
    a0 += row[k] * vec[k];
    a1 += row[k1] * vec[k1];
    k = k1 + 1; if (k == K) k = 0;
  }
  if (i < K) a0 += row[k] * vec[k];
  return a0 + a1;
}

// In-place LU with partial pivoting of H[m x m] (row-major, ld = ldh) carrying `ncols`
// right-hand-side columns of Rhs[m x ncols] (ld = ldr) through the same row operations
// (i.e. afterwards Rhs = L^-1 P Rhs).  piv (1-based, LAPACK convention) may be nullptr.
// Column-parallel: lane j owns column j of [H | Rhs]; two group syncs per pivot step.
// MMAX bounds the per-lane copy of the pivot column (registers when m is a constant).
template <int G, int MMAX, typename R>
__device__ __forceinline__ void g_lu_factor(const Grp<G>& g, int m, R* H, int ldh,
                                            R* Rhs, int ldr, int ncols, int* piv) {
  R colk[MMAX];
  for (int k = 0; k < m; ++k) {
    // phase 1: every lane reads column k (rows k..m-1) and finds the pivot redundantly
    int p = k;
    R best = R(-1);
#pragma unroll
    for (int i = 0; i < MMAX; ++i) {
      if (i >= k && i < m) {
        colk[i] = H[i * ldh + k];
        const R a = fabs(colk[i]);
        if (a > best) { best = a; p = i; }   // first maximum wins (LAPACK idamax)
      }
    }
    R pv = colk[k];
#pragma unroll
    for (int i = 0; i < MMAX; ++i) if (i == p) pv = colk[i];
    const R rp = R(1) / pv;
    g.sync();
    // phase 2: owners apply the row interchange + elimination to their columns
    const int tot = m + ncols;
    for (int j = g.lane; j < tot; j += G) {
      R* col; int ld;
      if (j < m) { col = H + j; ld = ldh; } else { col = Rhs + (j - m); ld = ldr; }
      R akj = col[k * ld];
      if (p != k) { const R t = col[p * ld]; col[p * ld] = akj; col[k * ld] = t; akj = t; }
      if (j == k) {
        // multipliers: rows i>k of column k after the interchange
#pragma unroll
        for (int i = 0; i < MMAX; ++i) {
          if (i > k && i < m) {
            const R aik = (i == p) ? colk[k] : colk[i];
            col[i * ld] = aik * rp;
          }
        }
      } else if (j > k) {
#pragma unroll
        for (int i = 0; i < MMAX; ++i) {
          if (i > k && i < m) {
            const R aik = (i == p) ? colk[k] : colk[i];
            col[i * ld] -= (aik * rp) * akj;
          }
        }
      }
    }
    if (piv && g.lane == 0) piv[k] = p + 1;
    g.sync();
  }
}

// a / d.  An exactly-zero numerator (rows of clamped controls, active PNQP coordinates: they occur all the time) would
// take the compiler's fp64 division down its out-of-line slow path (|a| < 6.6e-37; ~60 dependent instructions, and one
// lane is enough to send the warp there); 0 / d = 0 * d for finite non-zero d, sign included.
template <typename R>
__device__ __forceinline__ R div_z(R a, R d) {
  const R ad = fabs(d);
  if (a == R(0) && ad > R(0) && ad < R(INFINITY)) return a * d;
  return a / d;
}

// Back substitution U x = b for `ncols` columns of Rhs (in place); lane-per-column.
template <int G, typename R>
__device__ __forceinline__ void g_back_subst(const Grp<G>& g, int m, const R* H, int ldh,
                                             R* Rhs, int ldr, int ncols) {
  for (int j = g.lane; j < ncols; j += G) {
    for (int i = m - 1; i >= 0; --i) {
      R acc = Rhs[i * ldr + j];
      for (int l = i + 1; l < m; ++l) acc -= H[i * ldh + l] * Rhs[l * ldr + j];
      Rhs[i * ldr + j] = div_z(acc, H[i * ldh + i]);
    }
  }
}

// Full solve with an existing factorisation (LU, piv): Rhs <- (P^T L U)^-1 Rhs, lane-per-column.
template <int G, typename R>
__device__ __forceinline__ void g_lu_solve(const Grp<G>& g, int m, const R* H, int ldh, const int* piv,
                                           R* Rhs, int ldr, int ncols) {
  for (int j = g.lane; j < ncols; j += G) {
    for (int k = 0; k < m; ++k) {
      const int p = piv[k] - 1;
      if (p != k) { const R t = Rhs[k * ldr + j]; Rhs[k * ldr + j] = Rhs[p * ldr + j]; Rhs[p * ldr + j] = t; }
    }
    for (int i = 1; i < m; ++i) {
      R acc = Rhs[i * ldr + j];
      for (int l = 0; l < i; ++l) acc -= H[i * ldh + l] * Rhs[l * ldr + j];
      Rhs[i * ldr + j] = acc;
    }
    for (int i = m - 1; i >= 0; --i) {
      R acc = Rhs[i * ldr + j];
      for (int l = i + 1; l < m; ++l) acc -= H[i * ldh + l] * Rhs[l * ldr + j];
      Rhs[i * ldr + j] = div_z(acc, H[i * ldh + i]);
    }
  }
}

}  // namespace dmpc
