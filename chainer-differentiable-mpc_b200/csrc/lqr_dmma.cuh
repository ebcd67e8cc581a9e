// Large-state Riccati sweep on the FP64 tensor path (DMMA, mma.sync.m8n8k4.f64) for sm_100a - FIRST GENERATION
// (one CTA per element, operands in shared memory).  The default path is lqr_dmma_warp.cuh; this kernel is kept for
// A/B runs (DMPC_DMMA_CTA=1) and provides the helpers both share (dmma884, fast_rcp, warp_gj_inverse, abs_hi).
//
// BASELINE config 5 (n=32, m=8, T=100): F^T V F is a real dense contraction (73 % of the flops), so it runs
// on DMMA; tcgen05 has no f64 kind.  One CTA of four warps owns one batch element for the
// whole horizon (reference lqr/lqr_recursion.py:79-152):
//     A    all warps : V F[:, n:] (one row block each)               mv = V f_t + v
//     B    warp 0    : Quu = Cuu + Fu^T (V Fu); Gauss-Jordan inverse of the 8 x 8 block in registers;
//                      q = c_t + F_t^T mv
//          warps 1-3 : the other 16 tiles of Mx = V F_t and 24 tiles of Q = C_t + F_t^T Mx (in place on the
//                      staged C_t) - the inverse is hidden behind this work
//     C+D  all warps : K = -Quu^-1 Qux (DMMA), k = -Quu^-1 qu, V = Qxx + Qxu K, v = qx + Qxu k
// then (optionally) the rollout x_{t+1} = F_t [x_t; K_t x_t + k_t] + f_t through a 3-deep cp.async ring.
// Tiles are streamed with 16-byte cp.async into a double-buffered stage with padded leading dimensions
// (ld = 4 mod 8 doubles) so that every DMMA fragment load is shared-memory bank-conflict free.
#pragma once
#include "common.cuh"
#include "lqr_kernels.cuh"

namespace dmpc {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Shared-memory fragment load that the compiler may not sink below later volatile asm (the DMMAs):
// lets the kernel issue the loads of k-step i+1 before the DMMAs of k-step i.
__device__ __forceinline__ double lds_f64(const double* p) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
  return v;
}

// 1/x to ~1 ulp: MUFU seed (20 bits) + two Newton steps; the LU only needs a good reciprocal
// (LAPACK dgetf2 also scales by the reciprocal of the pivot).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = r * __fma_rn(-x, r, 2.0);      // 20 -> 40 bits
  r = r * __fma_rn(-x, r, 2.0);      // 40 -> 80 bits
  r = __fma_rn(r, __fma_rn(-x, r, 1.0), r);   // final correction in residual form
  return r;
}

template <int N, int M>
struct DmmaCfg {
  static constexpr int S = N + M, NB = N / 8, SB = S / 8, NW = NB, NT = NW * 32;
  static constexpr int LDV = N + 4, LDF = S + 4, LDK = N + 4, LDQI = M + 4;
  static constexpr int OC = 0, Oc = OC + S * LDF, OF = Oc + S, Of = OF + N * LDF, STG = Of + N;
  static constexpr int OV = 2 * STG, Ov = OV + N * LDV, OMx = Ov + N, Omv = OMx + N * LDF, OQi = Omv + N,
                       TOTAL = OQi + M * LDQI;
  static constexpr int OKk = OMx;     // K_t tile aliases the Mx region, dead after phase B
  static_assert(M == 8 && N == 32, "DMMA path is instantiated for n = 32, m = 8");
  static_assert(M * LDK <= N * LDF, "Kk alias must fit in Mx");
  static_assert(STG % 2 == 0 && OV % 2 == 0 && OMx % 2 == 0 && OQi % 2 == 0, "16B alignment");
};

__device__ __forceinline__ unsigned abs_hi(double x) { return (unsigned)__double2hiint(x) & 0x7fffffffu; }

// Gauss-Jordan inverse of an 8 x 8 matrix by one warp: lane j < 8 holds column j of A, lane
// 8 <= j < 16 holds column j-8 of I; after 8 pivot steps lanes 8..15 hold the columns of A^-1.
// The pivot column is broadcast with shuffles; the (partial) pivot is chosen on the leading 32
// bits of |a_ik| (ties within 2^-20 resolve to the first row - as stable as LAPACK's exact max).
template <int M>
__device__ __forceinline__ void warp_gj_inverse(double (&c)[M]) {
  constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
  for (int k = 0; k < M; ++k) {
    double pc[M];
#pragma unroll
    for (int i = 0; i < M; ++i) pc[i] = __shfl_sync(FULL, c[i], k);
    int piv = k;
    unsigned best = abs_hi(pc[k]);
#pragma unroll
    for (int i = k + 1; i < M; ++i) { const unsigned a = abs_hi(pc[i]); if (a > best) { best = a; piv = i; } }
    if (piv != k) {                                  // warp-uniform
#pragma unroll
      for (int i = k + 1; i < M; ++i) {
        if (piv == i) {
          double tmp = c[i]; c[i] = c[k]; c[k] = tmp;
          tmp = pc[i]; pc[i] = pc[k]; pc[k] = tmp;
        }
      }
    }
    const double rp = fast_rcp(pc[k]);
    c[k] *= rp;
#pragma unroll
    for (int i = 0; i < M; ++i) if (i != k) c[i] = __fma_rn(-pc[i], c[k], c[i]);
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// Phase B work of warp W in {1,2,3} (warp 0 inverts Quu meanwhile):
//   P1'  Mx[:, :n] = V F[:, :n]: row block W (4 tiles) + a share of row block 0
//        (W=1: column blocks 0,1; W=2: 2; W=3: 3)
//   P2'  Q = C + F^T Mx except tile (u,u): row block W (5 tiles) + tiles 3(W-1)..3(W-1)+2 of
//        [(0,0) (0,1) (0,2) (0,3) (0,4) (4,0) (4,1) (4,2) (4,3)]
// Tile ownership is compile-time so that all fragments live in registers; fragment loads are
// volatile and issued one k-step ahead of the DMMAs that consume them.
template <int N, int M, int W>
__device__ __forceinline__ void phase_b_tiles(const double* V, double* Mx, const double* Ft, double* Q, int gr, int tg) {
  using Cfg = DmmaCfg<N, M>;
  constexpr int NB = Cfg::NB, NT = Cfg::NT, LDV = Cfg::LDV, LDF = Cfg::LDF;
  constexpr int NX = (W == 1) ? 2 : 1;              // row-block-0 tiles of this warp
  constexpr int JX0 = (W == 1) ? 0 : W;             // their first column block
  // ---------------- P1'
  {
    double acc[NB][2], accx[NX][2];
#pragma unroll
    for (int j = 0; j < NB; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
#pragma unroll
    for (int j = 0; j < NX; ++j) { accx[j][0] = 0.0; accx[j][1] = 0.0; }
    const double* Va = V + (W * 8 + gr) * LDV + tg;
    const double* V0 = V + gr * LDV + tg;
    const double* Fb = Ft + tg * LDF + gr;
    double a = lds_f64(Va), a0 = lds_f64(V0), b[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) b[j] = lds_f64(Fb + j * 8);
#pragma unroll
    for (int k0 = 0; k0 < N; k0 += 4) {
      double an = 0.0, a0n = 0.0, bn[NB];
      if (k0 + 4 < N) {
        an = lds_f64(Va + k0 + 4); a0n = lds_f64(V0 + k0 + 4);
#pragma unroll
        for (int j = 0; j < NB; ++j) bn[j] = lds_f64(Fb + (k0 + 4) * LDF + j * 8);
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) dmma884(acc[j][0], acc[j][1], a, b[j]);
#pragma unroll
      for (int j = 0; j < NX; ++j) dmma884(accx[j][0], accx[j][1], a0, b[JX0 + j]);
      if (k0 + 4 < N) {
        a = an; a0 = a0n;
#pragma unroll
        for (int j = 0; j < NB; ++j) b[j] = bn[j];
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j)
      *reinterpret_cast<double2*>(Mx + (W * 8 + gr) * LDF + j * 8 + tg * 2) = make_double2(acc[j][0], acc[j][1]);
#pragma unroll
    for (int j = 0; j < NX; ++j)
      *reinterpret_cast<double2*>(Mx + gr * LDF + (JX0 + j) * 8 + tg * 2) = make_double2(accx[j][0], accx[j][1]);
  }
  named_bar_sync(1, NT - 32);                  // Mx complete among warps 1..3 (VFu came through barrier (1))
  // ---------------- P2'
  {
    constexpr int XI[3] = {(3 * (W - 1) + 0 < 5) ? 0 : NB, (3 * (W - 1) + 1 < 5) ? 0 : NB, (3 * (W - 1) + 2 < 5) ? 0 : NB};
    constexpr int XJ[3] = {(3 * (W - 1) + 0 < 5) ? 3 * (W - 1) + 0 : 3 * (W - 1) + 0 - 5,
                           (3 * (W - 1) + 1 < 5) ? 3 * (W - 1) + 1 : 3 * (W - 1) + 1 - 5,
                           (3 * (W - 1) + 2 < 5) ? 3 * (W - 1) + 2 : 3 * (W - 1) + 2 - 5};
    double acc[NB + 1][2], ex[3][2];
#pragma unroll
    for (int j = 0; j <= NB; ++j) {
      const double2 c2 = *reinterpret_cast<const double2*>(Q + (W * 8 + gr) * LDF + j * 8 + tg * 2);
      acc[j][0] = c2.x; acc[j][1] = c2.y;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double2 c2 = *reinterpret_cast<const double2*>(Q + (XI[j] * 8 + gr) * LDF + XJ[j] * 8 + tg * 2);
      ex[j][0] = c2.x; ex[j][1] = c2.y;
    }
    const double* Fa = Ft + tg * LDF + gr;          // F^T fragments: row block r -> Fa[k0*LDF + 8 r]
    const double* Mb = Mx + tg * LDF + gr;
    double a = lds_f64(Fa + W * 8), a0 = lds_f64(Fa), au = lds_f64(Fa + NB * 8), b[NB + 1];
#pragma unroll
    for (int j = 0; j <= NB; ++j) b[j] = lds_f64(Mb + j * 8);
#pragma unroll
    for (int k0 = 0; k0 < N; k0 += 4) {
      double an = 0.0, a0n = 0.0, aun = 0.0, bn[NB + 1];
      if (k0 + 4 < N) {
        an = lds_f64(Fa + (k0 + 4) * LDF + W * 8); a0n = lds_f64(Fa + (k0 + 4) * LDF); aun = lds_f64(Fa + (k0 + 4) * LDF + NB * 8);
#pragma unroll
        for (int j = 0; j <= NB; ++j) bn[j] = lds_f64(Mb + (k0 + 4) * LDF + j * 8);
      }
#pragma unroll
      for (int j = 0; j <= NB; ++j) dmma884(acc[j][0], acc[j][1], a, b[j]);
#pragma unroll
      for (int j = 0; j < 3; ++j) dmma884(ex[j][0], ex[j][1], XI[j] == 0 ? a0 : au, b[XJ[j]]);
      if (k0 + 4 < N) {
        a = an; a0 = a0n; au = aun;
#pragma unroll
        for (int j = 0; j <= NB; ++j) b[j] = bn[j];
      }
    }
    // every Q tile is read and written by one warp only -> in-place stores need no barrier
#pragma unroll
    for (int j = 0; j <= NB; ++j)
      *reinterpret_cast<double2*>(Q + (W * 8 + gr) * LDF + j * 8 + tg * 2) = make_double2(acc[j][0], acc[j][1]);
#pragma unroll
    for (int j = 0; j < 3; ++j)
      *reinterpret_cast<double2*>(Q + (XI[j] * 8 + gr) * LDF + XJ[j] * 8 + tg * 2) = make_double2(ex[j][0], ex[j][1]);
  }
}


// warp-level 8x8 output tile: acc += A(8 x K) * B(K x 8), fragments fetched by the given functors
#define DMPC_TILE_LOOP(acc0, acc1, K, AEXPR, BEXPR)            \
  _Pragma("unroll") for (int k0 = 0; k0 < (K); k0 += 4) {      \
    const double a_ = (AEXPR);                                 \
    const double b_ = (BEXPR);                                 \
    dmma884(acc0, acc1, a_, b_);                               \
  }

template <int N, int M>
__global__ void __launch_bounds__(DmmaCfg<N, M>::NT, 3) lqr_factor_dmma_kernel(LqrParams<double> p) {
  using Cfg = DmmaCfg<N, M>;
  constexpr int S = Cfg::S, NB = Cfg::NB, NT = Cfg::NT;
  constexpr int LDV = Cfg::LDV, LDF = Cfg::LDF, LDK = Cfg::LDK, LDQI = Cfg::LDQI;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  double* V = sm + Cfg::OV; double* v = sm + Cfg::Ov; double* Mx = sm + Cfg::OMx; double* mv = sm + Cfg::Omv;
  double* Kk = sm + Cfg::OKk; double* Qi = sm + Cfg::OQi;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gr = lane >> 2, tg = lane & 3;
  const int T = p.T, B = p.B;
  const size_t tb = (size_t)B;
  const bool have_f = p.f != nullptr;
  const bool save_fac = (p.flags & LQR_SAVE_FAC) && p.fac;
  // persistent CTAs: the grid is (CTAs per SM) x (SMs) and every CTA strides over the batch, so the
  // launch can leave shared memory free for a memory-bound kernel running beside it (bench.py --chunks)
  for (int e = blockIdx.x; e < B; e += gridDim.x) {
  // per-thread 16-byte chunk of a row: 20 chunks per row of S doubles, 6 rows in flight per pass
  const int ld_cc = tid % (S / 2), ld_r0 = tid / (S / 2);
  constexpr int LD_RSTEP = NT / (S / 2);          // 6 full rows per pass (threads >= 120 idle)
  const bool ld_active = ld_r0 < LD_RSTEP;

  auto load_tiles = [&](int t, int st) {
    double* base = sm + st * Cfg::STG;
    const size_t idx = (size_t)t * tb + e;
    if (ld_active) {
      const double* Cg = p.C + idx * S * S + ld_cc * 2;
      double* Cd = base + Cfg::OC + ld_cc * 2;
      for (int r = ld_r0; r < S; r += LD_RSTEP) cp_async16(Cd + r * LDF, Cg + r * S);
      if (t < T - 1) {
        const double* Fg = p.F + idx * N * S + ld_cc * 2;
        double* Fd = base + Cfg::OF + ld_cc * 2;
        for (int r = ld_r0; r < N; r += LD_RSTEP) cp_async16(Fd + r * LDF, Fg + r * S);
      }
    }
    if (tid < S / 2) cp_async16(base + Cfg::Oc + tid * 2, p.c + idx * S + tid * 2);
    else if (t < T - 1 && have_f && tid >= 32 && tid < 32 + N / 2)
      cp_async16(base + Cfg::Of + (tid - 32) * 2, p.f + idx * N + (tid - 32) * 2);
    cp_async_commit();
  };

  load_tiles(T - 1, 0);
  cp_async_wait<0>();
  __syncthreads();
  int st = 0;
  for (int t = T - 1; t >= 0; --t) {
    // tiles of step t are resident (waited for before the barrier that closed step t+1);
    // prefetch step t-1 into the other stage
    if (t > 0) load_tiles(t - 1, st ^ 1);
    double* base = sm + st * Cfg::STG;
    double* Q = base + Cfg::OC;        // C_t, overwritten in place by Q_t
    double* q = base + Cfg::Oc;        // c_t -> q_t
    const double* Ft = base + Cfg::OF;
    const double* ft = base + Cfg::Of;
    const size_t idx = (size_t)t * tb + e;
    if (t < T - 1) {
      // ---- phase A (all warps): VFu = V F[:, n:]  (row block = warp) ; mv = V f + v
      {
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < N; k0 += 16) {        // four independent accumulation chains
          dmma884(a0, a1, V[(warp * 8 + gr) * LDV + k0 + tg], Ft[(k0 + tg) * LDF + N + gr]);
          dmma884(b0, b1, V[(warp * 8 + gr) * LDV + k0 + 4 + tg], Ft[(k0 + 4 + tg) * LDF + N + gr]);
          dmma884(c0, c1, V[(warp * 8 + gr) * LDV + k0 + 8 + tg], Ft[(k0 + 8 + tg) * LDF + N + gr]);
          dmma884(d0, d1, V[(warp * 8 + gr) * LDV + k0 + 12 + tg], Ft[(k0 + 12 + tg) * LDF + N + gr]);
        }
        a0 = (a0 + b0) + (c0 + d0); a1 = (a1 + b1) + (c1 + d1);
        *reinterpret_cast<double2*>(Mx + (warp * 8 + gr) * LDF + N + tg * 2) = make_double2(a0, a1);
        const int i = tid >> 2, part = tid & 3;            // NT == 4 N
        double sacc = 0.0;
        if (have_f) {
#pragma unroll
          for (int k = 0; k < N / 4; ++k) sacc += V[i * LDV + 4 * k + part] * ft[4 * k + part];   // interleaved split: bank-conflict free
        }
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
        if (part == 0) mv[i] = sacc + v[i];
      }
      __syncthreads();                                                   // (1)
    }
    // ---- phase B: warp 0 -> Quu tile + Gauss-Jordan inverse ; warps 1..3 -> rest of Mx and Q
    if (warp == 0) {
      if (t < T - 1) {
        const double2 c2 = *reinterpret_cast<const double2*>(Q + (N + gr) * LDF + N + tg * 2);
        double a0 = c2.x, a1 = c2.y, b0 = 0.0, b1 = 0.0, c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < N; k0 += 16) {
          dmma884(a0, a1, Ft[(k0 + tg) * LDF + N + gr], Mx[(k0 + tg) * LDF + N + gr]);
          dmma884(b0, b1, Ft[(k0 + 4 + tg) * LDF + N + gr], Mx[(k0 + 4 + tg) * LDF + N + gr]);
          dmma884(c0, c1, Ft[(k0 + 8 + tg) * LDF + N + gr], Mx[(k0 + 8 + tg) * LDF + N + gr]);
          dmma884(d0, d1, Ft[(k0 + 12 + tg) * LDF + N + gr], Mx[(k0 + 12 + tg) * LDF + N + gr]);
        }
        a0 = (a0 + b0) + (c0 + d0); a1 = (a1 + b1) + (c1 + d1);
        *reinterpret_cast<double2*>(Q + (N + gr) * LDF + N + tg * 2) = make_double2(a0, a1);
        __syncwarp();
      }
      double c[M];
#pragma unroll
      for (int i = 0; i < M; ++i)
        c[i] = (lane < M) ? Q[(N + i) * LDF + N + lane] : ((lane - M == i) ? 1.0 : 0.0);
      warp_gj_inverse<M>(c);
      if (lane >= M && lane < 2 * M) {
        double* fg = save_fac ? p.fac + idx * (M * M + N * M) : nullptr;
#pragma unroll
        for (int i = 0; i < M; ++i) { Qi[i * LDQI + lane - M] = c[i]; if (fg) fg[i * M + lane - M] = c[i]; }
      }
      if (t < T - 1) {
        // q = c + F^T mv (in place on the staged c): warp 0 has slack after the inverse
#pragma unroll
        for (int o = lane; o < S * 4; o += 32) {
          const int i = o >> 2, part = o & 3;
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int k = 0; k < N / 4; k += 2) {
            s0 += Ft[(4 * k + part) * LDF + i] * mv[4 * k + part];
            s1 += Ft[(4 * k + 4 + part) * LDF + i] * mv[4 * k + 4 + part];
          }
          double sacc = s0 + s1;
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
          if (part == 0) q[i] += sacc;
        }
      }
    } else if (t < T - 1) {
      switch (warp) {
        case 1: phase_b_tiles<N, M, 1>(V, Mx, Ft, Q, gr, tg); break;
        case 2: phase_b_tiles<N, M, 2>(V, Mx, Ft, Q, gr, tg); break;
        default: phase_b_tiles<N, M, 3>(V, Mx, Ft, Q, gr, tg); break;
      }
    }
    __syncthreads();                                                     // (2)
    // ---- phase C+D (all warps, column block = warp): K = -Quu^-1 Qux, k = -Quu^-1 qu, then
    //      V = Qxx + Qxu K, v = qx + Qxu k.  The reference's extra terms K^T (Qux + Quu K) and
    //      K^T (qu + Quu k) (lqr_recursion.py:151-152) vanish identically for the exact gain and
    //      are O(eps cond(Quu)) here - below its own rounding noise.
    {
      double a0 = 0.0, a1 = 0.0;
      DMPC_TILE_LOOP(a0, a1, M, Qi[gr * LDQI + k0 + tg], Q[(N + k0 + tg) * LDF + warp * 8 + gr])
      a0 = -a0; a1 = -a1;
      *reinterpret_cast<double2*>(Kk + gr * LDK + warp * 8 + tg * 2) = make_double2(a0, a1);
      *reinterpret_cast<double2*>(p.Ks + idx * M * N + gr * N + warp * 8 + tg * 2) = make_double2(a0, a1);
      double kk = 0.0;                       // k_l on lanes < M of every warp (redundant, avoids a barrier)
      if (lane < M) {
#pragma unroll
        for (int l = 0; l < M; ++l) kk += Qi[lane * LDQI + l] * q[N + l];
        kk = -kk;
        if (warp == 0) p.ks[idx * M + lane] = kk;
      }
      if (save_fac) {
        double* fg = p.fac + idx * (M * M + N * M) + M * M;
        for (int o = tid; o < N * M; o += NT) { const int i = o / M, j = o - i * M; fg[o] = Q[i * LDF + N + j]; }
      }
      __syncwarp();
      if (t > 0) {
        const int jb = warp;
        double acc[NB][2];
#pragma unroll
        for (int ib = 0; ib < NB; ++ib) {
          const double2 c2 = *reinterpret_cast<const double2*>(Q + (ib * 8 + gr) * LDF + jb * 8 + tg * 2);
          acc[ib][0] = c2.x; acc[ib][1] = c2.y;
        }
#pragma unroll
        for (int k0 = 0; k0 < M; k0 += 4) {
          const double b1 = Kk[(k0 + tg) * LDK + jb * 8 + gr];
#pragma unroll
          for (int ib = 0; ib < NB; ++ib) dmma884(acc[ib][0], acc[ib][1], Q[(ib * 8 + gr) * LDF + N + k0 + tg], b1);
        }
#pragma unroll
        for (int ib = 0; ib < NB; ++ib)
          *reinterpret_cast<double2*>(V + (ib * 8 + gr) * LDV + jb * 8 + tg * 2) = make_double2(acc[ib][0], acc[ib][1]);
        double kb[M];
#pragma unroll
        for (int l = 0; l < M; ++l) kb[l] = __shfl_sync(0xffffffffu, kk, l);
        if (lane < 8) {
          const int i = warp * 8 + lane;
          double a = q[i];
#pragma unroll
          for (int l = 0; l < M; ++l) a += Q[i * LDF + N + l] * kb[l];
          v[i] = a;
        }
      }
    }
    cp_async_wait<0>();                                                  // tiles of step t-1 have landed
    __syncthreads();                                                     // (4) closes step t and publishes them
    st ^= 1;
  }

  // ---- fused rollout (lqr_recursion.py:160-200): the CTA re-streams K_t, k_t, F_t, f_t through a
  //      3-deep cp.async ring carved from the (now dead) Riccati stage buffers.  While this CTA is in
  //      its memory-bound rollout the other CTAs of the SM keep the DMMA pipe busy.
  if (p.flags & LQR_DO_ROLLOUT) {
    constexpr int RK = 0, Rk = RK + M * N, RF = Rk + M, Rf = RF + N * LDF, RSTG = Rf + N;   // 1704 doubles
    static_assert(3 * RSTG <= 2 * Cfg::STG, "rollout ring must fit in the Riccati stages");
    double* xcur = sm + Cfg::OV;           // [S]  (V is dead)
    double* xnew = sm + Cfg::OV + 64;      // [N]
    auto load_roll = [&](int t, int slot) {
      double* base = sm + slot * RSTG;
      const size_t idx = (size_t)t * tb + e;
      for (int c = tid; c < M * N / 2; c += NT) cp_async16(base + RK + c * 2, p.Ks + idx * M * N + c * 2);
      if (tid < M / 2) cp_async16(base + Rk + tid * 2, p.ks + idx * M + tid * 2);
      if (t < T - 1) {
        if (ld_active) {
          const double* Fg = p.F + idx * N * S + ld_cc * 2;
          double* Fd = base + RF + ld_cc * 2;
          for (int r = ld_r0; r < N; r += LD_RSTEP) cp_async16(Fd + r * LDF, Fg + r * S);
        }
        if (have_f && tid >= 32 && tid < 32 + N / 2) cp_async16(base + Rf + (tid - 32) * 2, p.f + idx * N + (tid - 32) * 2);
      }
    };
    // prologue: stages 0,1 (one commit group per timestep, empty groups keep the accounting uniform)
    load_roll(0, 0); cp_async_commit();
    if (T > 1) load_roll(1, 1);
    cp_async_commit();
    if (tid < N) xcur[tid] = p.x0[(size_t)e * N + tid];
    for (int t = 0; t < T; ++t) {
      if (t + 2 < T) load_roll(t + 2, (t + 2) % 3);
      cp_async_commit();
      cp_async_wait<2>();
      __syncthreads();
      const double* base = sm + (t % 3) * RSTG;
      const double* Kt = base + RK; const double* kt = base + Rk; const double* Ft = base + RF; const double* ft = base + Rf;
      {   // u = K x + k : 16 lanes per control
        const int o = tid >> 4, part = tid & 15;
        double a = Kt[o * N + 2 * part] * xcur[2 * part] + Kt[o * N + 2 * part + 1] * xcur[2 * part + 1];
        a += __shfl_xor_sync(0xffffffffu, a, 8);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        if (part == 0) xcur[N + o] = a + kt[o];
      }
      __syncthreads();
      {
        const size_t idx = (size_t)t * tb + e;
        if (tid < N) { if (p.x) p.x[idx * N + tid] = xcur[tid]; }
        else if (tid < S) { if (p.u) p.u[idx * M + tid - N] = xcur[tid]; }
        if (p.tau_out && tid < S) p.tau_out[idx * S + tid] = xcur[tid];
      }
      if (t < T - 1) {   // x' = F [x;u] + f : 4 lanes per state, interleaved split (conflict free with ld = S + 4)
        const int i = tid >> 2, part = tid & 3;
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < S / 4; ++j) a += Ft[i * LDF + 4 * j + part] * xcur[4 * j + part];
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        if (part == 0) xnew[i] = a + (have_f ? ft[i] : 0.0);
        __syncthreads();
        if (tid < N) xcur[tid] = xnew[tid];
      }
      __syncthreads();
    }
  }
  __syncthreads();
  }   // element loop
}

}  // namespace dmpc
