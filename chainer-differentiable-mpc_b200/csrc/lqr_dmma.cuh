// Large-state Riccati sweep on the FP64 tensor path (DMMA, mma.sync.m8n8k4.f64) for sm_100a.
//
// BASELINE config 5 (n=32, m=8, T=100): F^T V F is a real dense contraction (73 % of the flops),
// so it runs on DMMA; tcgen05 has no f64 kind.  One CTA of (n+m)/8 warps owns one
// batch element for the whole horizon; warp w owns column block w of every tile product:
//     P1  Mx = V F_t                    (n x s,  K = n)     mv = V f_t + v
//     P2  Q  = C_t + F_t^T Mx  (in place on the staged C_t)  q  = c_t + F_t^T mv
//     LU  warp 0: in-register pivoted LU of Quu carrying [-Qux | -qu | I] -> K_t, k_t, Quu^-1
//     P3  P = [Qux|qu] + Quu [K|k];  V = Qxx + Qxu K + K^T P;  v likewise
// (reference lqr/lqr_recursion.py:79-152).  Tiles are streamed with 16-byte cp.async into a
// double-buffered stage with padded leading dimensions (ld = 4 mod 8 doubles) so that every DMMA
// fragment load is shared-memory bank-conflict free.
#pragma once
#include "common.cuh"
#include "lqr_kernels.cuh"

namespace dmpc {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 1/x to ~1 ulp: MUFU seed (20 bits) + two Newton steps; the LU only needs a good reciprocal
// (LAPACK dgetf2 also scales by the reciprocal of the pivot).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = r * __fma_rn(-x, r, 2.0);
  r = r * __fma_rn(-x, r, 2.0);
  r = r * __fma_rn(-x, r, 2.0);
  return r;
}

template <int N, int M>
struct DmmaCfg {
  static constexpr int S = N + M, NB = N / 8, SB = S / 8, NT = SB * 32;
  static constexpr int LDV = N + 4, LDF = S + 4, LDK = N + 4;
  static constexpr int OC = 0, Oc = OC + S * LDF, OF = Oc + S, Of = OF + N * LDF, STG = Of + N;
  static constexpr int OV = 2 * STG, Ov = OV + N * LDV, OMx = Ov + N, Omv = OMx + N * LDF, TOTAL = Omv + N;
  static constexpr int OKk = OMx, OP = OMx + M * LDK;     // alias the dead Mx region after P2
  static_assert(M == 8 && N % 8 == 0 && N >= 16, "DMMA path: m == 8, n multiple of 8");
  static_assert(2 * M * LDK <= N * LDF, "Kk/P alias must fit in Mx");
  static_assert(STG % 2 == 0 && OV % 2 == 0 && OMx % 2 == 0, "16B alignment");
};

template <int N, int M>
__global__ void __launch_bounds__(DmmaCfg<N, M>::NT, 3) lqr_factor_dmma_kernel(LqrParams<double> p) {
  using Cfg = DmmaCfg<N, M>;
  constexpr int S = Cfg::S, NB = Cfg::NB, SB = Cfg::SB, NT = Cfg::NT;
  constexpr int LDV = Cfg::LDV, LDF = Cfg::LDF, LDK = Cfg::LDK;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  double* V = sm + Cfg::OV; double* v = sm + Cfg::Ov; double* Mx = sm + Cfg::OMx; double* mv = sm + Cfg::Omv;
  double* Kk = sm + Cfg::OKk; double* Pm = sm + Cfg::OP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gr = lane >> 2, tg = lane & 3;
  const int T = p.T, B = p.B;
  const int e = blockIdx.x;
  const size_t tb = (size_t)B;
  const bool have_f = p.f != nullptr;
  const bool save_fac = (p.flags & LQR_SAVE_FAC) && p.fac;

  auto load_tiles = [&](int t, int st) {
    double* base = sm + st * Cfg::STG;
    const size_t idx = (size_t)t * tb + e;
    const double* Cg = p.C + idx * S * S;
    for (int c = tid; c < S * (S / 2); c += NT) {
      const int row = c / (S / 2), cc = c - row * (S / 2);
      cp_async16(base + Cfg::OC + row * LDF + cc * 2, Cg + row * S + cc * 2);
    }
    const double* cg = p.c + idx * S;
    for (int c = tid; c < S / 2; c += NT) cp_async16(base + Cfg::Oc + c * 2, cg + c * 2);
    if (t < T - 1) {
      const double* Fg = p.F + idx * N * S;
      for (int c = tid; c < N * (S / 2); c += NT) {
        const int row = c / (S / 2), cc = c - row * (S / 2);
        cp_async16(base + Cfg::OF + row * LDF + cc * 2, Fg + row * S + cc * 2);
      }
      if (have_f) {
        const double* fg = p.f + idx * N;
        for (int c = tid; c < N / 2; c += NT) cp_async16(base + Cfg::Of + c * 2, fg + c * 2);
      }
    }
    cp_async_commit();
  };

  load_tiles(T - 1, 0);
  int st = 0;
  for (int t = T - 1; t >= 0; --t) {
    if (t > 0) { load_tiles(t - 1, st ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    double* base = sm + st * Cfg::STG;
    double* Q = base + Cfg::OC;        // C_t, overwritten in place by Q_t
    double* q = base + Cfg::Oc;        // c_t -> q_t
    const double* Ft = base + Cfg::OF;
    const double* ft = base + Cfg::Of;
    if (t < T - 1) {
      // ---- P1: Mx = V F (warp = column block), mv = V f + v
      {
        double acc[NB][2];
#pragma unroll
        for (int ib = 0; ib < NB; ++ib) { acc[ib][0] = 0.0; acc[ib][1] = 0.0; }
#pragma unroll 2
        for (int k0 = 0; k0 < N; k0 += 4) {
          const double b = Ft[(k0 + tg) * LDF + warp * 8 + gr];
#pragma unroll
          for (int ib = 0; ib < NB; ++ib) {
            const double a = V[(ib * 8 + gr) * LDV + k0 + tg];
            dmma884(acc[ib][0], acc[ib][1], a, b);
          }
        }
#pragma unroll
        for (int ib = 0; ib < NB; ++ib)
          *reinterpret_cast<double2*>(Mx + (ib * 8 + gr) * LDF + warp * 8 + tg * 2) = make_double2(acc[ib][0], acc[ib][1]);
        if (tid < N * 4) {
          const int i = tid >> 2, part = tid & 3;
          double sacc = 0.0;
          if (have_f) {
#pragma unroll
            for (int k = 0; k < N / 4; ++k) sacc += V[i * LDV + part * (N / 4) + k] * ft[part * (N / 4) + k];
          }
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
          if (part == 0) mv[i] = sacc + v[i];
        }
      }
      __syncthreads();
      // ---- P2: Q = C + F^T Mx (in place), q = c + F^T mv (in place)
      {
        double acc[SB][2];
#pragma unroll
        for (int ib = 0; ib < SB; ++ib) {
          const double2 c2 = *reinterpret_cast<const double2*>(Q + (ib * 8 + gr) * LDF + warp * 8 + tg * 2);
          acc[ib][0] = c2.x; acc[ib][1] = c2.y;
        }
#pragma unroll 2
        for (int k0 = 0; k0 < N; k0 += 4) {
          const double b = Mx[(k0 + tg) * LDF + warp * 8 + gr];
#pragma unroll
          for (int ib = 0; ib < SB; ++ib) {
            const double a = Ft[(k0 + tg) * LDF + ib * 8 + gr];
            dmma884(acc[ib][0], acc[ib][1], a, b);
          }
        }
        // q first (reads c), then the in-place stores of this warp's column block
        if (tid < S * 4) {
          const int i = tid >> 2, part = tid & 3;
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < N / 4; ++k) sacc += Ft[(part * (N / 4) + k) * LDF + i] * mv[part * (N / 4) + k];
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
          if (part == 0) q[i] += sacc;
        }
#pragma unroll
        for (int ib = 0; ib < SB; ++ib)
          *reinterpret_cast<double2*>(Q + (ib * 8 + gr) * LDF + warp * 8 + tg * 2) = make_double2(acc[ib][0], acc[ib][1]);
      }
      __syncthreads();
    }
    const size_t idx = (size_t)t * tb + e;
    // ---- LU (warp 0): columns of [Quu | -Qux | -qu | I] live in registers, lane j <-> columns j and j+32
    if (warp == 0) {
      double ca[M], cb[M];
      const int j2 = lane + 32;                 // second column index (valid if < NC)
      constexpr int NC = M + N + 1 + M;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        if (lane < M) ca[i] = Q[(N + i) * LDF + N + lane];
        else ca[i] = -Q[(N + i) * LDF + (lane - M)];
        const int r = j2 - M;                   // rhs index of the second column
        double vb = 0.0;
        if (r < N) vb = -Q[(N + i) * LDF + r];
        else if (r == N) vb = -q[N + i];
        else if (r - N - 1 == i) vb = 1.0;
        cb[i] = vb;
      }
#pragma unroll
      for (int k = 0; k < M; ++k) {
        double pc[M];
#pragma unroll
        for (int i = 0; i < M; ++i) pc[i] = __shfl_sync(0xffffffffu, ca[i], k);
        int piv = k;
        double best = fabs(pc[k]);
#pragma unroll
        for (int i = k + 1; i < M; ++i) { const double a = fabs(pc[i]); if (a > best) { best = a; piv = i; } }
        double pv = pc[k];
#pragma unroll
        for (int i = k + 1; i < M; ++i) if (i == piv) pv = pc[i];
        const double rp = fast_rcp(pv);
        // row interchange k <-> piv on this lane's columns and on the broadcast pivot column
#pragma unroll
        for (int i = k + 1; i < M; ++i) {
          if (i == piv) {
            double tmp = ca[i]; ca[i] = ca[k]; ca[k] = tmp;
            tmp = cb[i]; cb[i] = cb[k]; cb[k] = tmp;
            pc[i] = pc[k];
          }
        }
#pragma unroll
        for (int i = k + 1; i < M; ++i) {
          const double l = pc[i] * rp;
          if (lane != k) ca[i] = __fma_rn(-l, ca[k], ca[i]); else ca[i] = l;
          cb[i] = __fma_rn(-l, cb[k], cb[i]);
        }
      }
      // back substitution: U(i,l) is row i of column l (lane l, l < M)
#pragma unroll
      for (int i = M - 1; i >= 0; --i) {
        const double uii = __shfl_sync(0xffffffffu, ca[i], i);
        const double ri = fast_rcp(uii);
        double xa = ca[i], xb = cb[i];
#pragma unroll
        for (int l = i + 1; l < M; ++l) {
          const double uil = __shfl_sync(0xffffffffu, ca[i], l);
          xa = __fma_rn(-uil, ca[l], xa);
          xb = __fma_rn(-uil, cb[l], xb);
        }
        if (lane >= M) ca[i] = xa * ri;
        cb[i] = xb * ri;
      }
      // scatter: Kk (smem), K/k (global), Quu^-1 (global fac)
      double* Kg = p.Ks + idx * M * N; double* kg = p.ks + idx * M;
      double* fg = save_fac ? p.fac + idx * (M * M + N * M) : nullptr;
#pragma unroll
      for (int i = 0; i < M; ++i) {
        if (lane >= M) { Kk[i * LDK + lane - M] = ca[i]; Kg[i * N + lane - M] = ca[i]; }
        const int r = j2 - M;
        if (r < N) { Kk[i * LDK + r] = cb[i]; Kg[i * N + r] = cb[i]; }
        else if (r == N) { Kk[i * LDK + N] = cb[i]; kg[i] = cb[i]; }
        else if (r < N + 1 + M && fg) fg[i * M + (r - N - 1)] = cb[i];
      }
      (void)NC;
    } else if (save_fac) {
      // the other warps park Qxu for the adjoint while warp 0 factorises
      double* fg = p.fac + idx * (M * M + N * M) + M * M;
      for (int o = tid - 32; o < N * M; o += NT - 32) { const int i = o / M, j = o - i * M; fg[o] = Q[i * LDF + N + j]; }
    }
    __syncthreads();
    if (t > 0) {
      // ---- P3: P = [Qux|qu] + Quu [K|k] ; V = Qxx + Qxu K + K^T P ; v = qx + Qxu k + K^T p
      if (warp < NB) {
        const int jb = warp;
        {
          const double2 c2 = *reinterpret_cast<const double2*>(Q + (N + gr) * LDF + jb * 8 + tg * 2);
          double p0 = c2.x, p1 = c2.y;
#pragma unroll
          for (int k0 = 0; k0 < M; k0 += 4) {
            const double a = Q[(N + gr) * LDF + N + k0 + tg];
            const double b = Kk[(k0 + tg) * LDK + jb * 8 + gr];
            dmma884(p0, p1, a, b);
          }
          *reinterpret_cast<double2*>(Pm + gr * LDK + jb * 8 + tg * 2) = make_double2(p0, p1);
        }
        __syncwarp();
        double acc[NB][2];
#pragma unroll
        for (int ib = 0; ib < NB; ++ib) {
          const double2 c2 = *reinterpret_cast<const double2*>(Q + (ib * 8 + gr) * LDF + jb * 8 + tg * 2);
          acc[ib][0] = c2.x; acc[ib][1] = c2.y;
        }
#pragma unroll
        for (int k0 = 0; k0 < M; k0 += 4) {
          const double b1 = Kk[(k0 + tg) * LDK + jb * 8 + gr];
          const double b2 = Pm[(k0 + tg) * LDK + jb * 8 + gr];
#pragma unroll
          for (int ib = 0; ib < NB; ++ib) {
            const double a1 = Q[(ib * 8 + gr) * LDF + N + k0 + tg];        // Qxu
            dmma884(acc[ib][0], acc[ib][1], a1, b1);
            const double a2 = Kk[(k0 + tg) * LDK + ib * 8 + gr];           // K^T
            dmma884(acc[ib][0], acc[ib][1], a2, b2);
          }
        }
#pragma unroll
        for (int ib = 0; ib < NB; ++ib)
          *reinterpret_cast<double2*>(V + (ib * 8 + gr) * LDV + jb * 8 + tg * 2) = make_double2(acc[ib][0], acc[ib][1]);
      } else if (warp == NB) {
        // vector part on the last warp: p = qu + Quu k ; v = qx + Qxu k + K^T p
        double pj = 0.0;
        if (lane < M) {
          pj = q[N + lane];
#pragma unroll
          for (int l = 0; l < M; ++l) pj += Q[(N + lane) * LDF + N + l] * Kk[l * LDK + N];
        }
        double pb[M];
#pragma unroll
        for (int l = 0; l < M; ++l) pb[l] = __shfl_sync(0xffffffffu, pj, l);
        for (int i = lane; i < N; i += 32) {
          double a = q[i], b = 0.0;
#pragma unroll
          for (int l = 0; l < M; ++l) {
            a += Q[i * LDF + N + l] * Kk[l * LDK + N];
            b += Kk[l * LDK + i] * pb[l];
          }
          v[i] = a + b;
        }
      }
    }
    __syncthreads();
    st ^= 1;
  }
}

}  // namespace dmpc
