// KKT adjoint of DiffLqr with TWO horizon sweeps instead of three (reference lqr/differentiable_lqr.py:78-142).
//
// The reference (and lqr_dtau_kernel + adjoint_out_kernel) does: (1) backward Riccati-vector sweep of the adjoint LQR
// problem, (2) forward rollout d-tau, (3) backward lambda / d-lambda recursions + outer products.  Sweeps (2) and (3)
// run in opposite directions, so F_t is streamed three times and C_t (top n rows) once more.
//
// Identity used here: on the optimal trajectory of an LQR problem the costate is the gradient of the cost-to-go,
//       lambda_t = V_t x_t + v_t          (forward problem)         d-lambda_t = V_t dx_t + v'_t   (adjoint problem)
// [proof: lambda_t = C_x tau + c_x + A^T lambda_{t+1} = Qxx x + Qxu u + q_x with Q = C + F^T V_{t+1} F, and u = K x + k
//  gives (Qxx + Qxu K) x + (q_x + Qxu k) = V_t x + v_t; no symmetry of C is needed].  V_t, v_t are what the forward
// Riccati sweep already holds in registers; it writes them once (lqr_factor_dmma_warp_kernel, Vsave).  With them
// lambda_{t+1}, d-lambda_{t+1} are available in a FORWARD sweep, so sweep (3) folds into sweep (2):
//
//   adj sweep 1 (t down, lqr_dtau_kernel<.., FUSED>):  k'_t = -Quu^-1 (g_u + B^T v'_{t+1}),  v'_t = q_x + Qxu k'_t   -> dc[t,:,n:], df[t-1] / dx0
//   adj sweep 2 (t up,   adjoint_fused_kernel):        du_t = K_t dx_t + k'_t;  dx_{t+1} = F_t dtau_t;
//                                                      lambda_{t+1} = V_{t+1} x_{t+1} + v_{t+1};  dlambda_{t+1} = V_{t+1} dx_{t+1} + v'_{t+1}
//                                                      dC_t, dc_t, dF_t, df_t written once
// DRAM traffic per config-5 solve: 1.33 MB + 4.4 MB instead of 2.52 MB + 4.45 MB (F read 2x instead of 3x, C not read).
#pragma once
#include "common.cuh"
#include "lqr_kernels.cuh"

namespace dmpc {

template <typename R>
struct AdjFusedParams {
  int T, B, n, m, flags;        // flags: ADJ_QUIRK_DC | ADJ_QUIRK_DF
  const R* F;                   // [T-1,B,n,s]
  const R* Ks;                  // [T,B,m,n]
  const R* Vv;                  // [T,B,n*n+n]  V_t | v_t from the forward sweep (t >= 1 valid)
  const R* x; const R* u;       // tau
  const R* vp;                  // [T-1,B,n] row t-1 = v'_t (t >= 1) from adjoint sweep 1; MAY ALIAS df: row t is consumed
                                //           (cp.async of stage t) before df_t is written in step t
  const R* dx0;                 // [B,n]    v'_0 = dlambda_0 (already the final dx0 output)
  R* dc;                        // [T,B,s]  in: k'_t in [.., n:]   out: dtau_t (RED: in only, v'_t in [.., :n])
  R* dC; R* dF; R* df;          // outputs (df nullable)
  R* red;                       // RED: [B][s*s + s + n*s + n] per-element sums over t (dC | dc | dF | df), nothing else written
};

struct AdjFusedLayout { int oF, oK, oV, ov, oxn, otau, okp, ovpn, stage, st0, st1, dtau, dxn, lam, dlam, dlamp, total; };

template <typename R>
__host__ __device__ inline AdjFusedLayout adj_fused_layout(int n, int m) {
  const int W = 16 / (int)sizeof(R);
  const int s = n + m;
  AdjFusedLayout L;
  int o = 0;
  L.oF = o; o += rup(n * s, W);
  L.oK = o; o += rup(m * n, W);
  L.oV = o; o += rup(n * n + n, W);       // V_{t+1} | v_{t+1} (contiguous in HBM)
  L.ov = L.oV + n * n;
  L.oxn = o; o += rup(n, W);              // x_{t+1}
  L.otau = o; o += rup(n, W) + rup(m, W); // x_t | u_t
  L.okp = o; o += rup(m, W);              // k'_t
  L.ovpn = o; o += rup(n, W);             // v'_{t+1}
  L.stage = o;
  o = 0;
  L.st0 = o; o += L.stage;
  L.st1 = o; o += L.stage;
  L.dtau = o; o += rup(s, W);
  L.dxn = o; o += rup(n, W);
  L.lam = o; o += rup(n, W);
  L.dlam = o; o += rup(n, W);
  L.dlamp = o; o += rup(n, W);
  L.total = o;
  return L;
}

// sum_k row[k] * vec[k] split over the 4 lanes of a quad (lane q takes k = q, q+4, ...), rotated by `rot` quads so that
// the 8 quads of a warp (8 consecutive rows of a row-major matrix) start in different banks.  K % 4 == 0.
template <typename R>
__device__ __forceinline__ R quad_dot(const R* row, const R* vec, int K, int q, int rot) {
  R a = R(0);
  int k = (4 * rot + q) % K;
  for (int i = 0; i < K; i += 4) {
    a += row[k] * vec[k];
    k += 4; if (k >= K) k -= K;
  }
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  return a;
}

// RED = ADJ_REDUCE_TB: the per-step outer products are summed over t in registers (every output (i,j) is owned by one thread
// for the whole horizon) and one record per element is written for reduce_partials_kernel; dC, dc, dF, df are never
// materialised and v'_t travels in the first n entries of the d-tau workspace (which is then input only).
template <typename R, int N, int M, int G, bool RED = false>
__global__ void __launch_bounds__(G) adjoint_fused_kernel(AdjFusedParams<R> p) {
  static_assert(N % 4 == 0 && M % 4 == 0 && G % 32 == 0, "quad-split dots");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int n = N, m = M, s = N + M;
  const int T = p.T, B = p.B;
  const int e = blockIdx.x;
  const Grp<G> g;
  const AdjFusedLayout L = adj_fused_layout<R>(n, m);
  R* sm = reinterpret_cast<R*>(smem_raw);
  R* dtau = sm + L.dtau; R* dxn = sm + L.dxn; R* lam = sm + L.lam; R* dlam = sm + L.dlam; R* dlamp = sm + L.dlamp;
  const size_t tb = (size_t)B;
  constexpr int nxoff = (N * (int)sizeof(R) + 15) / 16 * 16 / (int)sizeof(R);
  const int quad = threadIdx.x >> 2, q = threadIdx.x & 3;
  constexpr int NQ = G / 4;
  const bool quirk_dC = (p.flags & ADJ_QUIRK_DC) != 0, quirk_df = (p.flags & ADJ_QUIRK_DF) != 0;

  auto load_tiles = [&](int t, int st) {
    R* base = sm + (st ? L.st1 : L.st0);
    const size_t idx = (size_t)t * tb + e;
    g_cp_async(g, base + L.oK, p.Ks + idx * m * n, m * n);
    g_cp_async(g, base + L.otau, p.x + idx * n, n);
    g_cp_async(g, base + L.otau + nxoff, p.u + idx * m, m);
    g_cp_async(g, base + L.okp, p.dc + idx * s + n, m);
    if (t < T - 1) {
      const size_t idn = idx + tb;
      g_cp_async(g, base + L.oF, p.F + idx * n * s, n * s);
      g_cp_async(g, base + L.oV, p.Vv + idn * (n * n + n), n * n + n);
      g_cp_async(g, base + L.oxn, p.x + idn * n, n);
      if (RED) g_cp_async(g, base + L.ovpn, p.dc + idn * s, n);
      else g_cp_async(g, base + L.ovpn, p.vp + idx * n, n);
    }
    cp_async_commit();
  };
  constexpr int NR = G / s;
  static_assert(NR >= 1 && G >= s, "one column per thread");
  constexpr int KC = (s + NR - 1) / NR, KF = (n + NR - 1) / NR;
  const int tj = g.lane % s, ti = g.lane / s;
  R accC[RED ? KC : 1], accF[RED ? KF : 1], accc = R(0), accf = R(0);
  if (RED) {
#pragma unroll
    for (int k = 0; k < KC; ++k) accC[k] = R(0);
#pragma unroll
    for (int k = 0; k < KF; ++k) accF[k] = R(0);
  }
  load_tiles(0, 0);
  for (int o = g.lane; o < n; o += G) { dtau[o] = R(0); dlamp[o] = p.dx0[(size_t)e * n + o]; }   // dx_0 = 0, dlambda_0 = v'_0
  int st = 0;
  for (int t = 0; t < T; ++t) {
    if (t < T - 1) { load_tiles(t + 1, st ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    g.sync();
    const R* base = sm + (st ? L.st1 : L.st0);
    const R* Ft = base + L.oF; const R* Kt = base + L.oK; const R* Vn = base + L.oV; const R* vn = base + L.ov;
    const R* xn = base + L.oxn; const R* xt = base + L.otau; const R* ut = base + L.otau + nxoff;
    const R* kpt = base + L.okp; const R* vpn = base + L.ovpn;
    const bool more = t < T - 1;
    // ---- A: du_t = K_t dx_t + k'_t ;  lambda_{t+1} = V_{t+1} x_{t+1} + v_{t+1}
    for (int o = quad; o < m + (more ? n : 0); o += NQ) {
      if (o < m) { const R a = quad_dot(Kt + o * n, dtau, n, q, o); if (q == 0) dtau[n + o] = a + kpt[o]; }
      else { const int i = o - m; const R a = quad_dot(Vn + i * n, xn, n, q, i); if (q == 0) lam[i] = a + vn[i]; }
    }
    g.sync();
    // ---- B: dx_{t+1} = F_t dtau_t
    if (more) {
      for (int i = quad; i < n; i += NQ) { const R a = quad_dot(Ft + i * s, dtau, s, q, i); if (q == 0) dxn[i] = a; }
    }
    g.sync();
    // ---- C: dlambda_{t+1} = V_{t+1} dx_{t+1} + v'_{t+1}
    if (more) {
      for (int i = quad; i < n; i += NQ) { const R a = quad_dot(Vn + i * n, dxn, n, q, i); if (q == 0) dlam[i] = a + vpn[i]; }
    }
    g.sync();
    // ---- D: gradients of step t (differentiable_lqr.py:128-134)
    const size_t idx = (size_t)t * tb + e;
    auto tau = [&](int j) -> R { return j < n ? xt[j] : ut[j - n]; };
    // thread (ti, tj) owns column tj of the rows ti, ti + NR, ...: tau_j, dtau_j stay in registers and the row operands
    // are warp broadcasts (2 shared-memory reads per output instead of 4, no bank conflicts); rows are written whole
    if (ti < NR) {
      const R tau_j = tau(tj), dt_j = dtau[tj];
      R* dCg = RED ? nullptr : p.dC + idx * s * s + tj;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int i = ti + k * NR;
        if (i < s) {
          const R a = dtau[i] * tau_j, b = tau(i) * dt_j;
          const R v = quirk_dC ? (R(0.5) * a + b) : (R(0.5) * (a + b));
          if (RED) accC[k] += v; else dCg[i * s] = v;
        }
      }
      if (more) {
        R* dFg = RED ? nullptr : p.dF + idx * n * s + tj;
#pragma unroll
        for (int k = 0; k < KF; ++k) {
          const int i = ti + k * NR;
          if (i < n) {
            const R v = dlam[i] * tau_j + lam[i] * dt_j;
            if (RED) accF[k] += v; else dFg[i * s] = v;
          }
        }
      }
    }
    if (RED) {
      if (g.lane < s) accc += dtau[g.lane];
      if (more && g.lane < n) accf += quirk_df ? dlamp[g.lane] : dlam[g.lane];
    } else {
      for (int o = g.lane; o < s; o += G) p.dc[idx * s + o] = dtau[o];
      if (more && p.df) for (int o = g.lane; o < n; o += G) p.df[idx * n + o] = quirk_df ? dlamp[o] : dlam[o];
    }
    g.sync();
    if (more) for (int o = g.lane; o < n; o += G) { dtau[o] = dxn[o]; dlamp[o] = dlam[o]; }
    g.sync();
    st ^= 1;
  }
  if (RED) {
    R* out = p.red + (size_t)e * (s * s + s + n * s + n);
    if (ti < NR) {
#pragma unroll
      for (int k = 0; k < KC; ++k) { const int i = ti + k * NR; if (i < s) out[i * s + tj] = accC[k]; }
#pragma unroll
      for (int k = 0; k < KF; ++k) { const int i = ti + k * NR; if (i < n) out[s * s + s + i * s + tj] = accF[k]; }
    }
    if (g.lane < s) out[s * s + g.lane] = accc;
    if (g.lane < n) out[s * s + s + n * s + g.lane] = accf;
  }
}

}  // namespace dmpc
