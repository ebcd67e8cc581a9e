// Large-state Riccati sweep, second generation: ONE WARP owns one batch element and keeps the whole
// recursion state in DMMA fragment registers (sm_100a, mma.sync.m8n8k4.f64).
//
// Why: a CTA-per-element kernel that fetches every DMMA operand from shared memory (round 1's first generation,
// profiles/r1/r1f_factor_cta_full.summary.txt) is bound by the shared-memory pipe (l1tex data-pipe wavefronts 70 % of
// peak, 2995 wavefronts per element-step against 1600 DMMA-pipe cycles).  Here the operands stay in registers:
//
//   * the accumulator layout of a DMMA output tile (lane (g,t) holds D[g][2t], D[g][2t+1]) IS a valid
//     B (or A) fragment of the next product if the contraction index inside each 8-block is enumerated
//     in the order k(t,e) = 2t+e.  A contraction is invariant under a permutation of k applied to both
//     operands, so V_t (accumulators of step t+1) feeds step t directly, W^T = F^T V^T feeds Q = C + F^T W,
//     and Qxu feeds V = Qxx + Qxu K - no shared-memory round trip, no layout conversion.
//   * the only fragments read from shared memory are those of F_t (staged by cp.async, two reads per
//     element: once for W^T, once for Q) and the 8-row K/Qux/Quu^-1 panels of the gain computation.
//   * C_t never touches shared memory: it is the accumulator initialiser of Q and is loaded straight
//     into the accumulator registers (LDG.128, one row block ahead).
//
// Per step (reference lqr/lqr_recursion.py:79-152), all by one warp:
//     mv   = V f_t + v                                   (FMA + 2 shuffles)
//     W^T  = F_t^T V^T                       160 DMMA    (A = F fragments, B = V registers)
//     Q    = C_t + F_t^T W, q = c_t + F_t^T mv   200 DMMA (A = F fragments, B = W^T registers), row block u first
//     Quu^-1 by in-register Gauss-Jordan (warp_gj_inverse), K = -Quu^-1 Qux (8 DMMA), k = -Quu^-1 qu
//     V    = Qxx + Qxu K, v = qx + Qxu k      32 DMMA    (A = Qxu registers, accumulate in place on Qxx)
// then (optionally) the rollout x_{t+1} = F_t [x_t; K_t x_t + k_t] + f_t.  Warps never synchronise with each
// other; 8 warps per SM (255 registers each) keep two independent recursions on every DMMA pipe.
#pragma once
#include "common.cuh"
#include "lqr_kernels.cuh"

namespace dmpc {

// DMMA as a pure function of its operands (non-volatile: ptxas may schedule loads around it)
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__device__ __forceinline__ double quad_sum(double a) {      // sum over the 4 lanes t = 0..3 of a row group
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  return a;
}

__device__ __forceinline__ unsigned abs_hi(double x) { return (unsigned)__double2hiint(x) & 0x7fffffffu; }   // |x| high word

// One pivot step of the in-register Gauss-Jordan inverse of Quu (lanes 0..7 hold the columns of Quu, lanes 8..15 the
// columns of I; afterwards lanes 8..15 hold the columns of Quu^-1), branch-free so that the
// compiler can interleave it with the independent DMMA stream of the Q passes.
// 1/x to ~1 ulp: MUFU seed (20 bits), one Newton step (40 bits), one correction in residual form
__device__ __forceinline__ double fast_rcp2(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = r * __fma_rn(-x, r, 2.0);
  return __fma_rn(r, __fma_rn(-x, r, 1.0), r);
}

template <int M>
__device__ __forceinline__ void gj_step(double (&c)[M], const int k) {
  constexpr unsigned FULL = 0xffffffffu;
  double pc[M];
#pragma unroll
  for (int i = 0; i < M; ++i) pc[i] = __shfl_sync(FULL, c[i], k);
  // argmax_{i >= k} |pc[i]| as a max-tree over keys (leading 28 bits of |x|) << 3 | (7 - i): ties (and magnitudes
  // equal to 2^-17 relative) resolve to the smallest row, as in the sequential scan
  unsigned key[M];
#pragma unroll
  for (int i = 0; i < M; ++i) key[i] = (i >= k) ? ((abs_hi(pc[i]) & ~7u) | (unsigned)(M - 1 - i)) : 0u;
#pragma unroll
  for (int w = 1; w < M; w <<= 1)
#pragma unroll
    for (int i = 0; i + w < M; i += 2 * w) key[i] = max(key[i], key[i + w]);
  const int piv = (k == M - 1) ? k : (M - 1 - (int)(key[0] & 7u));
  double ck = c[k], pk = pc[k];
#pragma unroll
  for (int i = 0; i < M; ++i)
    if (i > k) {
      const bool sw = (piv == i);
      ck = sw ? c[i] : ck; pk = sw ? pc[i] : pk;
      c[i] = sw ? c[k] : c[i]; pc[i] = sw ? pc[k] : pc[i];
    }
  const double rp = fast_rcp2(pk);
  ck *= rp;
  c[k] = ck;
#pragma unroll
  for (int i = 0; i < M; ++i)
    if (i != k) c[i] = __fma_rn(-pc[i], ck, c[i]);
}

// I/O element type of the kernel.  double: the tensors are fp64 in HBM.  float (the fp32 API): tensors are float in
// HBM *and* in the staging buffers (half the traffic, half the staging bytes), widened to fp64 when a fragment is read;
// the recursion itself runs on the same fp64 DMMA path, so the fp32 result is the correctly rounded fp64 one.
template <typename IO> __device__ __forceinline__ double2 ld2(const IO* p);
template <> __device__ __forceinline__ double2 ld2<double>(const double* p) { return *reinterpret_cast<const double2*>(p); }
template <> __device__ __forceinline__ double2 ld2<float>(const float* p) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  return make_double2((double)v.x, (double)v.y);
}
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }
__device__ __forceinline__ void st2(float* p, double a, double b) { *reinterpret_cast<float2*>(p) = make_float2((float)a, (float)b); }


struct WarpCfg {
  static constexpr int N = 32, M = 8, S = 40;
  // F buffer: row pitch 42 doubles -> the four fragment rows 2t+e of a half-warp fall in distinct 32-byte
  // bank groups (4*LDF = 8 mod 32).  C buffer: dense (pitch 40), read as accumulator tiles (16 B per lane).
  // float I/O: pitch 44 floats (16-byte aligned rows for cp.async; the rows 2t+e of a fragment read land 24 banks apart).
  static constexpr int LDF = 42;
  template <typename IO> static constexpr int ldf() { return sizeof(IO) == 8 ? LDF : 44; }
  static constexpr int OF = 0, Of = OF + N * LDF, Oc = Of + N, OC = Oc + S, OSCR = OC + S * S;
  // per-warp scratch
  static constexpr int LDQ = 36;     // Qux panel [8][36]: B fragments by rows k0+t
  static constexpr int LDU = 12;     // Quu / Quu^-1 [8][12]: A fragments by rows g
  static constexpr int LDK = 34;     // K panel [8][34]: B fragments by rows 2t+e (aliases the Qux panel)
  static constexpr int OQux = OSCR, OK = OQux, OQuu = OQux + M * LDQ, OQi = OQuu + M * LDU, Oqu = OQi + M * LDU,
                       Okk = Oqu + M, Omv = Okk + M, Obar = Omv + N;                     // two mbarriers per warp
  static constexpr int Oqall = Obar + 2, TOTAL = Oqall + S;                          // q = c + F^T mv, all 40 rows
  // rollout: two stages {F, f, K_t [8][36], k_t} carved from the same region, then x|u and x_next
  static constexpr int LDKR = 36;
  static constexpr int RF = 0, Rf = RF + N * LDF, RK = Rf + N, Rk = RK + M * LDKR, RSTG = Rk + M;
  static constexpr int Oxs = 2 * RSTG, Oxn = Oxs + S;
  static_assert(Oxn + N <= TOTAL, "rollout buffers must fit");
  static_assert(M * LDK <= M * LDQ, "aliases must fit");
  static_assert(2 * (4 * TOTAL * 8 + 1024) <= 228 * 1024, "two 4-warp CTAs per SM");
  static_assert(Of % 2 == 0 && Oc % 2 == 0 && OC % 2 == 0 && OSCR % 2 == 0 && OQuu % 2 == 0 && OQi % 2 == 0 &&
                Oqu % 2 == 0 && Okk % 2 == 0 && RSTG % 2 == 0 && Rf % 2 == 0 && RK % 2 == 0 && Rk % 2 == 0, "16-byte alignment");
};


// ---- bulk asynchronous copies (async proxy, completion on an mbarrier): one instruction moves a whole row /
//      row block, so staging costs neither address registers nor issue slots
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DMPC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DMPC_DONE;\n"
      "bra DMPC_WAIT;\n"
      "DMPC_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// F_t ([32][40] contiguous in global memory) -> padded rows in shared memory: one row per cp.async instruction
// (20 lanes x 16 bytes), so every address is base + compile-time immediate
template <typename IO>
__device__ __forceinline__ void stage_F_rows(IO* Fs, const IO* Fg, int lane) {
  constexpr int EP = 16 / sizeof(IO);              // elements per 16-byte chunk
  if (lane < WarpCfg::S / EP) {
#pragma unroll
    for (int r = 0; r < WarpCfg::N; ++r) cp_async16(Fs + r * WarpCfg::ldf<IO>() + lane * EP, Fg + r * WarpCfg::S + lane * EP);
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

template <int WPC, typename IO>
__global__ void __launch_bounds__(WPC * 32, 8 / WPC) lqr_factor_dmma_warp_kernel(LqrParams<IO> p) {
  using Cfg = WarpCfg;
  constexpr int N = Cfg::N, M = Cfg::M, S = Cfg::S, LDF = Cfg::ldf<IO>(), LDQ = Cfg::LDQ, LDU = Cfg::LDU, LDK = Cfg::LDK;
  constexpr int W = sizeof(IO), EP = 16 / W;        // bytes per I/O element, elements per 16-byte chunk
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gr = lane >> 2, tg = lane & 3;
  const int T = p.T, B = p.B;
  // Persistent warps: at most two CTAs per SM are launched and warp gw owns the elements gw, gw + nw, gw + 2 nw, ...
  // (no CTA relaunch between elements, no idle warps in a partially filled last wave: -3.5 % on config 5).  Skewing the
  // sweep / rollout order between the warps of a CTA (sweep k elements, then roll k out) was measured too and is within
  // noise of this plain order (profiles/r2/forward_variants_ab.txt).
  const int nw = gridDim.x * WPC, gw = blockIdx.x * WPC + warp;
  if (gw >= B) return;                      // warps never synchronise CTA-wide below this point
  double* sm = reinterpret_cast<double*>(smem_raw) + (size_t)warp * Cfg::TOTAL;
  const size_t tb = (size_t)B;
  const bool have_f = p.f != nullptr;
  const bool save_fac = (p.flags & LQR_SAVE_FAC) && p.fac;

  // two mbarriers per warp (Riccati: {F,f,c} and C; rollout: one per stage); every phase opened below is waited for
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + Cfg::Obar);
  if (lane == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncwarp();
  unsigned par0 = 0, par1 = 0;

  for (int e = gw; e < B; e += nw) {
  // The staging region is re-used from the rollout of the previous element (generic-proxy reads and writes) by the bulk
  // copies of this sweep (async proxy): order them once per element.
  __syncwarp();
  fence_proxy_async();
  if (p.flags & LQR_DO_FACTOR) {
    IO* Fs = reinterpret_cast<IO*>(sm + Cfg::OF); IO* fs = reinterpret_cast<IO*>(sm + Cfg::Of);
    IO* cs = reinterpret_cast<IO*>(sm + Cfg::Oc); IO* Cs = reinterpret_cast<IO*>(sm + Cfg::OC);
    double* Qux_s = sm + Cfg::OQux; double* Quu_s = sm + Cfg::OQuu; double* Qi_s = sm + Cfg::OQi;
    double* K_s = sm + Cfg::OK; double* qu_s = sm + Cfg::Oqu; double* kk_s = sm + Cfg::Okk; double* mv_s = sm + Cfg::Omv;

    // Single-buffered, refilled just in time: a row block of C_{t-1} is requested (bulk copy) as soon as pass i of
    // step t has read its accumulators, a full step ahead of its use; F_{t-1} (padded rows, cp.async), f_{t-1},
    // c_{t-1} are requested after the last Q pass and land behind the gain computation.  (An L2 bulk prefetch one
    // step ahead was measured and removed: the prefetched lines did not survive until their use and were read
    // twice from DRAM - +0.77 MB per solve, 6 % slower, profiles/r1/l2_prefetch_ab.txt.)
    unsigned long long* barF = bars;
    unsigned long long* barC = bars + 1;
    auto stage_C_rows = [&](int t, int i) {            // lane 0 only
      bulk_g2s(Cs + i * 8 * S, p.C + ((size_t)t * tb + e) * (S * S) + i * 8 * S, 8 * S * W, barC);
    };
    auto stage_Ffc = [&](int t) {                      // all lanes; t < T-1
      const size_t idx = (size_t)t * tb + e;
      stage_F_rows(Fs, p.F + idx * (N * S), lane);
      cp_async_commit();
      if (lane == 0) {
        mbar_arrive_expect_tx(barF, (have_f ? N * W : 0) + S * W);
        bulk_g2s(cs, p.c + idx * S, S * W, barF);
        if (have_f) bulk_g2s(fs, p.f + idx * N, N * W, barF);
      }
    };

    // V_t as accumulator-layout registers: Vr[r][kb][e] = V[8r+g][8kb+2t+e];  vr[r] = v[8r+g]
    // v_t and q_x live in shared memory between their producer and consumer (v aliases mv, q_x aliases Quu)
    double Vr[4][4][2];
    double* v_s = mv_s; double* qx_s = Quu_s;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) { Vr[r][kb][0] = 0.0; Vr[r][kb][1] = 0.0; }
    }

    if (lane == 0) {                                   // step T-1 needs C and c only
      mbar_arrive_expect_tx(barC, S * S * W);
      bulk_g2s(Cs, p.C + ((size_t)(T - 1) * tb + e) * (S * S), S * S * W, barC);
      mbar_arrive_expect_tx(barF, S * W);
      bulk_g2s(cs, p.c + ((size_t)(T - 1) * tb + e) * S, S * W, barF);
    }

    for (int t = T - 1; t >= 0; --t) {
      cp_async_wait<0>();
      mbar_wait(barF, par0); par0 ^= 1;
      mbar_wait(barC, par1); par1 ^= 1;
      __syncwarp();                                    // C_t, F_t, f_t, c_t are resident
      const size_t idx = (size_t)t * tb + e;
      const bool last = (t == T - 1);

      double WT[5][4][2];          // WT[j][r][e] = W[8r+2t+e][8j+g],  W = V F
      if (!last) {
        // ---- mv = V f + v
        double mvr[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          double a = 0.0;
          if (have_f) {
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
              const double2 f2 = ld2<IO>(fs + kb * 8 + 2 * tg);
              a = __fma_rn(Vr[r][kb][0], f2.x, a);
              a = __fma_rn(Vr[r][kb][1], f2.y, a);
            }
            a = quad_sum(a);
          }
          mvr[r] = a + v_s[r * 8 + gr];
        }
        __syncwarp();                                  // every lane has read v before mv overwrites it
        if (tg == 0) {
#pragma unroll
          for (int r = 0; r < 4; ++r) mv_s[r * 8 + gr] = mvr[r];
        }
        // ---- W^T = F^T V^T : tile (j, r) = sum_k F[k][8j+g] * V[8r+g'][k],  k = 8kb+2t+e
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          double a[4][2];
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            a[kb][0] = Fs[(kb * 8 + 2 * tg) * LDF + j * 8 + gr];
            a[kb][1] = Fs[(kb * 8 + 2 * tg + 1) * LDF + j * 8 + gr];
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) { WT[j][r][0] = 0.0; WT[j][r][1] = 0.0; }
#pragma unroll
          for (int kb = 0; kb < 4; ++kb)
#pragma unroll
            for (int ee = 0; ee < 2; ++ee)
#pragma unroll
              for (int r = 0; r < 4; ++r) dmma(WT[j][r], a[kb][ee], Vr[r][kb][ee]);
        }
        __syncwarp();                                  // mv_s complete
      }

      // ---- q = c + F^T mv for all 40 rows as ONE scalar section.  DFMAs share the FP64 pipe with the DMMAs: issued
      //      between them (one per k-step, as round 1 did) each one waits for the DMMA in flight and costs a whole DMMA
      //      slot (profiles/r2/fwd_pass_stalls.txt); issued back to back here they do not break up the DMMA streams of
      //      the Q passes.  (Folding q, mv and v into extra DMMA tiles instead - 480 DMMAs per step - measured slower:
      //      profiles/r2/forward_variants_ab.txt.)
      double* qall_s = sm + Cfg::Oqall;
      if (!last) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;           // rows 0..31: lane l owns row l (column l of F_t)
#pragma unroll
        for (int k = 0; k < N; k += 4) {
          a0 = __fma_rn((double)Fs[(k + 0) * LDF + lane], mv_s[k + 0], a0);
          a1 = __fma_rn((double)Fs[(k + 1) * LDF + lane], mv_s[k + 1], a1);
          a2 = __fma_rn((double)Fs[(k + 2) * LDF + lane], mv_s[k + 2], a2);
          a3 = __fma_rn((double)Fs[(k + 3) * LDF + lane], mv_s[k + 3], a3);
        }
        double b0 = 0.0;                                           // rows 32..39: row 32 + g by the four lanes of a quad
#pragma unroll
        for (int k = 0; k < N; k += 4) b0 = __fma_rn((double)Fs[(k + tg) * LDF + N + gr], mv_s[k + tg], b0);
        b0 = quad_sum(b0);
        qall_s[lane] = ((a0 + a1) + (a2 + a3)) + (double)cs[lane];
        if (tg == 0) qall_s[N + gr] = b0 + (double)cs[N + gr];
      } else {
        qall_s[lane] = (double)cs[lane];
        if (lane < M) qall_s[N + lane] = (double)cs[N + lane];
      }
      __syncwarp();

      // ---- Q = C + F^T W, row block u first.  After that pass the Gauss-Jordan inverse of Quu, K = -Quu^-1 Qux and
      //      k = -Quu^-1 qu run as a second scalar section; the four Qxx passes that follow are pure DMMA streams.
      double Qxx[4][4][2], Qxu[4][2];
      IO* fg = save_fac ? p.fac + idx * (M * M + N * M) : nullptr;
#pragma unroll
      for (int pi = 0; pi < 5; ++pi) {
        const int i = (pi == 0) ? 4 : pi - 1;
        double acc[5][2];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const double2 c2 = ld2<IO>(Cs + (i * 8 + gr) * S + j * 8 + 2 * tg);
          acc[j][0] = c2.x; acc[j][1] = c2.y;
        }
        if (pi == 1) {
          double cinv[M];
#pragma unroll
          for (int ii = 0; ii < M; ++ii) cinv[ii] = (lane < M) ? Quu_s[ii * LDU + lane] : ((lane - M == ii) ? 1.0 : 0.0);
#pragma unroll
          for (int k = 0; k < M; ++k) gj_step<M>(cinv, k);
          if (lane >= M && lane < 2 * M) {             // lanes 8..15 hold the columns of Quu^-1
#pragma unroll
            for (int ii = 0; ii < M; ++ii) { Qi_s[ii * LDU + lane - M] = cinv[ii]; if (fg) fg[ii * M + lane - M] = (IO)cinv[ii]; }
          }
          __syncwarp();                                // Quu^-1 complete (q_x, stored by the passes below, aliases the Quu panel)
          double Kt[4][2];
#pragma unroll
          for (int c = 0; c < 4; ++c) { Kt[c][0] = 0.0; Kt[c][1] = 0.0; }
#pragma unroll
          for (int k0 = 0; k0 < M; k0 += 4) {
            const double a = Qi_s[gr * LDU + k0 + tg];
#pragma unroll
            for (int c = 0; c < 4; ++c) dmma(Kt[c], a, Qux_s[(k0 + tg) * LDQ + c * 8 + gr]);
          }
          const double2 qi2 = *reinterpret_cast<const double2*>(Qi_s + gr * LDU + 2 * tg);
          const double2 qu2 = *reinterpret_cast<const double2*>(qu_s + 2 * tg);
          const double kk = -quad_sum(__fma_rn(qi2.x, qu2.x, qi2.y * qu2.y));
          if (tg == 0) { kk_s[gr] = kk; p.ks[idx * M + gr] = (IO)kk; }
          __syncwarp();                                // every lane has read the Qux panel: K may overwrite it
          IO* Kg = p.Ks + idx * (M * N) + gr * N + 2 * tg;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const double2 k2 = make_double2(-Kt[c][0], -Kt[c][1]);
            *reinterpret_cast<double2*>(K_s + gr * LDK + c * 8 + 2 * tg) = k2;
            st2(Kg + c * 8, k2.x, k2.y);
          }
        }
        if (!last) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int ee = 0; ee < 2; ++ee) {
              const double a = Fs[(r * 8 + 2 * tg + ee) * LDF + i * 8 + gr];
#pragma unroll
              for (int j = 0; j < 5; ++j) dmma(acc[j], a, WT[j][r][ee]);
            }
        }
        // Refill this row block of the C buffer for the next step (async proxy, bulk copy).  It is issued BEHIND the DMMAs
        // of the pass: they read every lane's accumulator registers, so every lane's loads of the row block above have
        // completed when the copy is issued.  Round 1 issued it right after the loads with only a warp barrier in between
        // (loads are ISSUED then, not complete): once every ~1e6 element-steps the refill overtook a load and one element's
        // recursion went wrong (scratch/stress_c5.py, profiles/r2/forward_race.txt).  The last horizon step has no DMMAs:
        // there the loads are ordered by an explicit proxy fence (once per element).
        if (t > 0) {
          if (last) { fence_proxy_async(); __syncwarp(); }
          if (lane == 0) {
            if (pi == 0) mbar_arrive_expect_tx(barC, S * S * W);
            stage_C_rows(t - 1, i);
          }
        }
        const double qa = qall_s[i * 8 + gr];
        if (i == 4) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2*>(Qux_s + gr * LDQ + j * 8 + 2 * tg) = make_double2(acc[j][0], acc[j][1]);
          *reinterpret_cast<double2*>(Quu_s + gr * LDU + 2 * tg) = make_double2(acc[4][0], acc[4][1]);
          if (tg == 0) qu_s[gr] = qa;
          __syncwarp();                                // the Qux | Quu | qu panels are complete
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) { Qxx[i][j][0] = acc[j][0]; Qxx[i][j][1] = acc[j][1]; }
          Qxu[i][0] = acc[4][0]; Qxu[i][1] = acc[4][1];
          if (tg == 0) qx_s[i * 8 + gr] = qa;
        }
      }
      __syncwarp();                                    // F_t, f_t, c_t, mv are dead; K, k, q_x are complete
      if (t > 0) stage_Ffc(t - 1);
      if (fg) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
          st2(fg + M * M + (r * 8 + gr) * M + 2 * tg, Qxu[r][0], Qxu[r][1]);
      }
      __syncwarp();
      // ---- V = Qxx + Qxu K, v = qx + Qxu k.  The reference's extra terms K^T (Qux + Quu K) and K^T (qu + Quu k)
      //      (lqr_recursion.py:151-152) vanish identically for the exact gain.
      if (t > 0) {
        const double2 kp = *reinterpret_cast<const double2*>(kk_s + 2 * tg);
#pragma unroll
        for (int ee = 0; ee < 2; ++ee)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const double b = K_s[(2 * tg + ee) * LDK + c * 8 + gr];
#pragma unroll
            for (int r = 0; r < 4; ++r) dmma(Qxx[r][c], Qxu[r][ee], b);
          }
        IO* Vg = p.Vsave ? p.Vsave + idx * (N * N + N) : nullptr;     // V_t | v_t for the fused adjoint (lambda = V x + v)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const double vn = qx_s[r * 8 + gr] + quad_sum(__fma_rn(Qxu[r][0], kp.x, Qxu[r][1] * kp.y));
          if (tg == 0) { v_s[r * 8 + gr] = vn; if (Vg) Vg[N * N + r * 8 + gr] = (IO)vn; }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            Vr[r][c][0] = Qxx[r][c][0]; Vr[r][c][1] = Qxx[r][c][1];
            if (Vg) st2(Vg + (r * 8 + gr) * N + c * 8 + 2 * tg, Vr[r][c][0], Vr[r][c][1]);
          }
        }
      }
    }
  }

  // ---- rollout (lqr_recursion.py:160-200): the warp re-streams K_t, k_t, F_t, f_t through two cp.async stages.  The
  //      steps are short, so this phase is latency/HBM bound (10 KB per step); other warps of the SM are in their
  //      DMMA-bound Riccati sweep meanwhile
  if (p.flags & LQR_DO_ROLLOUT) {
    constexpr int LDKR = Cfg::LDKR;
    __threadfence_block();
    __syncwarp();                                  // K_t, k_t written above by other lanes of this warp
    double* xs = sm + Cfg::Oxs;                    // [x; u]
    double* xn = sm + Cfg::Oxn;
    __syncwarp();
    auto load_roll = [&](int t, int st) {
      double* base = sm + st * Cfg::RSTG;
      const size_t idx = (size_t)t * tb + e;
      const bool hasF = t < T - 1;
      if (hasF) stage_F_rows<IO>(reinterpret_cast<IO*>(base + Cfg::RF), p.F + idx * (N * S), lane);
      const IO* Kg = p.Ks + idx * (M * N);
      constexpr int CPR = N / EP;                  // 16-byte chunks per row of K (fp64: 128 chunks in all, 16 per row)
#pragma unroll
      for (int it = 0; it < M * CPR / 32; ++it) {
        const int q = it * 32 + lane, row = q / CPR, cc = q % CPR;
        cp_async16(reinterpret_cast<IO*>(base + Cfg::RK) + row * LDKR + cc * EP, Kg + q * EP);
      }
      if (lane < M / EP) cp_async16(reinterpret_cast<IO*>(base + Cfg::Rk) + lane * EP, p.ks + idx * M + lane * EP);
      if (hasF && have_f && lane >= 16 && lane - 16 < N / EP)
        cp_async16(reinterpret_cast<IO*>(base + Cfg::Rf) + (lane - 16) * EP, p.f + idx * N + (lane - 16) * EP);
      cp_async_commit();
    };
    load_roll(0, 0);
    xs[lane] = (double)p.x0[(size_t)e * N + lane];
    int st = 0;
    for (int t = 0; t < T; ++t) {
      cp_async_wait<0>();
      __syncwarp();
      if (t + 1 < T) load_roll(t + 1, st ^ 1);
      const IO* Fs = reinterpret_cast<const IO*>(sm + st * Cfg::RSTG + Cfg::RF);
      const IO* fs = reinterpret_cast<const IO*>(sm + st * Cfg::RSTG + Cfg::Rf);
      const IO* Kt = reinterpret_cast<const IO*>(sm + st * Cfg::RSTG + Cfg::RK);
      const IO* kt = reinterpret_cast<const IO*>(sm + st * Cfg::RSTG + Cfg::Rk);
      {   // u = K x + k : control g, interleaved quarter of the states per lane (bank-conflict free with LDKR = 36)
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < N / 4; ++i) a = __fma_rn(Kt[gr * LDKR + 4 * i + tg], xs[4 * i + tg], a);
        a = quad_sum(a);
        if (tg == 0) xs[N + gr] = a + kt[gr];
      }
      __syncwarp();
      const size_t idx = (size_t)t * tb + e;
      if (p.x) p.x[idx * N + lane] = (IO)xs[lane];
      if (p.u && lane < M) p.u[idx * M + lane] = (IO)xs[N + lane];
      if (p.tau_out) { p.tau_out[idx * S + lane] = (IO)xs[lane]; if (lane < M) p.tau_out[idx * S + N + lane] = (IO)xs[N + lane]; }
      if (t < T - 1) {   // x' = F [x; u] + f : one state per lane, four accumulation chains
        double a0 = have_f ? fs[lane] : 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int j = 0; j < S; j += 4) {
          a0 = __fma_rn(Fs[lane * LDF + j], xs[j], a0);
          a1 = __fma_rn(Fs[lane * LDF + j + 1], xs[j + 1], a1);
          a2 = __fma_rn(Fs[lane * LDF + j + 2], xs[j + 2], a2);
          a3 = __fma_rn(Fs[lane * LDF + j + 3], xs[j + 3], a3);
        }
        xn[lane] = (a0 + a1) + (a2 + a3);
        __syncwarp();
        xs[lane] = xn[lane];
      }
      st ^= 1;
    }
  }
  }   // elements of this warp
}

}  // namespace dmpc
