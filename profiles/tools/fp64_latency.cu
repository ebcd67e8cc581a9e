// Dependent-issue latencies on B200 (clock64 around a single-warp dependent chain):
// DFMA, DMMA m8n8k4 (1, 2, 4, 8 independent accumulators), 64-bit SHFL, MUFU.RCP64H + Newton.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(long long* out, double* sink) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 0.999999, c = 1e-12;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < 256; ++i) a = __fma_rn(a, b, c);
  t1 = clock64(); out[0] = (t1 - t0);
  // DMMA chains
  double c0[8][2];
  for (int i = 0; i < 8; ++i) { c0[i][0] = 0; c0[i][1] = 0; }
  double x = 1.0 + threadIdx.x * 1e-6, y = 1.0 - threadIdx.x * 1e-6;
#define RUN(NACC, SLOT) \
  t0 = clock64(); \
  _Pragma("unroll 4") for (int i = 0; i < 128; ++i) { \
    _Pragma("unroll") for (int j = 0; j < NACC; ++j) \
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0[j][0]), "+d"(c0[j][1]) : "d"(x), "d"(y)); \
  } \
  t1 = clock64(); out[SLOT] = (t1 - t0);
  RUN(1, 1) RUN(2, 2) RUN(4, 3) RUN(8, 4)
  // 64-bit shuffle chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 256; ++i) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
  t1 = clock64(); out[5] = (t1 - t0);
  // rcp seed + 3 newton
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < 64; ++i) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); r = r * __fma_rn(-a, r, 2.0); r = r * __fma_rn(-a, r, 2.0); r = __fma_rn(r, __fma_rn(-a, r, 1.0), r); a = r + 1.5; }
  t1 = clock64(); out[6] = (t1 - t0);
  double s = a; for (int i = 0; i < 8; ++i) s += c0[i][0] + c0[i][1];
  sink[threadIdx.x] = s;
}
int main() {
  long long* d; double* s; cudaMalloc(&d, 64); cudaMalloc(&s, 32 * 8);
  lat<<<1, 32>>>(d, s); lat<<<1, 32>>>(d, s); cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d, 56, cudaMemcpyDeviceToHost);
  printf("{\"dfma_dep_clk\": %.1f, \"dmma_dep_clk_1acc\": %.1f, \"dmma_clk_per_op_2acc\": %.1f, \"dmma_clk_per_op_4acc\": %.1f, \"dmma_clk_per_op_8acc\": %.1f, \"shfl64_dep_clk\": %.1f, \"rcp_newton3_plus_add_clk\": %.1f}\n",
         h[0] / 256.0, h[1] / 128.0, h[2] / 256.0, h[3] / 512.0, h[4] / 1024.0, h[5] / 256.0, h[6] / 64.0);
  return 0;
}
