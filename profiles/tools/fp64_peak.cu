// FP64 peak microbenchmark for the roofline denominators ("builder must
// microbenchmark DFMA/DMMA peaks").  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// Prints one JSON line: DFMA and DMMA (mma.sync.m8n8k4.f64) TFLOP/s and SM clock.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a[8];
  const double b = 1.0000001, c = 1e-9;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = __fma_rn(a[i], b, c);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms; double best_fma = 0, best_mma = 0;
  for (int tpb = 128; tpb <= 1024; tpb *= 2) {
    for (int bps = 1; bps <= 2; ++bps) {
      const int iters = 20000;
      dfma_kernel<<<sms * bps, tpb>>>(out, 1000);
      cudaEventRecord(e0); dfma_kernel<<<sms * bps, tpb>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      double tf = 2.0 * 64 * iters * (double)tpb * sms * bps / (ms * 1e-3) / 1e12;
      if (tf > best_fma) best_fma = tf;
      dmma_kernel<<<sms * bps, tpb>>>(out, 1000);
      cudaEventRecord(e0); dmma_kernel<<<sms * bps, tpb>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      double tm = 2.0 * 256 * 8 * iters * (double)(tpb / 32) * sms * bps / (ms * 1e-3) / 1e12;
      if (tm > best_mma) best_mma = tm;
      fprintf(stderr, "tpb %d bps %d: dfma %.2f TF, dmma %.2f TF\n", tpb, bps, tf, tm);
    }
  }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.3f, \"dmma_m8n8k4_tflops\": %.3f, \"sm_clock_khz_max\": %d}\n",
         p.name, sms, best_fma, best_mma, clk);
  return 0;
}
