// Host-side launch interfaces between capi.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include "lqr_kernels.cuh"

namespace dmpc {

// Which (n, m) shapes have compile-time-specialised kernels (everything else runs the
// runtime-shape instantiation of the same code).
struct ShapeInfo { int n, m, G; bool specialised; };

template <typename R> int launch_lqr_solve(const LqrParams<R>& p, cudaStream_t st, long long* nlaunch);
template <typename R> int launch_lqr_dtau(const DtauParams<R>& p, cudaStream_t st, long long* nlaunch);
template <typename R> int launch_adjoint_out(const AdjOutParams<R>& p, cudaStream_t st, long long* nlaunch);
template <typename R> int launch_reduce_partials(const R* red, int B, int rsz, R* out, cudaStream_t st, long long* nlaunch);

constexpr int kMaxSmem = 227 * 1024;

}  // namespace dmpc
