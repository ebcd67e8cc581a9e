#pragma once
#include "mpc_kernels.cuh"
namespace dmpc {
template <typename R> int launch_mpc_forward(const MpcFwdParams<R>& p, cudaStream_t st, long long* nl);
template <typename R> int launch_pnqp(const PnqpParams<R>& p, cudaStream_t st, long long* nl);
template <typename R> int launch_active_mask(const R* u, const R* lo, const R* hi, unsigned char* out, size_t count,
                                             cudaStream_t st, long long* nl);
template <typename R> int launch_traj(const TrajParams<R>& p, cudaStream_t st, long long* nl);
}  // namespace dmpc
