// extern "C" boundary of libdiffmpc_b200.so (declared in include/diffmpc_b200.h).
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/diffmpc_b200.h"
#include "launch.h"

struct dmpc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  std::string err;
};

using namespace dmpc;

namespace {
inline cudaStream_t pick(dmpc_handle h, void* s) { return s ? (cudaStream_t)s : h->stream; }
inline int fail(dmpc_handle h, int code, const char* what) {
  if (h) h->err = what;
  return code;
}
inline int cuda_fail(dmpc_handle h, cudaError_t e, const char* where) {
  if (h) h->err = std::string(where) + ": " + cudaGetErrorString(e);
  return DMPC_ERR_CUDA;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(h, e_, #call); } while (0)
inline int set_dev(dmpc_handle h) { return cudaSetDevice(h->device) == cudaSuccess ? 0 : DMPC_ERR_CUDA; }
}  // namespace

template <typename R>
static int lqr_solve_impl(dmpc_handle h, int T, int B, int n, int m, const void* x0, const void* C, const void* c,
                          const void* F, const void* f, void* x, void* u, void* Ks, void* ks, void* fac, int flags,
                          cudaStream_t st) {
  LqrParams<R> p;
  memset(&p, 0, sizeof(p));
  p.T = T; p.B = B; p.n = n; p.m = m;
  p.flags = 0;
  if (flags & DMPC_LQR_FACTOR) p.flags |= LQR_DO_FACTOR;
  if (flags & DMPC_LQR_ROLLOUT) p.flags |= LQR_DO_ROLLOUT;
  if ((flags & DMPC_LQR_SAVE_FAC) && fac) p.flags |= LQR_SAVE_FAC;
  p.x0 = (const R*)x0; p.C = (const R*)C; p.c = (const R*)c; p.F = (const R*)F; p.f = (const R*)f;
  p.c_scale = R(1);
  p.x = (R*)x; p.u = (R*)u; p.Ks = (R*)Ks; p.ks = (R*)ks; p.fac = (R*)fac;
  int rc = launch_lqr_solve<R>(p, st, &h->launches);
  if (rc) h->err = "lqr_solve launch failed";
  return rc;
}

template <typename R>
static int lqr_adjoint_impl(dmpc_handle h, int T, int B, int n, int m, const void* C, const void* c, const void* F,
                            const void* x, const void* u, const void* gx, const void* gu, const void* Ks,
                            const void* fac, void* dx0, void* dC, void* dc, void* dF, void* df, int flags,
                            cudaStream_t st) {
  DtauParams<R> d;
  d.T = T; d.B = B; d.n = n; d.m = m;
  d.F = (const R*)F; d.gx = (const R*)gx; d.gu = (const R*)gu; d.Ks = (const R*)Ks; d.fac = (const R*)fac;
  d.dc = (R*)dc;
  int rc = launch_lqr_dtau<R>(d, st, &h->launches);
  if (rc) { h->err = "lqr_dtau launch failed"; return rc; }
  AdjOutParams<R> a;
  memset(&a, 0, sizeof(a));
  a.T = T; a.B = B; a.n = n; a.m = m; a.F_T = T - 1;
  a.flags = (flags & DMPC_ADJ_STRICT_REFERENCE) ? (ADJ_QUIRK_DC | ADJ_QUIRK_DF) : 0;
  a.C = (const R*)C; a.c = (const R*)c; a.F = (const R*)F; a.x = (const R*)x; a.u = (const R*)u;
  a.dtau = (const R*)dc; a.gx = (const R*)gx; a.gu = (const R*)gu;
  a.dx0 = (R*)dx0; a.dC = (R*)dC; a.dc = (R*)dc; a.dF = (R*)dF; a.df = (R*)df;
  rc = launch_adjoint_out<R>(a, st, &h->launches);
  if (rc) h->err = "adjoint_out launch failed";
  return rc;
}

extern "C" {

int dmpc_version(void) { return 100; }

const char* dmpc_status_string(int s) {
  switch (s) {
    case DMPC_OK: return "ok";
    case DMPC_ERR_BAD_SHAPE: return "bad shape";
    case DMPC_ERR_BAD_BOUNDS: return "lower is larger than upper";
    case DMPC_ERR_NONFINITE: return "non-finite value";
    case DMPC_ERR_CUDA: return "CUDA error";
    case DMPC_ERR_UNSUPPORTED: return "unsupported configuration";
    case DMPC_ERR_NULL: return "null pointer";
    case DMPC_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    default: return "unknown status";
  }
}

int dmpc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int dmpc_create(int device, dmpc_handle* out) {
  if (!out) return DMPC_ERR_NULL;
  *out = nullptr;
  int n = dmpc_device_count();
  if (n <= 0 || device < 0 || device >= n) return DMPC_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return DMPC_ERR_CUDA;
  dmpc_ctx* h = new dmpc_ctx();
  h->device = device;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return DMPC_ERR_CUDA; }
  *out = h;
  return DMPC_OK;
}

int dmpc_destroy(dmpc_handle h) {
  if (!h) return DMPC_ERR_NULL;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return DMPC_OK;
}

const char* dmpc_last_error(dmpc_handle h) { return h ? h->err.c_str() : "null handle"; }
long long dmpc_launch_count(dmpc_handle h) { return h ? h->launches : 0; }

int dmpc_malloc(dmpc_handle h, size_t bytes, void** d_ptr) {
  if (!h || !d_ptr) return DMPC_ERR_NULL;
  if (set_dev(h)) return DMPC_ERR_CUDA;
  CK(cudaMalloc(d_ptr, bytes ? bytes : 16));
  return DMPC_OK;
}
int dmpc_free(dmpc_handle h, void* d_ptr) {
  if (!h) return DMPC_ERR_NULL;
  if (set_dev(h)) return DMPC_ERR_CUDA;
  CK(cudaFree(d_ptr));
  return DMPC_OK;
}
int dmpc_host_alloc(dmpc_handle h, size_t bytes, void** h_ptr) {
  if (!h || !h_ptr) return DMPC_ERR_NULL;
  if (set_dev(h)) return DMPC_ERR_CUDA;
  CK(cudaHostAlloc(h_ptr, bytes ? bytes : 16, cudaHostAllocDefault));
  return DMPC_OK;
}
int dmpc_host_free(dmpc_handle h, void* h_ptr) {
  if (!h) return DMPC_ERR_NULL;
  CK(cudaFreeHost(h_ptr));
  return DMPC_OK;
}
int dmpc_memcpy_h2d(dmpc_handle h, void* d, const void* s, size_t bytes, void* stream) {
  if (!h) return DMPC_ERR_NULL;
  if (set_dev(h)) return DMPC_ERR_CUDA;
  if (bytes) CK(cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, pick(h, stream)));
  return DMPC_OK;
}
int dmpc_memcpy_d2h(dmpc_handle h, void* d, const void* s, size_t bytes, void* stream) {
  if (!h) return DMPC_ERR_NULL;
  if (set_dev(h)) return DMPC_ERR_CUDA;
  if (bytes) CK(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToHost, pick(h, stream)));
  return DMPC_OK;
}
int dmpc_memset(dmpc_handle h, void* d, int value, size_t bytes, void* stream) {
  if (!h) return DMPC_ERR_NULL;
  if (set_dev(h)) return DMPC_ERR_CUDA;
  if (bytes) CK(cudaMemsetAsync(d, value, bytes, pick(h, stream)));
  return DMPC_OK;
}
int dmpc_sync(dmpc_handle h, void* stream) {
  if (!h) return DMPC_ERR_NULL;
  if (set_dev(h)) return DMPC_ERR_CUDA;
  CK(cudaStreamSynchronize(pick(h, stream)));
  return DMPC_OK;
}

size_t dmpc_lqr_fac_elems(int T, int B, int n, int m) { return (size_t)T * B * (size_t)(m * m + n * m); }

int dmpc_lqr_solve(dmpc_handle h, int dtype, int T, int B, int n, int m, const void* d_x0, const void* d_C,
                   const void* d_c, const void* d_F, int F_T, const void* d_f, void* d_x, void* d_u, void* d_Ks,
                   void* d_ks, void* d_fac, int flags, void* stream) {
  if (!h) return DMPC_ERR_NULL;
  if (T < 1 || B < 1 || n < 1 || m < 1) return fail(h, DMPC_ERR_BAD_SHAPE, "T,B,n,m must be >= 1");
  if (T > 1 && F_T != T - 1 && F_T != T) return fail(h, DMPC_ERR_BAD_SHAPE, "F must have T-1 or T time rows");
  if (!d_Ks || !d_ks) return fail(h, DMPC_ERR_NULL, "Ks/ks buffers are required");
  if ((flags & DMPC_LQR_FACTOR) && (!d_C || !d_c || (T > 1 && !d_F))) return fail(h, DMPC_ERR_NULL, "C,c,F required");
  if ((flags & DMPC_LQR_ROLLOUT) && (!d_x0 || !d_x || !d_u || (T > 1 && !d_F))) return fail(h, DMPC_ERR_NULL, "x0,x,u,F required");
  if (!(flags & (DMPC_LQR_FACTOR | DMPC_LQR_ROLLOUT))) return fail(h, DMPC_ERR_UNSUPPORTED, "nothing to do");
  if (set_dev(h)) return DMPC_ERR_CUDA;
  cudaStream_t st = pick(h, stream);
  if (dtype == DMPC_F64) return lqr_solve_impl<double>(h, T, B, n, m, d_x0, d_C, d_c, d_F, d_f, d_x, d_u, d_Ks, d_ks, d_fac, flags, st);
  if (dtype == DMPC_F32) return lqr_solve_impl<float>(h, T, B, n, m, d_x0, d_C, d_c, d_F, d_f, d_x, d_u, d_Ks, d_ks, d_fac, flags, st);
  return fail(h, DMPC_ERR_UNSUPPORTED, "dtype");
}

int dmpc_lqr_adjoint(dmpc_handle h, int dtype, int T, int B, int n, int m, const void* d_C, const void* d_c,
                     const void* d_F, const void* d_x, const void* d_u, const void* d_gx, const void* d_gu,
                     const void* d_Ks, const void* d_fac, void* d_dx0, void* d_dC, void* d_dc, void* d_dF,
                     void* d_df, int flags, void* stream) {
  if (!h) return DMPC_ERR_NULL;
  if (T < 1 || B < 1 || n < 1 || m < 1) return fail(h, DMPC_ERR_BAD_SHAPE, "T,B,n,m must be >= 1");
  if (!d_C || !d_c || !d_x || !d_u || !d_gx || !d_gu || !d_Ks || !d_fac || !d_dx0 || !d_dC || !d_dc || (T > 1 && (!d_F || !d_dF)))
    return fail(h, DMPC_ERR_NULL, "lqr_adjoint: required buffer is NULL");
  if (set_dev(h)) return DMPC_ERR_CUDA;
  cudaStream_t st = pick(h, stream);
  if (dtype == DMPC_F64) return lqr_adjoint_impl<double>(h, T, B, n, m, d_C, d_c, d_F, d_x, d_u, d_gx, d_gu, d_Ks, d_fac, d_dx0, d_dC, d_dc, d_dF, d_df, flags, st);
  if (dtype == DMPC_F32) return lqr_adjoint_impl<float>(h, T, B, n, m, d_C, d_c, d_F, d_x, d_u, d_gx, d_gu, d_Ks, d_fac, d_dx0, d_dC, d_dc, d_dF, d_df, flags, st);
  return fail(h, DMPC_ERR_UNSUPPORTED, "dtype");
}

}  // extern "C"
