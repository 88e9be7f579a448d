// Host side of the batched dense Newton step (ks_dense.cuh): the strided-batched fp64 GEMM.
// Product build: cuBLAS, resolved at run time with dlopen("libcublas.so.12") -- the copy PyTorch has already mapped when the
// library is used from Python, the system one otherwise -- so libks_b200.so has no link-time dependency on it.  cuBLAS is
// used for these plain GEMMs only (J J^T, the Cholesky panel and trailing updates); everything else is our own kernels.
// Emulation build (tests/emul): a naive host loop.
#pragma once
#include "ks_host.cuh"

#ifndef KS_HOST_EMUL
#include <cublas_v2.h>
#include <dlfcn.h>
#endif

namespace ksdh {
using namespace ksh;

#ifdef KS_HOST_EMUL
// C = alpha op(A) op(B) + beta C, column-major, one matrix per batch entry
inline int dgemm_batched(cudaStream_t, bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda, long long sa,
                         const double* B, int ldb, long long sb, double beta, double* C, int ldc, long long sc, int batch) {
  for (int q = 0; q < batch; ++q) {
    const double* a = A + (size_t)q * sa;
    const double* b = B + (size_t)q * sb;
    double* c = C + (size_t)q * sc;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int l = 0; l < k; ++l) s += (ta ? a[(size_t)i * lda + l] : a[(size_t)l * lda + i]) * (tb ? b[(size_t)l * ldb + j] : b[(size_t)j * ldb + l]);
        c[(size_t)j * ldc + i] = alpha * s + (beta == 0.0 ? 0.0 : beta * c[(size_t)j * ldc + i]);
      }
  }
  return 0;
}
#else
struct CublasApi {
  void* lib = nullptr;
  cublasStatus_t (*create)(cublasHandle_t*) = nullptr;
  cublasStatus_t (*set_stream)(cublasHandle_t, cudaStream_t) = nullptr;
  cublasStatus_t (*set_workspace)(cublasHandle_t, void*, size_t) = nullptr;
  cublasStatus_t (*dgemm_sb)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const double*, const double*, int,
                             long long, const double*, int, long long, const double*, double*, int, long long, int) = nullptr;
  cublasHandle_t handle[kMaxDevices] = {};
};
inline CublasApi& cublas_api() {
  static CublasApi api;
  return api;
}
inline int cublas_ensure() {
  CublasApi& api = cublas_api();
  if (!api.lib) {
    const char* names[] = {"libcublas.so.12", "libcublas.so", "/usr/local/cuda/lib64/libcublas.so.12"};
    for (const char* nm : names) {
      api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) return fail(KS_ERR_LAUNCH, "dense Newton step: cannot load libcublas.so.12 (needed for the batched fp64 GEMMs)");
    api.create = reinterpret_cast<decltype(api.create)>(dlsym(api.lib, "cublasCreate_v2"));
    api.set_stream = reinterpret_cast<decltype(api.set_stream)>(dlsym(api.lib, "cublasSetStream_v2"));
    api.set_workspace = reinterpret_cast<decltype(api.set_workspace)>(dlsym(api.lib, "cublasSetWorkspace_v2"));
    api.dgemm_sb = reinterpret_cast<decltype(api.dgemm_sb)>(dlsym(api.lib, "cublasDgemmStridedBatched"));
    if (!api.create || !api.set_stream || !api.dgemm_sb) return fail(KS_ERR_LAUNCH, "dense Newton step: cuBLAS symbols missing");
  }
  const int dev = current_device();
  if (!api.handle[dev] && api.create(&api.handle[dev]) != CUBLAS_STATUS_SUCCESS)
    return fail(KS_ERR_LAUNCH, "dense Newton step: cublasCreate failed");
  return 0;
}
inline int dgemm_batched(cudaStream_t s, bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda, long long sa,
                         const double* B, int ldb, long long sb, double beta, double* C, int ldc, long long sc, int batch) {
  if (m <= 0 || n <= 0 || batch <= 0) return 0;
  if (int rc = cublas_ensure()) return rc;
  CublasApi& api = cublas_api();
  cublasHandle_t h = api.handle[current_device()];
  if (api.set_stream(h, s) != CUBLAS_STATUS_SUCCESS) return fail(KS_ERR_LAUNCH, "cublasSetStream failed");
  const cublasStatus_t st = api.dgemm_sb(h, ta ? CUBLAS_OP_T : CUBLAS_OP_N, tb ? CUBLAS_OP_T : CUBLAS_OP_N, m, n, k, &alpha, A, lda, sa, B, ldb,
                                         sb, &beta, C, ldc, sc, batch);
  return st == CUBLAS_STATUS_SUCCESS ? 0 : fail(KS_ERR_LAUNCH, "cublasDgemmStridedBatched failed");
}
#endif

}  // namespace ksdh
