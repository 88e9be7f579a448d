// Host-side helpers shared by the translation units of libks_b200.so (error text, SM count, geometry).
#pragma once
#include <cmath>
#include <cstdio>

#include "../../include/ks_b200.h"
#include "ks_common.cuh"
#include "ks_pointwise.cuh"

namespace ksh {
inline thread_local char g_err[256] = "";
inline int fail(int code, const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
// Per-device caches.  cudaFuncSetAttribute and the SM count belong to the CURRENT device, and one process may drive
// several GPUs (BatchedOrbitKS switches devices per call), so both are keyed by cudaGetDevice().
constexpr int kMaxDevices = 64;
inline int current_device() {
#ifdef KS_HOST_EMUL
  return 0;
#else
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
#endif
}
inline int sm_count() {
#ifdef KS_HOST_EMUL
  return ks_emul::device_sm_count;
#else
  static int cached[kMaxDevices] = {};
  const int dev = current_device();
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
#endif
}
// Opt a kernel into `bytes` of dynamic shared memory once per device.  One DynSmemOptIn object per kernel
// instantiation (a function-local static at the launch site).
struct DynSmemOptIn {
  bool done[kMaxDevices] = {};
  template <class K>
  bool ensure(K kernel, int bytes) {
    const int dev = current_device();
    if (done[dev]) return true;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
    done[dev] = true;
    return true;
  }
};
inline int mode_cols(int cls, int m) { return (cls == KS_ORBIT || cls == KS_RELATIVE) ? m - 2 : m / 2 - 1; }
inline int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}
inline kspw::Geom make_geom(int cls, int n, int m) {
  kspw::Geom g;
  g.n = n;
  g.m = m;
  g.h = n / 2 - 1;
  g.k = m / 2 - 1;
  g.cols = mode_cols(cls, m);
  g.nm = (n - 1) * g.cols;
  g.nmax = n > m ? n : m;
  g.log_n = ilog2_exact(n);
  g.log_m = ilog2_exact(m);
  g.ct = 1.0 / std::sqrt(2.0 * n);
  g.cx = 1.0 / std::sqrt(2.0 * m);
  return g;
}
}  // namespace ksh

// tile path entry points, one per symmetry class (ks_tile_c{0..3}.cu); `a.scratch` is the workspace, `chunk` the number of
// orbits one pass sequence covers
int ks_tile_run_c0(int op, int n, int m, int chunk, const KsEvalArgs& a, cudaStream_t s);
int ks_tile_run_c1(int op, int n, int m, int chunk, const KsEvalArgs& a, cudaStream_t s);
int ks_tile_run_c2(int op, int n, int m, int chunk, const KsEvalArgs& a, cudaStream_t s);
int ks_tile_run_c3(int op, int n, int m, int chunk, const KsEvalArgs& a, cudaStream_t s);
// ks_api.cu: 0 = plain pass kernels, 1 = TMA-staged rows and columns (ks_tile_tma.cuh), 2 = rows only, 3 = columns only
int ks_tile_tma_mode();
// 32x32 register-resident kernels, one translation unit per symmetry class (ks_w32_c{0..3}.cu); `variant` selects the
// kernel build for eqn / costgrad (3, 4, 5: see ks_debug_warp32_variant)
int ks_w32_run_c0(int op, int variant, const KsEvalArgs& a, cudaStream_t s);
int ks_w32_run_c1(int op, int variant, const KsEvalArgs& a, cudaStream_t s);
int ks_w32_run_c2(int op, int variant, const KsEvalArgs& a, cudaStream_t s);
int ks_w32_run_c3(int op, int variant, const KsEvalArgs& a, cudaStream_t s);
