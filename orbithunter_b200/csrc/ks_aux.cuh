// Small batched device operations either side of the solver loops: rediscretisation of the modes (Orbit.resize),
// field rescaling (Orbit.rescale) and the equilibrium / zero-solution norms of OrbitKS.preprocess.
// Reference: orbithunter/core.py:1167-1226, :1831-1853; orbithunter/ks/orbits.py:1428-1596, :2969-3008, :3325-3394, :3735-3804.
// All three are HBM-bound elementwise / reduction kernels: one CTA per orbit, coalesced 8-byte accesses.
#pragma once
#include "../../include/ks_b200.h"
#include "ks_common.cuh"

namespace ksaux {
constexpr int kThreads = 256;

// sum or max of one value per thread over the CTA, result in every thread (fixed order: deterministic)
template <bool MAX>
KS_DEVI double cta_reduce(double v, double* red) {
  const int t = threadIdx.x;
  red[t] = v;
  __syncthreads();
  for (int s = kThreads / 2; s > 0; s >>= 1) {
    if (t < s) red[t] = MAX ? fmax(red[t], red[t + s]) : red[t] + red[t + s];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

struct ResizeArgs {
  const double* in;
  double* out;
  int cls, n, m, n2, m2, cols, cols2, batch;
  double st, sx;  // sqrt(n2 / n), sqrt(m2 / m)
};
// out[r2][c2] = (in[r][c] * st) * sx where (r, c) carries the same (time index, Re/Im_t, wavenumber, Re/Im_x), else 0
__global__ void __launch_bounds__(kThreads) ks_resize_kernel(const ResizeArgs a) {
  const int b = blockIdx.x;
  const int h = a.n / 2 - 1, k = a.m / 2 - 1, h2 = a.n2 / 2 - 1, k2 = a.m2 / 2 - 1;
  const bool full = (a.cls == KS_ORBIT || a.cls == KS_RELATIVE);
  const double* src = a.in + (size_t)b * (a.n - 1) * a.cols;
  double* dst = a.out + (size_t)b * (a.n2 - 1) * a.cols2;
  const int total = (a.n2 - 1) * a.cols2;
  for (int i = threadIdx.x; i < total; i += kThreads) {
    const int r2 = i / a.cols2, c2 = i - r2 * a.cols2;
    const int imt = r2 > h2 ? 1 : 0, j = imt ? r2 - h2 : r2;                 // row -> (Re/Im of the time coefficient, index j)
    const int imx = (full && c2 >= k2) ? 1 : 0, kap = (imx ? c2 - k2 : c2);  // column -> (Re/Im of the space coefficient, kappa - 1)
    double v = 0.0;
    if (j <= h && kap < k) {
      const int r = imt ? h + j : j, c = imx ? k + kap : kap;
      v = src[r * a.cols + c];
      if (a.n2 != a.n) v = __dmul_rn(v, a.st);
      if (a.m2 != a.m) v = __dmul_rn(v, a.sx);
    }
    dst[i] = v;
  }
}

struct RescaleArgs {
  double* field;
  double* norms;
  int n, m, batch, method;
  double magnitude;
};
__global__ void __launch_bounds__(kThreads) ks_rescale_kernel(const RescaleArgs a) {
  __shared__ double red[kThreads];
  double* u = a.field + (size_t)blockIdx.x * a.n * a.m;
  const int total = a.n * a.m, t = threadIdx.x;
  if (a.method == 3) {  // sign(u) |u|^p
    for (int i = t; i < total; i += kThreads) {
      const double v = u[i];
      u[i] = v == 0.0 ? 0.0 : copysign(pow(fabs(v), a.magnitude), v);
    }
    if (a.norms && t == 0) a.norms[blockIdx.x] = 0.0;
    return;
  }
  double nrm;
  if (a.method == 0) {
    double mx = 0.0;
    for (int i = t; i < total; i += kThreads) mx = fmax(mx, fabs(u[i]));
    nrm = cta_reduce<true>(mx, red);
  } else if (a.method == 1) {  // induced 1-norm of the n x m array: largest column sum
    double mx = 0.0;
    for (int c = t; c < a.m; c += kThreads) {
      double s = 0.0;
      for (int r = 0; r < a.n; ++r) s += fabs(u[r * a.m + c]);
      mx = fmax(mx, s);
    }
    nrm = cta_reduce<true>(mx, red);
  } else {
    double s = 0.0;
    for (int i = t; i < total; i += kThreads) s = fma(u[i], u[i], s);
    nrm = sqrt(cta_reduce<false>(s, red));
  }
  for (int i = t; i < total; i += kThreads) u[i] = a.magnitude * u[i] / nrm;  // same operation order as the reference
  if (a.norms && t == 0) a.norms[blockIdx.x] = nrm;
}

struct PreprocessArgs {
  const double* modes;
  const double* params;
  int32_t* code;
  double* norms;
  int cls, n, m, cols, batch;
};
__global__ void __launch_bounds__(kThreads) ks_preprocess_kernel(const PreprocessArgs a) {
  __shared__ double red[kThreads];
  const int b = blockIdx.x, t = threadIdx.x, h = a.n / 2 - 1;
  const double* M = a.modes + (size_t)b * (a.n - 1) * a.cols;
  const double T = a.params[b * 3];
  const int total = (a.n - 1) * a.cols;
  double su = 0.0, st = 0.0;
  for (int i = t; i < total; i += kThreads) {
    const int r = i / a.cols, j = r > h ? r - h : r;
    const double v = M[i], w = (T != 0.0) ? ks::kTwoPi * j / T : 0.0;
    su = fma(v, v, su);
    st = fma(w * v, w * v, st);
  }
  const double nu = sqrt(cta_reduce<false>(su, red)), nt = sqrt(cta_reduce<false>(st, red));
  if (t == 0) {
    // thresholds: OrbitKS family field.size * 1e-9 (ks/orbits.py:1582, :1589), RelativeOrbitKS 1e-5 (:2987, :3000)
    const double thr = a.cls == KS_RELATIVE ? 1e-5 : (double)a.n * a.m * 1e-9;
    int code = 0;
    if (a.cls == KS_RELATIVE) code = (nu < thr || T == 0.0) ? 1 : (nt < thr ? 2 : 0);
    else code = T == 0.0 ? 2 : (nu < thr ? 1 : (nt < thr ? 2 : 0));
    a.code[b] = code;
    if (a.norms) {
      a.norms[b * 2] = nu;
      a.norms[b * 2 + 1] = nt;
    }
  }
}
}  // namespace ksaux
