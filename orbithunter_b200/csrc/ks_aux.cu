// extern "C" entry points of the auxiliary batched operations (ks_aux.cuh).
#include "ks_aux.cuh"

#include <cmath>

#include "ks_host.cuh"

extern "C" int ks_resize_modes(int cls, int n, int m, int n_new, int m_new, int batch, const double* in, double* out, void* stream) {
  using namespace ksh;
  auto bad = [](int v) { return v < 4 || (v % 2) || v > 8192; };
  if (cls < 0 || cls > 3 || bad(n) || bad(m) || bad(n_new) || bad(m_new)) return fail(KS_ERR_SHAPE, "resize: need even n, m >= 4");
  if (batch < 0) return fail(KS_ERR_ARG, "negative batch");
  if (batch == 0) return 0;
  if (!in || !out || in == out) return fail(KS_ERR_ARG, "resize: null or aliased array argument");
  ksaux::ResizeArgs a{in, out, cls, n, m, n_new, m_new, mode_cols(cls, m), mode_cols(cls, m_new), batch,
                      std::sqrt((double)n_new / (double)n), std::sqrt((double)m_new / (double)m)};
  KS_LAUNCH(ksaux::ks_resize_kernel, batch, ksaux::kThreads, 0, static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "resize kernel launch failed");
}

extern "C" int ks_rescale_field(int n, int m, int batch, double* field, double magnitude, int method, double* norms_out,
                                void* stream) {
  using namespace ksh;
  if (n < 1 || m < 1) return fail(KS_ERR_SHAPE, "rescale: empty field");
  if (method < 0 || method > 3) return fail(KS_ERR_ARG, "rescale: method must be 0 (inf), 1 (L1), 2 (L2) or 3 (LP)");
  if (batch < 0) return fail(KS_ERR_ARG, "negative batch");
  if (batch == 0) return 0;
  if (!field) return fail(KS_ERR_ARG, "rescale: null field");
  ksaux::RescaleArgs a{field, norms_out, n, m, batch, method, magnitude};
  KS_LAUNCH(ksaux::ks_rescale_kernel, batch, ksaux::kThreads, 0, static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "rescale kernel launch failed");
}

extern "C" int ks_preprocess(int cls, int n, int m, int batch, const double* modes, const double* params, int32_t* code_out,
                             double* norms_out, void* stream) {
  using namespace ksh;
  if (cls < 0 || cls > 3 || n < 4 || m < 4 || (n % 2) || (m % 2)) return fail(KS_ERR_SHAPE, "preprocess: need even n, m >= 4");
  if (batch < 0) return fail(KS_ERR_ARG, "negative batch");
  if (batch == 0) return 0;
  if (!modes || !params || !code_out) return fail(KS_ERR_ARG, "preprocess: null array argument");
  ksaux::PreprocessArgs a{modes, params, code_out, norms_out, cls, n, m, mode_cols(cls, m), batch};
  KS_LAUNCH(ksaux::ks_preprocess_kernel, batch, ksaux::kThreads, 0, static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "preprocess kernel launch failed");
}
