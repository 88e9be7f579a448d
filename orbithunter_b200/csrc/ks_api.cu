// extern "C" entry points of libks_b200.so (declared in include/ks_b200.h) and kernel dispatch.
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "../../include/ks_b200.h"
#include "ks_adjoint.cuh"
#include "ks_common.cuh"
#include "ks_generic.cuh"
#include "ks_warp32.cuh"

namespace {
thread_local char g_err[256] = "";
int fail(int code, const char* msg) {
  std::snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int sm_count() {
#ifdef KS_HOST_EMUL
  return ks_emul::device_sm_count;
#else
  static int cached = 0;
  if (!cached) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached = n;
  }
  return cached;
#endif
}
bool shape_ok(int cls, int n, int m) {
  return cls >= 0 && cls <= 3 && n >= 4 && m >= 4 && (n % 2 == 0) && (m % 2 == 0) && n <= 4096 && m <= 4096;
}
int mode_cols(int cls, int m) { return (cls == KS_ORBIT || cls == KS_RELATIVE) ? m - 2 : m / 2 - 1; }
int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}
bool g_force_generic = false;  // debug/test switch: route 32x32 through the generic kernel as a cross-check
bool g_warp32_smem_stash = false;  // which warp32 build variant to launch (see ks_warp32.cuh Cfg)
bool use_warp32(int n, int m) { return n == 32 && m == 32 && !g_force_generic; }
int warp32_grid(int batch, int warps_per_cta) {
  const int npairs = (batch + 1) / 2;
  const int ctas = (npairs + warps_per_cta - 1) / warps_per_cta;
  const int sms = sm_count();
  return ctas < sms ? ctas : sms;
}
int generic_grid(int batch) {
  const int cap = 2 * sm_count();
  return batch < cap ? batch : cap;
}
ksg::Geom make_geom(int cls, int n, int m) {
  ksg::Geom g;
  g.n = n;
  g.m = m;
  g.h = n / 2 - 1;
  g.k = m / 2 - 1;
  g.cols = mode_cols(cls, m);
  g.nm = (n - 1) * g.cols;
  g.nmax = n > m ? n : m;
  g.log_n = ilog2_exact(n);
  g.log_m = ilog2_exact(m);
  g.ct = 1.0 / sqrt(2.0 * n);
  g.cx = 1.0 / sqrt(2.0 * m);
  return g;
}

template <int CLS, int OP, bool TS>
int launch_warp32_v(const KsEvalArgs& a, cudaStream_t s) {
  using Cfg = ksw32::Cfg<TS>;
  auto kern = ksw32::ks_warp32_kernel<CLS, OP, TS>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes) != cudaSuccess)
      return fail(KS_ERR_LAUNCH, "cudaFuncSetAttribute(warp32) failed");
    attr_done = true;
  }
  KS_LAUNCH(kern, warp32_grid(a.batch, Cfg::warps_per_cta), Cfg::warps_per_cta * 32, Cfg::smem_bytes, s, a);
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "warp32 kernel launch failed");
}
template <int CLS, int OP>
int launch_warp32(const KsEvalArgs& a, cudaStream_t s) {
  return g_warp32_smem_stash ? launch_warp32_v<CLS, OP, true>(a, s) : launch_warp32_v<CLS, OP, false>(a, s);
}
template <int CLS, int OP>
int launch_generic(const ksg::GenArgs& ga, cudaStream_t s) {
  auto kern = ksg::ks_generic_kernel<CLS, OP>;
  const size_t smem = ksg::smem_bytes(ga.g.n, ga.g.m);
  if (smem > 227 * 1024) return fail(KS_ERR_SHAPE, "grid too large for the generic kernel's shared-memory line buffers");
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return fail(KS_ERR_LAUNCH, "cudaFuncSetAttribute(generic) failed");
  KS_LAUNCH(kern, generic_grid(ga.e.batch), ksg::kThreads, smem, s, ga);
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "generic kernel launch failed");
}

template <int OP>
int dispatch(int cls, int n, int m, const KsEvalArgs& a, void* ws, size_t ws_bytes, void* stream, int from_basis = 0,
             int to_basis = 0, const double* tin = nullptr, double* tout = nullptr) {
  if (!shape_ok(cls, n, m)) return fail(KS_ERR_SHAPE, "unsupported class / discretization (need even n, m >= 4)");
  if (a.batch < 0) return fail(KS_ERR_ARG, "negative batch");
  if (a.batch == 0) return 0;
  const size_t need = ks_workspace_bytes(cls, n, m, a.batch);
  if (need > 0 && (ws == nullptr || ws_bytes < need)) return fail(KS_ERR_WORKSPACE, "workspace too small (see ks_workspace_bytes)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  KsEvalArgs e = a;
  e.scratch = static_cast<double*>(ws);
  if constexpr (OP != ksg::KS_OP_TRANSFORM) {
    if (use_warp32(n, m)) {
      switch (cls) {
        case KS_ORBIT: return launch_warp32<KS_ORBIT, OP>(e, s);
        case KS_RELATIVE: return launch_warp32<KS_RELATIVE, OP>(e, s);
        case KS_SHIFT_REFLECTION: return launch_warp32<KS_SHIFT_REFLECTION, OP>(e, s);
        default: return launch_warp32<KS_ANTISYMMETRIC, OP>(e, s);
      }
    }
  }
  ksg::GenArgs ga;
  ga.e = e;
  ga.g = make_geom(cls, n, m);
  ga.from_basis = from_basis;
  ga.to_basis = to_basis;
  ga.tin = tin;
  ga.tout = tout;
  ga.slot_doubles = ksg::slot_doubles(n, m, ga.g.cols);
  switch (cls) {
    case KS_ORBIT: return launch_generic<KS_ORBIT, OP>(ga, s);
    case KS_RELATIVE: return launch_generic<KS_RELATIVE, OP>(ga, s);
    case KS_SHIFT_REFLECTION: return launch_generic<KS_SHIFT_REFLECTION, OP>(ga, s);
    default: return launch_generic<KS_ANTISYMMETRIC, OP>(ga, s);
  }
}
}  // namespace

extern "C" {

int ks_b200_abi_version(void) { return 1; }
const char* ks_last_error(void) { return g_err; }

void ks_debug_force_generic(int on) { g_force_generic = on != 0; }
void ks_debug_warp32_variant(int smem_stash) { g_warp32_smem_stash = smem_stash != 0; }

int ks_mode_size(int cls, int n, int m) { return shape_ok(cls, n, m) ? (n - 1) * mode_cols(cls, m) : 0; }

size_t ks_workspace_bytes(int cls, int n, int m, int batch) {
  if (!shape_ok(cls, n, m) || batch <= 0) return 0;
  const int cols = mode_cols(cls, m);
  // generic path slots (always available: ks_transform uses it for every shape)
  size_t bytes = (size_t)generic_grid(batch) * ksg::slot_doubles(n, m, cols) * 8;
  if (use_warp32(n, m)) {
    const size_t w = (size_t)warp32_grid(batch, 4) * ksw32::kMaxWarpsPerCta * 2 * 31 * cols * 8;
    if (w > bytes) bytes = w;
  }
  return bytes;
}

int ks_transform(int cls, int n, int m, int batch, int from_basis, int to_basis, const double* in, double* out, void* ws,
                 size_t ws_bytes, void* stream) {
  if (from_basis < 0 || from_basis > 2 || to_basis < 0 || to_basis > 2) return fail(KS_ERR_ARG, "unknown basis id");
  if (batch > 0 && (!in || !out)) return fail(KS_ERR_ARG, "null array argument");
  KsEvalArgs a;
  std::memset(&a, 0, sizeof(a));
  a.batch = batch;
  return dispatch<ksg::KS_OP_TRANSFORM>(cls, n, m, a, ws, ws_bytes, stream, from_basis, to_basis, in, out);
}

int ks_eqn(int cls, int n, int m, int batch, const double* modes, const double* params, double* f_out, double* cost_out,
           void* ws, size_t ws_bytes, void* stream) {
  if (batch > 0 && (!modes || !params || !f_out)) return fail(KS_ERR_ARG, "null array argument");
  KsEvalArgs a;
  std::memset(&a, 0, sizeof(a));
  a.modes = modes;
  a.params = params;
  a.out_state = f_out;
  a.out_cost = cost_out;
  a.batch = batch;
  return dispatch<KS_OP_EQN>(cls, n, m, a, ws, ws_bytes, stream);
}

int ks_matvec(int cls, int n, int m, int batch, const double* modes, const double* params, const double* vmodes,
              const double* vparams, uint32_t cmask, double* out_state, void* ws, size_t ws_bytes, void* stream) {
  if (batch > 0 && (!modes || !params || !vmodes || !vparams || !out_state)) return fail(KS_ERR_ARG, "null array argument");
  KsEvalArgs a;
  std::memset(&a, 0, sizeof(a));
  a.modes = modes;
  a.params = params;
  a.vmodes = vmodes;
  a.vparams = vparams;
  a.out_state = out_state;
  a.batch = batch;
  a.cmask = cmask;
  return dispatch<KS_OP_MATVEC>(cls, n, m, a, ws, ws_bytes, stream);
}

int ks_rmatvec(int cls, int n, int m, int batch, const double* modes, const double* params, const double* vmodes,
               uint32_t cmask, double* out_state, double* out_params, void* ws, size_t ws_bytes, void* stream) {
  if (batch > 0 && (!modes || !params || !vmodes || !out_state)) return fail(KS_ERR_ARG, "null array argument");
  KsEvalArgs a;
  std::memset(&a, 0, sizeof(a));
  a.modes = modes;
  a.params = params;
  a.vmodes = vmodes;
  a.out_state = out_state;
  a.out_params = out_params;
  a.batch = batch;
  a.cmask = cmask;
  return dispatch<KS_OP_RMATVEC>(cls, n, m, a, ws, ws_bytes, stream);
}

int ks_costgrad(int cls, int n, int m, int batch, const double* modes, const double* params, uint32_t cmask,
                int preconditioning, double* cost_out, double* grad_state, double* grad_params, double* f_out, void* ws,
                size_t ws_bytes, void* stream) {
  if (batch > 0 && (!modes || !params || !grad_state)) return fail(KS_ERR_ARG, "null array argument");
  KsEvalArgs a;
  std::memset(&a, 0, sizeof(a));
  a.modes = modes;
  a.params = params;
  a.out_state = grad_state;
  a.out_params = grad_params;
  a.out_cost = cost_out;
  a.f_out = f_out;
  a.batch = batch;
  a.cmask = cmask;
  a.precond = preconditioning;
  return dispatch<KS_OP_COSTGRAD>(cls, n, m, a, ws, ws_bytes, stream);
}

int ks_precondition(int cls, int n, int m, int batch, const double* state_in, const double* gparams_in,
                    const double* params, uint32_t cmask, double pexp_t, double pexp_x, double* state_out,
                    double* gparams_out, void* stream) {
  if (!shape_ok(cls, n, m)) return fail(KS_ERR_SHAPE, "unsupported class / discretization (need even n, m >= 4)");
  if (batch < 0) return fail(KS_ERR_ARG, "negative batch");
  if (batch == 0) return 0;
  if (!state_in || !params || !state_out || (gparams_out && !gparams_in)) return fail(KS_ERR_ARG, "null array argument");
  ksg::PrecArgs a;
  a.in = state_in;
  a.gp_in = gparams_in;
  a.params = params;
  a.out = state_out;
  a.gp_out = gparams_out;
  a.batch = batch;
  a.n = n;
  a.m = m;
  a.cols = mode_cols(cls, m);
  a.cmask = cmask;
  a.pexp_t = pexp_t;
  a.pexp_x = pexp_x;
  const int nm = (n - 1) * a.cols;
  int bx = (nm + 255) / 256;
  if (bx > 64) bx = 64;
  for (int b0 = 0; b0 < batch; b0 += 65535) {  // gridDim.y limit
    ksg::PrecArgs c = a;
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    c.in += (size_t)b0 * nm;
    c.out += (size_t)b0 * nm;
    c.params += (size_t)b0 * 3;
    if (c.gp_in) c.gp_in += (size_t)b0 * 3;
    if (c.gp_out) c.gp_out += (size_t)b0 * 3;
    KS_LAUNCH(ksg::ks_precondition_kernel, dim3(bx, nb), dim3(256), 0, static_cast<cudaStream_t>(stream), c);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "precondition kernel launch failed");
}

// ---- adjoint descent driver (reference optimize.py:367-509) ------------------------------------------------------
namespace {
__global__ void adj_init_kernel(int* status, int* nit, int* live, double* step, double step0, int batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) {
    status[i] = -1;
    nit[i] = 0;
    live[i] = 1;
    step[i] = step0;
  }
}
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace

size_t ks_adjoint_descent_workspace_bytes(int cls, int n, int m, int batch) {
  if (!shape_ok(cls, n, m) || batch <= 0) return 0;
  const size_t nm = (size_t)(n - 1) * mode_cols(cls, m);
  size_t bytes = align256(ks_workspace_bytes(cls, n, m, batch));
  bytes += 3 * align256((size_t)batch * nm * 8);   // g, trial u, trial gradient
  bytes += 3 * align256((size_t)batch * 3 * 8);    // gp, trial params, trial gp
  bytes += align256((size_t)batch * 8);            // trial cost
  bytes += align256((size_t)batch * 4) + 256;      // live flags, live counter
  return bytes;
}

int ks_adjoint_descent(int cls, int n, int m, int batch, double* modes, double* params, uint32_t cmask, double tol,
                       double ftol, int maxiter, double min_step, int preconditioning, double step_size,
                       int32_t* status_out, int32_t* nit_out, double* cost_out, double* step_out, int check_every,
                       int* passes_out, void* ws, size_t ws_bytes, void* stream) {
  if (!shape_ok(cls, n, m)) return fail(KS_ERR_SHAPE, "unsupported class / discretization (need even n, m >= 4)");
  if (batch < 0 || maxiter < 0) return fail(KS_ERR_ARG, "negative batch or maxiter");
  if (passes_out) *passes_out = 0;
  if (batch == 0) return 0;
  if (!modes || !params || !status_out || !nit_out || !cost_out || !step_out) return fail(KS_ERR_ARG, "null array argument");
  if (!ws || ws_bytes < ks_adjoint_descent_workspace_bytes(cls, n, m, batch))
    return fail(KS_ERR_WORKSPACE, "workspace too small (see ks_adjoint_descent_workspace_bytes)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t nm = (size_t)(n - 1) * mode_cols(cls, m);
  unsigned char* p = static_cast<unsigned char*>(ws);
  const size_t eval_bytes = ks_workspace_bytes(cls, n, m, batch);
  void* eval_ws = p;
  p += align256(eval_bytes);
  auto take = [&](size_t bytes) {
    void* q = p;
    p += align256(bytes);
    return q;
  };
  ksadj::AdjState st;
  st.u = modes;
  st.up = params;
  st.g = static_cast<double*>(take(batch * nm * 8));
  st.tu = static_cast<double*>(take(batch * nm * 8));
  st.tg = static_cast<double*>(take(batch * nm * 8));
  st.gp = static_cast<double*>(take((size_t)batch * 24));
  st.tup = static_cast<double*>(take((size_t)batch * 24));
  st.tgp = static_cast<double*>(take((size_t)batch * 24));
  st.tcost = static_cast<double*>(take((size_t)batch * 8));
  st.live = static_cast<int*>(take((size_t)batch * 4));
  st.n_live = static_cast<int*>(take(4));
  st.cost = cost_out;
  st.step = step_out;
  st.status = status_out;
  st.nit = nit_out;
  st.batch = batch;
  st.nm = (int)nm;
  st.tol = tol;
  st.ftol = ftol;
  st.min_step = min_step;
  st.maxiter = maxiter;
  KS_LAUNCH(adj_init_kernel, (batch + 255) / 256, 256, 0, s, status_out, nit_out, st.live, step_out, step_size, batch);
  cudaMemcpyAsync(st.tu, modes, batch * nm * 8, cudaMemcpyDeviceToDevice, s);
  cudaMemcpyAsync(st.tup, params, (size_t)batch * 24, cudaMemcpyDeviceToDevice, s);
  cudaMemsetAsync(st.g, 0, batch * nm * 8, s);
  cudaMemsetAsync(st.gp, 0, (size_t)batch * 24, s);
  if (check_every <= 0) check_every = 64;
  // every live orbit either counts an iteration or halves its step in each pass: steps halve at most ~1100 times
  // before underflow and far fewer before min_step, iterations stop at maxiter
  const long max_passes = (long)maxiter + 1200;
  int passes = 0;
  for (long it = 0; it < max_passes; ++it) {
    KsEvalArgs a;
    std::memset(&a, 0, sizeof(a));
    a.modes = st.tu;
    a.params = st.tup;
    a.out_state = st.tg;
    a.out_params = st.tgp;
    a.out_cost = st.tcost;
    a.batch = batch;
    a.cmask = cmask;
    a.precond = preconditioning;
    const int rc = dispatch<KS_OP_COSTGRAD>(cls, n, m, a, eval_ws, eval_bytes, stream);
    if (rc != 0) return rc;
    cudaMemsetAsync(st.n_live, 0, 4, s);
    st.first = (it == 0) ? 1 : 0;
    KS_LAUNCH(ksadj::adj_commit_trial_kernel, batch, ksadj::kThreads, 0, s, st);
    ++passes;
    if ((it + 1) % check_every == 0 || it + 1 == max_passes) {
      int n_live = 0;
      if (cudaMemcpyAsync(&n_live, st.n_live, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
        return fail(KS_ERR_LAUNCH, "adjoint descent: device error while polling the live-orbit counter");
      if (n_live == 0) break;
    }
  }
  if (passes_out) *passes_out = passes;
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "adjoint descent launch failed");
}

}  // extern "C"
