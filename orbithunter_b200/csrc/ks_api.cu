// extern "C" entry points of libks_b200.so (declared in include/ks_b200.h) and kernel dispatch.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>

#include "../../include/ks_b200.h"
#include "ks_adjoint.cuh"
#include "ks_common.cuh"
#include "ks_dense.cuh"
#include "ks_dense_host.cuh"
#include "ks_generic.cuh"
#include "ks_krylov.cuh"
#include "ks_host.cuh"
#include "ks_tile.cuh"
#include "ks_warp32.cuh"
#include "ks_warp32x1.cuh"

using namespace ksh;
// setup copies / memsets of the solver drivers: surface a failing call where it happens (a bad pointer would otherwise
// leave garbage in the solver state and be reported much later, by a poll)
#define KS_CHECK_CUDA(call, what)                                                      \
  do {                                                                                 \
    if ((call) != cudaSuccess) return fail(KS_ERR_LAUNCH, what ": CUDA call failed");  \
  } while (0)
namespace {
bool shape_ok(int cls, int n, int m) {
  return cls >= 0 && cls <= 3 && n >= 4 && m >= 4 && (n % 2 == 0) && (m % 2 == 0) && n <= 4096 && m <= 4096;
}
int g_w32_pingpong = 0;
int g_w32_stagger_ns = 0;  // tuning: start delay between the warps of one scheduler in the 32x32 kernels
bool g_force_generic = false;  // debug/test switch: route 32x32 through the generic kernel as a cross-check
int g_warp32_variant = 4;  // 32x32 kernels: 4 = one orbit per warp (ks_warp32x1.cuh, default), 3 = two orbits per warp (ks_warp32.cuh,
                           // which also serves matvec / rmatvec for variants 5, 6, 7)
int g_dense_outer = 4;  // dense Newton step: Cholesky blocks per outer panel of the two-level factorisation (1: rank-64 updates throughout)
int g_dense_jac_mode = 2;  // dense Newton step, 32x32: 2 = matrix-free J J^T (no Jacobian), 1 = Jacobian by the x1 kernel's Jacobian mode + GEMM,
                           // 0 = Jacobian by a matvec over a stored identity + GEMM
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
// 32x32 kernels: instantiated per symmetry class in ks_w32_c{0..3}.cu (see ks_w32_host.cuh)
int w32_run(int cls, int op, const KsEvalArgs& a0, cudaStream_t s) {
  KsEvalArgs a = a0;
  a.stagger_ns = g_w32_stagger_ns;
  a.pingpong = g_w32_pingpong;
  switch (cls) {
    case KS_ORBIT: return ks_w32_run_c0(op, g_warp32_variant, a, s);
    case KS_RELATIVE: return ks_w32_run_c1(op, g_warp32_variant, a, s);
    case KS_SHIFT_REFLECTION: return ks_w32_run_c2(op, g_warp32_variant, a, s);
    default: return ks_w32_run_c3(op, g_warp32_variant, a, s);
  }
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


// ---- tile path (n, m in {64, 128, 256}): kernels and launch code live in ks_tile_c{0..3}.cu (one class per translation
// unit, compiled in parallel); see ks_tile_host.cuh ----
int g_adj_inner = 128;  // 32x32 adjoint descent: passes per launch of the multi-pass kernel (0 = one pass per launch)
int g_gmres_cluster = 1;  // 0 = one CTA per orbit GMRES step (cross-check), 1 = auto cluster size, 2/4/8 = at least that size
int g_tile_mode = 1;  // 0 = generic kernel, 1 = tile path with the CTA-resident 64x64 kernel, 2 = tile path, multi-pass only
int g_tile_chunk_mb = 256;
int g_tile_tma = 0;  // multi-pass tiles: 0 = plain pass kernels (default: faster, see ks_tile_tma.cuh), 1 = TMA-staged persistent rows and columns, 2 / 3 = rows / columns only
bool use_tile(int n, int m) { return g_tile_mode != 0 && !g_force_generic && kst::size_ok(n) && kst::size_ok(m); }
int tile_chunk(int n, int m, int cols, int batch) {
  const size_t per = kst::orbit_doubles(n, m, cols) * 8;
  size_t c = ((size_t)g_tile_chunk_mb << 20) / per;
  if (c < 1) c = 1;
  return c < (size_t)batch ? (int)c : batch;
}
size_t tile_workspace_bytes(int n, int m, int cols, int batch) {
  return (size_t)tile_chunk(n, m, cols, batch) * kst::orbit_doubles(n, m, cols) * 8 + 4 * 256;
}
template <int OP>
int launch_tile(int cls, int n, int m, const KsEvalArgs& a, cudaStream_t s) {
  // 64x64: the CTA-resident fused kernel (chunk < 0 selects it); everything else: the multi-pass tiles
  const int chunk = (n == 64 && m == 64 && g_tile_mode == 1) ? -1 : tile_chunk(n, m, mode_cols(cls, m), a.batch);
  switch (cls) {
    case KS_ORBIT: return ks_tile_run_c0(OP, n, m, chunk, a, s);
    case KS_RELATIVE: return ks_tile_run_c1(OP, n, m, chunk, a, s);
    case KS_SHIFT_REFLECTION: return ks_tile_run_c2(OP, n, m, chunk, a, s);
    default: return ks_tile_run_c3(OP, n, m, chunk, a, s);
  }
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
    if (use_tile(n, m)) return launch_tile<OP>(cls, n, m, e, s);
    if (use_warp32(n, m)) {
      return w32_run(cls, OP, e, s);
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

int ks_tile_tma_mode() { return g_tile_tma; }
int g_op_cache = 0;  // Krylov drivers: reuse N(u, u) across the operator applications of one inner solve (64x64); measured: -1 .. +7 %, off

extern "C" {

void ks_debug_operator_cache(int on) { g_op_cache = on != 0; }
void ks_debug_tile_tma(int mode) { g_tile_tma = (mode >= 0 && mode <= 3) ? mode : 0; }
int ks_b200_abi_version(void) { return 1; }
const char* ks_last_error(void) { return g_err; }

void ks_debug_force_generic(int on) { g_force_generic = on != 0; }
void ks_debug_adjoint_inner(int passes) { g_adj_inner = passes < 0 ? 0 : passes; }
void ks_debug_gmres_cluster(int mode) { g_gmres_cluster = (mode == 0 || mode == 1 || mode == 2 || mode == 4 || mode == 8) ? mode : 1; }
void ks_debug_tile(int enabled, int chunk_mb) {
  g_tile_mode = (enabled == 0 || enabled == 2) ? enabled : 1;
  if (chunk_mb > 0) g_tile_chunk_mb = chunk_mb;
}
void ks_debug_warp32_pingpong(int mode) { g_w32_pingpong = mode; }
void ks_debug_warp32_stagger(int ns) { g_w32_stagger_ns = ns < 0 ? 0 : ns; }
void ks_debug_dense_jacobian_mode(int mode) { g_dense_jac_mode = (mode >= 0 && mode <= 2) ? mode : 2; }
void ks_debug_dense_outer_blocks(int blocks) { g_dense_outer = blocks >= 1 && blocks <= 64 ? blocks : 4; }
void ks_debug_warp32_variant(int variant) { g_warp32_variant = (variant >= 3 && variant <= 7) ? variant : 4; }

int ks_mode_size(int cls, int n, int m) { return shape_ok(cls, n, m) ? (n - 1) * mode_cols(cls, m) : 0; }

size_t ks_workspace_bytes(int cls, int n, int m, int batch) {
  if (!shape_ok(cls, n, m) || batch <= 0) return 0;
  const int cols = mode_cols(cls, m);
  // generic path slots (always available: ks_transform uses it for every shape)
  size_t bytes = (size_t)generic_grid(batch) * ksg::slot_doubles(n, m, cols) * 8;
  if (kst::size_ok(n) && kst::size_ok(m)) {
    const size_t w = tile_workspace_bytes(n, m, cols, batch);
    if (w > bytes) bytes = w;
  }
  if (use_warp32(n, m)) {
    const size_t w = (size_t)warp32_grid(batch, 7) * ksw32::kMaxWarpsPerCta * 2 * 31 * cols * 8;
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

static int fill_grid(size_t total) { return (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184); }
static int jacobian_columns(int cls, int n, int m, uint32_t cmask) {
  int nfree = 0;
  for (int i = 0; i < 3; ++i) nfree += ks::free_slot(cmask | (cls == KS_RELATIVE ? 0u : (uint32_t)KS_FIX_S), i) >= 0 ? 1 : 0;
  return (n - 1) * mode_cols(cls, m) + nfree;
}

size_t ks_jacobian_workspace_bytes(int cls, int n, int m, int batch, uint32_t cmask) {
  if (!shape_ok(cls, n, m) || batch <= 0) return 0;
  const size_t N = (size_t)jacobian_columns(cls, n, m, cmask), nm = (size_t)(n - 1) * mode_cols(cls, m);
  if (use_warp32(n, m)) return ks_workspace_bytes(cls, n, m, (int)((size_t)batch * N < 0x7fffffffu ? (size_t)batch * N : 0x7fffffff));
  // one orbit at a time: its copies, the parameter copies, the identity, then the evaluation workspace of N products
  return ((N * nm + N * 3 + N * N) * 8 + 255) / 256 * 256 + ks_workspace_bytes(cls, n, m, (int)N);
}

int ks_jacobian(int cls, int n, int m, int batch, const double* modes, const double* params, uint32_t cmask, double* jac_t,
                void* ws, size_t ws_bytes, void* stream) {
  if (batch > 0 && (!modes || !params || !jac_t)) return fail(KS_ERR_ARG, "null array argument");
  if (!shape_ok(cls, n, m)) return fail(KS_ERR_SHAPE, "unsupported discretization");
  if (batch <= 0) return 0;
  if (cls != KS_RELATIVE) cmask |= KS_FIX_S;
  const int N = jacobian_columns(cls, n, m, cmask), nm = (n - 1) * mode_cols(cls, m);
  if ((size_t)batch * N > 0x7fffffffu) return fail(KS_ERR_ARG, "ks_jacobian: batch * columns exceeds the evaluation index range");
  if (ws_bytes < ks_jacobian_workspace_bytes(cls, n, m, batch, cmask)) return fail(KS_ERR_WORKSPACE, "ks_jacobian: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  KsEvalArgs a;
  std::memset(&a, 0, sizeof(a));
  a.cmask = cmask;
  if (use_warp32(n, m)) {  // unit vectors made in the kernel, u read in place (ks_warp32x1.cuh, Jacobian mode)
    a.modes = modes; a.params = params; a.out_state = jac_t; a.batch = batch * N; a.u_div = N; a.unit_v = 1;
    return dispatch<KS_OP_MATVEC>(cls, n, m, a, ws, ws_bytes, stream);
  }
  double* Urep = static_cast<double*>(ws);
  double* Prep = Urep + (size_t)N * nm;
  double* Vrep = Prep + (size_t)N * 3;
  const size_t head = (((size_t)N * nm + (size_t)N * 3 + (size_t)N * N) * 8 + 255) / 256 * 256;
  KS_LAUNCH(ksd::identity_kernel, fill_grid((size_t)N * N), 256, 0, s, Vrep, N, (size_t)N * N);
  for (int b = 0; b < batch; ++b) {
    KS_LAUNCH(ksd::replicate_kernel, fill_grid((size_t)N * nm), 256, 0, s, modes + (size_t)b * nm, Urep, nm, N, (size_t)N * nm);
    KS_LAUNCH(ksd::replicate_kernel, 8, 256, 0, s, params + (size_t)b * 3, Prep, 3, N, (size_t)N * 3);
    a.modes = Urep; a.params = Prep; a.vmodes = Vrep; a.out_state = jac_t + (size_t)b * N * nm; a.batch = N; a.v_ld = N; a.v_tail = 1;
    const int rc = dispatch<KS_OP_MATVEC>(cls, n, m, a, static_cast<char*>(ws) + head, ws_bytes - head, stream);
    if (rc != 0) return rc;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "ks_jacobian: kernel launch failed");
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

int ks_deriv(int cls, int n, int m, int batch, int axis, int order, int rows, int cols, const double* in,
             const double* params, double* out, void* stream) {
  if (!shape_ok(cls, n, m)) return fail(KS_ERR_SHAPE, "unsupported class / discretization (need even n, m >= 4)");
  if (batch < 0 || order < 0 || (axis != 0 && axis != 1)) return fail(KS_ERR_ARG, "bad batch / order / axis");
  if (batch == 0) return 0;
  if (!in || !params || !out || in == out) return fail(KS_ERR_ARG, "null or aliased array argument");
  const int k = m / 2 - 1;
  ksg::DerivArgs a;
  a.in = in; a.params = params; a.out = out; a.batch = batch; a.rows = rows; a.cols = cols; a.n = n; a.m = m;
  a.axis = axis; a.order = order;
  if (axis == 0) {
    if (rows != n - 1) return fail(KS_ERR_SHAPE, "time derivative needs the modes basis (n-1 rows)");
    a.half = 0;
  } else {
    if (cols == 2 * k) a.half = 0;
    else if (cols == k && (order % 2 == 0)) a.half = 1;
    else return fail(KS_ERR_SHAPE, "spatial derivative: need [Re|Im] columns, or k columns with an even order");
  }
  const int sz = rows * cols;
  int bx = (sz + 255) / 256;
  if (bx > 64) bx = 64;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    ksg::DerivArgs c = a;
    const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
    c.in += (size_t)b0 * sz; c.out += (size_t)b0 * sz; c.params += (size_t)b0 * 3;
    KS_LAUNCH(ksg::ks_deriv_kernel, dim3(bx, nb), dim3(256), 0, static_cast<cudaStream_t>(stream), c);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "derivative kernel launch failed");
}

// ---- adjoint descent driver (reference optimize.py:367-509) ------------------------------------------------------
namespace {
// fused path (32x32, x1 kernel): per-orbit selection of the accepted ping-pong buffer into the caller's arrays
// user layout [rows][cols] <-> packed lane-major layout of ks_warp32x1.cuh (lane = 2 c + h, slot 0 = x1(0),
// slot 2J-1 = x1(J) (row J), slot 2J = x2(J) (row 15+J); half classes: a-halves exist for odd J, b-halves for even J
// (ShiftReflection) or b-halves only (Antisymmetric))
__device__ __host__ inline int packed_user_index(int cls, int lane, int slot) {
  const int c = lane >> 1, h = lane & 1;
  if (c >= 15 || slot > 30) return -1;
  const int J = (slot + 1) >> 1, second = (slot > 0 && (slot & 1) == 0) ? 1 : 0;
  const bool full = (cls == KS_ORBIT || cls == KS_RELATIVE);
  if (!full) {
    const bool pres = (cls == KS_ANTISYMMETRIC) ? (h == 1) : ((J & 1) ? h == 0 : h == 1);
    if (!pres) return -1;
  }
  const int cols = full ? 30 : 15;
  return (second ? 15 + J : J) * cols + (full ? h * 15 : 0) + c;
}
__global__ void adj_pack_kernel(const double* modes, double* packed, int cls, int batch, int nm) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < ksx1::kPackedDoubles; i += blockDim.x) {
    const int pr = i / 64, r = i - pr * 64, lane = r >> 1, slot = 2 * pr + (r & 1);
    const int ui = packed_user_index(cls, lane, slot);
    packed[(size_t)b * ksx1::kPackedDoubles + i] = ui >= 0 ? modes[(size_t)b * nm + ui] : 0.0;
  }
}
__global__ void adj_gather_kernel(const KsAdjFused A, double* modes, double* params, int cls, int batch, int nm) {
  const int b = blockIdx.x;
  if (b >= batch) return;
  const int cur = A.cur[b];
  const double* src = A.u[cur] + (size_t)b * ksx1::kPackedDoubles;
  for (int i = threadIdx.x; i < ksx1::kPackedDoubles; i += blockDim.x) {
    const int pr = i / 64, r = i - pr * 64, lane = r >> 1, slot = 2 * pr + (r & 1);
    const int ui = packed_user_index(cls, lane, slot);
    if (ui >= 0) modes[(size_t)b * nm + ui] = src[i];
  }
  if (threadIdx.x < 3) params[b * 3 + threadIdx.x] = A.p[cur][b * 3 + threadIdx.x];
}
__global__ void adj_fused_init_kernel(int* status, int* nit, int* live, int* cur, double* step, double step0, int batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) {
    status[i] = -1;
    nit[i] = 0;
    live[i] = 1;
    cur[i] = 0;
    step[i] = step0;
  }
}
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
  // fused 32x32 path: ping-pong u / g / params / gradient-params, trial cost, cur flags
  size_t fused = 4 * align256((size_t)batch * ksx1::kPackedDoubles * 8) + 4 * align256((size_t)batch * 24) +
                 align256((size_t)batch * 8) + 2 * align256((size_t)batch * 4) + 256;
  return bytes > fused ? bytes : fused;
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
  auto take = [&](size_t bytes) {
    void* q = p;
    p += align256(bytes);
    return q;
  };
  if (check_every <= 0) check_every = 64;
  if (use_warp32(n, m) && g_warp32_variant >= 4) {
    // ---- fused path: ONE launch per pass (decide on the previous trial + evaluate the next one inside the x1 kernel)
    KsEvalArgs a;
    std::memset(&a, 0, sizeof(a));
    KsAdjFused& A = a.adj;
    const size_t pbytes = (size_t)batch * ksx1::kPackedDoubles * 8;
    for (int i = 0; i < 2; ++i) A.u[i] = static_cast<double*>(take(pbytes));
    for (int i = 0; i < 2; ++i) A.g[i] = static_cast<double*>(take(pbytes));
    for (int i = 0; i < 2; ++i) A.p[i] = static_cast<double*>(take((size_t)batch * 24));
    for (int i = 0; i < 2; ++i) A.gp[i] = static_cast<double*>(take((size_t)batch * 24));
    A.tcost = static_cast<double*>(take((size_t)batch * 8));
    A.live = static_cast<int*>(take((size_t)batch * 4));
    A.cur = static_cast<int*>(take((size_t)batch * 4));
    A.n_live = static_cast<int*>(take(4));
    A.cost = cost_out;
    A.step = step_out;
    A.status = status_out;
    A.nit = nit_out;
    A.tol = tol;
    A.ftol = ftol;
    A.min_step = min_step;
    A.maxiter = maxiter;
    a.batch = batch;
    a.cmask = cmask;
    a.precond = preconditioning;
    KS_LAUNCH(adj_fused_init_kernel, (batch + 255) / 256, 256, 0, s, status_out, nit_out, A.live, A.cur, step_out, step_size, batch);
    KS_LAUNCH(adj_pack_kernel, batch, 128, 0, s, (const double*)modes, A.u[0], cls, batch, (int)nm);
    KS_CHECK_CUDA(cudaMemcpyAsync(A.p[0], params, (size_t)batch * 24, cudaMemcpyDeviceToDevice, s), "solver driver setup");
    KS_CHECK_CUDA(cudaMemsetAsync(A.g[0], 0, pbytes, s), "solver driver setup");
    KS_CHECK_CUDA(cudaMemsetAsync(A.gp[0], 0, (size_t)batch * 24, s), "solver driver setup");
    const long max_passes = (long)maxiter + 1200;
    int passes = 0;
    if (g_adj_inner > 0) {
      // multi-pass kernel: every live orbit runs up to `inner` passes per launch with its state on chip; an orbit either
      // counts an iteration or halves its step in each pass, so ceil(max_passes / inner) launches always suffice
      A.inner = g_adj_inner;
      for (long done = 0; done < max_passes + A.inner; done += A.inner) {
        KS_CHECK_CUDA(cudaMemsetAsync(A.n_live, 0, 4, s), "solver driver setup");
        A.first = (done == 0) ? 1 : 0;
        const int rc = w32_run(cls, KS_OP_ADJMULTI, a, s);
        if (rc != 0) return rc;
        passes += A.inner;
        int n_live = 0;
        if (cudaMemcpyAsync(&n_live, A.n_live, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
          return fail(KS_ERR_LAUNCH, "adjoint descent: device error while polling the live-orbit counter");
        if (n_live == 0) break;
      }
      KS_LAUNCH(adj_gather_kernel, batch, 128, 0, s, A, modes, params, cls, batch, (int)nm);
      if (passes_out) *passes_out = passes;
      return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "fused adjoint descent launch failed");
    }
    for (long it = 0; it < max_passes; ++it) {
      KS_CHECK_CUDA(cudaMemsetAsync(A.n_live, 0, 4, s), "solver driver setup");
      A.first = (it == 0) ? 1 : 0;
      const int rc = w32_run(cls, KS_OP_ADJSTEP, a, s);
      if (rc != 0) return rc;
      ++passes;
      if ((it + 1) % check_every == 0 || it + 1 == max_passes) {
        int n_live = 0;
        if (cudaMemcpyAsync(&n_live, A.n_live, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
          return fail(KS_ERR_LAUNCH, "adjoint descent: device error while polling the live-orbit counter");
        if (n_live == 0) break;
      }
    }
    KS_LAUNCH(adj_gather_kernel, batch, 128, 0, s, A, modes, params, cls, batch, (int)nm);
    if (passes_out) *passes_out = passes;
    return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "fused adjoint descent launch failed");
  }
  const size_t eval_bytes = ks_workspace_bytes(cls, n, m, batch);
  void* eval_ws = p;
  p += align256(eval_bytes);
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
  KS_CHECK_CUDA(cudaMemcpyAsync(st.tu, modes, batch * nm * 8, cudaMemcpyDeviceToDevice, s), "solver driver setup");
  KS_CHECK_CUDA(cudaMemcpyAsync(st.tup, params, (size_t)batch * 24, cudaMemcpyDeviceToDevice, s), "solver driver setup");
  KS_CHECK_CUDA(cudaMemsetAsync(st.g, 0, batch * nm * 8, s), "solver driver setup");
  KS_CHECK_CUDA(cudaMemsetAsync(st.gp, 0, (size_t)batch * 24, s), "solver driver setup");
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
    KS_CHECK_CUDA(cudaMemsetAsync(st.n_live, 0, 4, s), "solver driver setup");
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

// ---- Newton-Krylov driver (reference optimize.py:1196-1404, :1682-1850) -----------------------------------------------
namespace {
__global__ void nk_init_kernel(int* status, int* nit, int* live, int* searching, const double* cost, double tol, int batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch) {
    status[i] = -1;
    nit[i] = 0;
    live[i] = cost[i] > tol ? 1 : 0;  // while cost > tol and status == -1            (:1314)
    searching[i] = 0;
  }
}
__global__ void count_flags_kernel(const int* flags, int batch, int* counter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch && flags[i]) atomicAdd(counter, 1);
}
int n_free(uint32_t cmask) { return 3 - ((cmask & 1) + ((cmask >> 1) & 1) + ((cmask >> 2) & 1)); }
struct NkLayout {
  size_t total;
  size_t off[24];
};
// carve sizes shared by the _workspace_bytes query and the driver
// dense Newton step (method 3): sizes of the per-chunk buffers, shared by the workspace query and the driver
struct DenseLayout {
  size_t total;
  size_t off[16];
};
DenseLayout dense_layout(int cls, int n, int m, int batch, int chunk) {
  const size_t nm = (size_t)(n - 1) * mode_cols(cls, m), N = nm + 3, C = (size_t)chunk;
  const size_t nblk = (nm + ksd::kNb - 1) / ksd::kNb;
  DenseLayout L;
  size_t p = 0;
  int k = 0;
  auto take = [&](size_t bytes) {
    L.off[k++] = p;
    p += align256(bytes);
  };
  take((size_t)batch * 4);                                 // 0 live list (device copy)
  take(C * nm * 8);                                        // 1 U chunk
  take(C * 3 * 8);                                         // 2 P chunk
  take(C * nm * 8);                                        // 3 rhs chunk (-F)
  take(C * nm * 8);                                        // 4 residual / right-hand side of the current solve
  take(C * nm * 8);                                        // 5 y
  take(C * N * 8);                                         // 6 dx chunk
  take(C * N * 3 * 8);                                     // 7 parameters replicated per Jacobian column
  take(C * nm * N * 8);                                    // 8 J (column-major Nm x N per orbit)
  take(C * N * N * 8);                                     // 9 G = J J^T (also: the identity the Jacobian is built from)
  take(C * N * N * 8);                                     // 10 L (also: the orbit replicated per Jacobian column)
  take(C * nblk * ksd::kNb * ksd::kNb * 8);                // 11 inverted diagonal blocks
  take(C * 8);                                             // 12 lambda
  take(C * 8);                                             // 13 mean diagonal of G
  take(C * 4 + 256);                                       // 14 fail flags + redo counter
  take(ks_workspace_bytes(cls, n, m, (int)(C * N)));       // 15 evaluation workspace of the Jacobian build
  L.total = p;
  return L;
}
NkLayout nk_layout(int cls, int n, int m, int batch, int method, int restart, uint32_t cmask) {
  const size_t nm = (size_t)(n - 1) * mode_cols(cls, m);
  const size_t N = nm + 3;  // upper bound on cdof (3 free parameters)
  (void)cmask;
  NkLayout L;
  size_t p = 0;
  int k = 0;
  auto take = [&](size_t bytes) {
    L.off[k++] = p;
    p += align256(bytes);
  };
  take(ks_workspace_bytes(cls, n, m, batch));                 // 0 eval workspace
  take(batch * nm * 8);                                       // 1 F / y
  take(batch * N * 8);                                        // 2 rhs b (cdof or Nm)
  take(batch * N * 8);                                        // 3 vin
  take(batch * N * 8);                                        // 4 w
  take(batch * N * 8);                                        // 5 x
  take(batch * nm * 8);                                       // 6 trial u
  take((size_t)batch * 24);                                   // 7 trial params
  take((size_t)batch * 8);                                    // 8 trial cost
  take((size_t)batch * 8);                                    // 9 step
  take((size_t)batch * 4);                                    // 10 live
  take((size_t)batch * 4);                                    // 11 searching
  take(256);                                                  // 12 counters
  take((size_t)batch * 32 * 8);                               // 13 solver scalars (GMRES SC_COUNT = 8, LSQR / LSMR LQ_COUNT = 32)
  take((size_t)batch * 8 * 4);                                // 14 solver ints
  if (method == 0) {
    take((size_t)batch * (restart + 1) * N * 8);              // 15 V
    take((size_t)batch * restart * (restart + 1) * 8);        // 16 H
    take((size_t)batch * restart * 2 * 8);                    // 17 Givens
    take((size_t)batch * (restart + 1) * 8);                  // 18 S
  } else if (method == 3) {
    take(dense_layout(cls, n, m, batch, restart).total);      // 15 dense region (`restart` carries the chunk size)
    take(8); take(8); take(8);                                // 16-18 unused
  } else {
    take(batch * nm * 8);                                     // 15 u
    take(batch * N * 8);                                      // 16 v
    take(batch * N * 8);                                      // 17 w vector
    take(batch * N * 8);                                      // 18 hbar (LSMR)
  }
  // 19: operator cache (N(u, u)) where the CTA-resident 64x64 kernel runs the operator
  const bool opc = g_op_cache && method != 3 && n == 64 && m == 64 && g_tile_mode == 1 && use_tile(n, m);
  take(opc ? batch * nm * 8 : 8);
  L.total = p;
  return L;
}
}  // namespace

size_t ks_newton_krylov_workspace_bytes(int cls, int n, int m, int batch, int method, int restart) {
  if (!shape_ok(cls, n, m) || batch <= 0 || method < 0 || method > 2 || (method == 0 && restart < 1)) return 0;
  return nk_layout(cls, n, m, batch, method, restart, 0).total;
}

}  // extern "C"
namespace {
// Outer Newton loop shared by ks_newton_krylov (methods 0-2) and ks_newton_dense (method 3, `restart` = chunk size,
// `inner_maxiter` = refinement steps).
int newton_driver(int cls, int n, int m, int batch, double* modes, double* params, uint32_t cmask, int method,
                  double tol, double ftol, int maxiter, double min_step, int preconditioning, int restart,
                  int inner_maxiter, double rtol, double atol, double btol, double conlim, int32_t* status_out,
                  int32_t* nit_out, double* cost_out, long long* counters_out, int check_every, void* ws,
                  size_t ws_bytes, void* stream) {
  if (!shape_ok(cls, n, m)) return fail(KS_ERR_SHAPE, "unsupported class / discretization (need even n, m >= 4)");
  if (batch < 0 || maxiter < 0 || method < 0 || method > 3) return fail(KS_ERR_ARG, "bad batch / maxiter / method");
  if (counters_out) counters_out[0] = counters_out[1] = counters_out[2] = counters_out[3] = 0;
  if (batch == 0) return 0;
  if (!modes || !params || !status_out || !nit_out || !cost_out) return fail(KS_ERR_ARG, "null array argument");
  const int nm = (n - 1) * mode_cols(cls, m);
  const int N = nm + n_free(cmask);
  if (method == 0) {
    if (restart < 1) return fail(KS_ERR_ARG, "gmres restart must be >= 1");
    if (restart > N) restart = N;  // scipy: restart = min(restart, n)
  }
  // _sparse_linalg_factory (optimize.py:1764, :1803-1850): gmres solves the normal equations unless the system is square,
  // i.e. cdof().size == eqn().size, which for these classes means every parameter is constrained
  const bool square = (method == 0 && n_free(cmask) == 0);
  const NkLayout lay = nk_layout(cls, n, m, batch, method, restart, cmask);
  if (!ws || ws_bytes < lay.total) return fail(KS_ERR_WORKSPACE, "workspace too small (see ks_newton_krylov_workspace_bytes)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned char* base = static_cast<unsigned char*>(ws);
  auto at = [&](int k) { return base + lay.off[k]; };
  void* eval_ws = at(0);
  const size_t eval_bytes = ks_workspace_bytes(cls, n, m, batch);
  double* Fy = reinterpret_cast<double*>(at(1));
  double* rhs = reinterpret_cast<double*>(at(2));
  double* vin = reinterpret_cast<double*>(at(3));
  double* wout = reinterpret_cast<double*>(at(4));
  double* xsol = reinterpret_cast<double*>(at(5));
  double* tu = reinterpret_cast<double*>(at(6));
  double* tup = reinterpret_cast<double*>(at(7));
  double* tcost = reinterpret_cast<double*>(at(8));
  double* step = reinterpret_cast<double*>(at(9));
  int* live = reinterpret_cast<int*>(at(10));
  int* searching = reinterpret_cast<int*>(at(11));
  int* counters = reinterpret_cast<int*>(at(12));  // [0] solver active, [1] line search active, [2] live orbits
  if (check_every <= 0) check_every = 8;
  long long n_ops = 0, n_steps = 0, n_ls = 0, n_outer = 0;
  const bool opc = g_op_cache && method != 3 && n == 64 && m == 64 && g_tile_mode == 1 && use_tile(n, m);
  double* n_cache = opc ? reinterpret_cast<double*>(at(19)) : nullptr;

  auto poll = [&](int idx, int& value) -> bool {
    return cudaMemcpyAsync(&value, counters + idx, 4, cudaMemcpyDeviceToHost, s) == cudaSuccess &&
           cudaStreamSynchronize(s) == cudaSuccess;
  };
  auto eval_eqn = [&](const double* u, const double* up, double* f, double* c, bool fill_cache = false) {
    KsEvalArgs a;
    std::memset(&a, 0, sizeof(a));
    a.modes = u; a.params = up; a.out_state = f; a.out_cost = c; a.batch = batch;
    if (fill_cache) a.n_out = n_cache;  // u is fixed until the line search: operator cache
    return dispatch<KS_OP_EQN>(cls, n, m, a, eval_ws, eval_bytes, stream);
  };
  auto op_matvec = [&](const double* v_cdof, double* y) {  // y = J v, v in cdof layout
    KsEvalArgs a;
    std::memset(&a, 0, sizeof(a));
    a.modes = modes; a.params = params; a.vmodes = v_cdof; a.vparams = nullptr; a.out_state = y; a.batch = batch;
    a.cmask = cmask; a.v_ld = N; a.v_tail = 1;
    a.n_cache = n_cache;
    ++n_ops;
    return dispatch<KS_OP_MATVEC>(cls, n, m, a, eval_ws, eval_bytes, stream);
  };
  auto op_rmatvec = [&](const double* y, double* w_cdof) {  // w = J^T y, w in cdof layout
    KsEvalArgs a;
    std::memset(&a, 0, sizeof(a));
    a.modes = modes; a.params = params; a.vmodes = y; a.out_state = w_cdof; a.batch = batch;
    a.cmask = cmask; a.out_ld = N; a.out_tail = 1;
    a.n_cache = n_cache;
    ++n_ops;
    return dispatch<KS_OP_RMATVEC>(cls, n, m, a, eval_ws, eval_bytes, stream);
  };

  // ---- method 3: the dense least-squares step for the `n_live` orbits still in the loop (ks_dense.cuh), chunk by chunk ----
  std::vector<int> h_live, h_idx;
  int rc_dense = 0;
  auto dense_step = [&](int n_live) -> int {
    const int chunk = restart, refine = inner_maxiter;
    const DenseLayout dl = dense_layout(cls, n, m, batch, chunk);
    unsigned char* db = at(15);
    auto dat = [&](int k) { return db + dl.off[k]; };
    int* d_idx = reinterpret_cast<int*>(dat(0));
    double* Uc = reinterpret_cast<double*>(dat(1));
    double* Pc = reinterpret_cast<double*>(dat(2));
    double* rhs0 = reinterpret_cast<double*>(dat(3));
    double* res = reinterpret_cast<double*>(dat(4));
    double* yv = reinterpret_cast<double*>(dat(5));
    double* dxc = reinterpret_cast<double*>(dat(6));
    double* Prep = reinterpret_cast<double*>(dat(7));
    double* J = reinterpret_cast<double*>(dat(8));
    double* G = reinterpret_cast<double*>(dat(9));
    double* Lf = reinterpret_cast<double*>(dat(10));
    double* Li = reinterpret_cast<double*>(dat(11));
    double* lambda = reinterpret_cast<double*>(dat(12));
    double* gmean = reinterpret_cast<double*>(dat(13));
    int* failf = reinterpret_cast<int*>(dat(14));
    int* redo = failf + chunk;
    void* jws = dat(15);
    const size_t jws_bytes = ks_workspace_bytes(cls, n, m, chunk * N);
    // live list: flags to the host (the driver has just synchronised on the live count), indices back
    h_live.resize(batch);
    if (cudaMemcpyAsync(h_live.data(), live, (size_t)batch * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
      return fail(KS_ERR_LAUNCH, "dense Newton step: reading the live flags failed");
    h_idx.clear();
    for (int b = 0; b < batch; ++b)
      if (h_live[b]) h_idx.push_back(b);
    if ((int)h_idx.size() != n_live) return fail(KS_ERR_LAUNCH, "dense Newton step: live list out of step");
    if (cudaMemcpyAsync(d_idx, h_idx.data(), h_idx.size() * 4, cudaMemcpyHostToDevice, s) != cudaSuccess)
      return fail(KS_ERR_LAUNCH, "dense Newton step: writing the live list failed");
    const int nblk = (nm + ksd::kNb - 1) / ksd::kNb;
    const long long sJ = (long long)nm * N, sG = (long long)nm * nm, sLi = (long long)nblk * ksd::kNb * ksd::kNb;
    static DynSmemOptIn diag_optin;
    if (!diag_optin.ensure(ksd::chol_diag_kernel, ksd::kDiagSmemBytes)) return fail(KS_ERR_LAUNCH, "cudaFuncSetAttribute(chol_diag) failed");
    const size_t vec_smem = (size_t)(nm + ksd::kNb) * 8, x_smem = (size_t)N * 8;
    for (int c0 = 0; c0 < n_live; c0 += chunk) {
      const int C = n_live - c0 < chunk ? n_live - c0 : chunk;
      const int* idx = d_idx + c0;
      KS_LAUNCH(ksd::gather_kernel, C, 256, 0, s, idx, (const double*)modes, (const double*)params, (const double*)Fy, Uc, Pc, rhs0, nm);
      // 32x32, matrix-free (default): G = J J^T column by column as J (J^T e_j) -- two launches of the one-orbit-per-warp
      // kernel (rmatvec over unit vectors made in the kernel, then matvec over the result) instead of a Jacobian and a
      // 2 Nm^2 N flop product; the products J^T y and J dx of the solve are single matvec / rmatvec launches.
      const bool mfree = use_warp32(n, m) && g_dense_jac_mode == 2;
      if (mfree) {
        KsEvalArgs a;  // W[c][j][:] = J^T e_j in the cdof layout; W lives where the Jacobian would
        std::memset(&a, 0, sizeof(a));
        a.modes = Uc; a.params = Pc; a.out_state = J; a.batch = C * nm; a.cmask = cmask; a.u_div = nm; a.unit_v = 1;
        a.out_ld = N; a.out_tail = 1;
        n_ops += nm;
        if ((rc_dense = dispatch<KS_OP_RMATVEC>(cls, n, m, a, jws, jws_bytes, stream)) != 0) return rc_dense;
      } else if (use_warp32(n, m) && g_dense_jac_mode) {
      // Jacobian: one batched matvec over the identity; column `col` of orbit c is evaluation number c * N + col
        // 32x32: the one-orbit-per-warp kernel builds the unit vectors itself and reads u / parameters of orbit c for
        // every column (no replicated copies, no identity in memory)
        KsEvalArgs a;
        std::memset(&a, 0, sizeof(a));
        a.modes = Uc; a.params = Pc; a.out_state = J; a.batch = C * N; a.cmask = cmask; a.u_div = N; a.unit_v = 1;
        n_ops += N;
        if ((rc_dense = dispatch<KS_OP_MATVEC>(cls, n, m, a, jws, jws_bytes, stream)) != 0) return rc_dense;
      } else {
        double* Urep = Lf;
        double* Vrep = G;
        KS_LAUNCH(ksd::replicate_kernel, 1184, 256, 0, s, (const double*)Uc, Urep, nm, N, (size_t)C * N * nm);
        KS_LAUNCH(ksd::replicate_kernel, 296, 256, 0, s, (const double*)Pc, Prep, 3, N, (size_t)C * N * 3);
        KS_LAUNCH(ksd::identity_kernel, 1184, 256, 0, s, Vrep, N, (size_t)C * N * N);
        KsEvalArgs a;
        std::memset(&a, 0, sizeof(a));
        a.modes = Urep; a.params = Prep; a.vmodes = Vrep; a.vparams = nullptr; a.out_state = J; a.batch = C * N;
        a.cmask = cmask; a.v_ld = N; a.v_tail = 1;
        n_ops += N;
        if ((rc_dense = dispatch<KS_OP_MATVEC>(cls, n, m, a, jws, jws_bytes, stream)) != 0) return rc_dense;
      }
      // G = J J^T, then G (+ lambda I) = L L^T; orbits whose factorisation breaks down are redone with a growing shift
      for (int attempt = 0; attempt < 5; ++attempt) {
        KS_CHECK_CUDA(cudaMemsetAsync(redo, 0, 4, s), "solver driver setup");
        KS_LAUNCH(ksd::lambda_kernel, (C + 255) / 256, 256, 0, s, (const double*)gmean, lambda, failf, redo, attempt, C);
        if (attempt > 0) {
          int n_redo = 0;
          if (cudaMemcpyAsync(&n_redo, redo, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
            return fail(KS_ERR_LAUNCH, "dense Newton step: device error while checking the factorisation");
          if (n_redo == 0) break;
        }
        if (mfree) {
          KsEvalArgs a;  // column j of G = J w_j, w_j = row j of W
          std::memset(&a, 0, sizeof(a));
          a.modes = Uc; a.params = Pc; a.vmodes = J; a.v_ld = N; a.v_tail = 1; a.out_state = G; a.batch = C * nm; a.cmask = cmask;
          a.u_div = nm;
          n_ops += nm;
          if ((rc_dense = dispatch<KS_OP_MATVEC>(cls, n, m, a, jws, jws_bytes, stream)) != 0) return rc_dense;
        } else {
        // G = J J^T is symmetric and only its lower triangle is read below: block rows x block columns up to the diagonal
        // (an 8 x 8 partition does 36 of the 64 block products; measured 2 % faster than 4 x 4, finer is slower)
          static const int parts_env = std::getenv("KS_DENSE_PARTS") ? std::atoi(std::getenv("KS_DENSE_PARTS")) : 0;  // tuning
          const int parts = parts_env > 0 ? parts_env : (nm >= 512 ? 8 : (nm >= 128 ? 2 : 1));  // 8: 36 of 64 block products
          const int bs = (((nm + parts - 1) / parts) + ksd::kNb - 1) / ksd::kNb * ksd::kNb;  // multiple of the Cholesky block
          for (int r0 = 0; r0 < nm; r0 += bs) {
            const int mr = nm - r0 < bs ? nm - r0 : bs;
            for (int c0g = 0; c0g <= r0; c0g += bs) {
              const int mc = nm - c0g < bs ? nm - c0g : bs;
              if ((rc_dense = ksdh::dgemm_batched(s, false, true, mr, mc, N, 1.0, J + r0, nm, sJ, J + c0g, nm, sJ, 0.0,
                                                  G + r0 + (size_t)c0g * nm, nm, sG, C)) != 0)
                return rc_dense;
            }
          }
        }
        if (attempt == 0) KS_LAUNCH(ksd::diag_mean_kernel, C, ksd::kThreads, 0, s, (const double*)G, gmean, nm);
        else KS_LAUNCH(ksd::shift_diag_kernel, C, 256, 0, s, G, (const double*)lambda, nm);
        for (int blk = 0; blk < nblk; ++blk) {
          const int k0 = blk * ksd::kNb, b = nm - k0 < ksd::kNb ? nm - k0 : ksd::kNb, k1 = k0 + b, rest = nm - k1;
          KS_LAUNCH(ksd::chol_diag_kernel, C, ksd::kThreads, ksd::kDiagSmemBytes, s, (const double*)G, Lf, Li, failf, nm, k0, b, blk);
          if (rest > 0) {
            // panel: L[k1:, k0:k1] = G[k1:, k0:k1] L_kk^-T ; trailing: G[k1:, k1:] -= L[k1:, k0:k1] L[k1:, k0:k1]^T
            if ((rc_dense = ksdh::dgemm_batched(s, false, true, rest, b, b, 1.0, G + k1 + (size_t)k0 * nm, nm, sG,
                                                Li + (size_t)blk * ksd::kNb * ksd::kNb, ksd::kNb, sLi, 0.0, Lf + k1 + (size_t)k0 * nm, nm, sG, C)) != 0)
              return rc_dense;
            // Two-level blocking: inside an outer panel of `outer` Cholesky blocks only the panel's own columns are updated after
            // each block (rank 64); the columns to its right get the whole panel's contribution at once (rank 64 * outer), in
            // column strips of 4 blocks, lower triangle only (each strip includes its diagonal square).  Rank-64 updates of the
            // whole trailing matrix ran at 15 TFLOP/s, the rank-256 ones run at the rate of the J J^T products.
            const int outer = g_dense_outer;
            const int K0 = (blk / outer) * outer * ksd::kNb;
            const int K1 = K0 + outer * ksd::kNb < nm ? K0 + outer * ksd::kNb : nm;
            if (K1 > k1) {
              if ((rc_dense = ksdh::dgemm_batched(s, false, true, nm - k1, K1 - k1, b, -1.0, Lf + k1 + (size_t)k0 * nm, nm, sG,
                                                  Lf + k1 + (size_t)k0 * nm, nm, sG, 1.0, G + k1 + (size_t)k1 * nm, nm, sG, C)) != 0)
                return rc_dense;
            } else {  // last block of the outer panel: everything to the right, rank K1 - K0
              for (int c1s = K1; c1s < nm; c1s += 4 * ksd::kNb) {
                const int wc = nm - c1s < 4 * ksd::kNb ? nm - c1s : 4 * ksd::kNb;
                if ((rc_dense = ksdh::dgemm_batched(s, false, true, nm - c1s, wc, K1 - K0, -1.0, Lf + c1s + (size_t)K0 * nm, nm, sG,
                                                    Lf + c1s + (size_t)K0 * nm, nm, sG, 1.0, G + c1s + (size_t)c1s * nm, nm, sG, C)) != 0)
                  return rc_dense;
              }
            }
          }
        }
      }
      // dx = J^T (L L^T)^-1 (-F), then `refine` rounds of r = -F - J dx, dx += J^T (L L^T)^-1 r
      for (int it = 0; it <= refine; ++it) {
        const double* r = rhs0;
        if (it > 0) {
          if (mfree) {  // res = rhs0 - J dx
            KsEvalArgs a;
            std::memset(&a, 0, sizeof(a));
            a.modes = Uc; a.params = Pc; a.vmodes = dxc; a.v_ld = N; a.v_tail = 1; a.out_state = res; a.batch = C; a.cmask = cmask;
            if ((rc_dense = dispatch<KS_OP_MATVEC>(cls, n, m, a, jws, jws_bytes, stream)) != 0) return rc_dense;
            KS_LAUNCH(ksd::residual_kernel, 148, 256, 0, s, res, (const double*)rhs0, (size_t)C * nm);
          } else {
            KS_LAUNCH(ksd::gemv_kernel, C, ksd::kThreads, x_smem, s, (const double*)J, (const double*)dxc, (const double*)rhs0, res, nm, N, 1, 0);
          }
          r = res;
        }
        KS_LAUNCH(ksd::tri_solve_kernel, C, ksd::kThreads, vec_smem, s, (const double*)Lf, (const double*)Li, r, yv, nm);
        if (mfree) {  // dx (+)= J^T y; the W area is free once the factorisation stands
          double* dst = it > 0 ? J : dxc;
          KsEvalArgs a;
          std::memset(&a, 0, sizeof(a));
          a.modes = Uc; a.params = Pc; a.vmodes = yv; a.out_state = dst; a.out_ld = N; a.out_tail = 1; a.batch = C; a.cmask = cmask;
          if ((rc_dense = dispatch<KS_OP_RMATVEC>(cls, n, m, a, jws, jws_bytes, stream)) != 0) return rc_dense;
          if (it > 0) KS_LAUNCH(ksd::add_kernel, 148, 256, 0, s, dxc, (const double*)dst, (size_t)C * N);
        } else {
          KS_LAUNCH(ksd::gemv_kernel, C, ksd::kThreads, vec_smem, s, (const double*)J, (const double*)yv, (const double*)nullptr, dxc, nm, N, 0,
                    it > 0 ? 1 : 0);
        }
      }
      KS_LAUNCH(ksd::scatter_kernel, C, 256, 0, s, idx, (const double*)dxc, xsol, N);
      n_steps += C;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "dense Newton step: kernel launch failed");
  };

  // initial cost and loop mask.  The line search evaluates the whole batch at (tu, tup) every step: slots of orbits that are
  // never live keep this copy of their orbit instead of uninitialised workspace
  KS_CHECK_CUDA(cudaMemcpyAsync(tu, modes, (size_t)batch * nm * 8, cudaMemcpyDeviceToDevice, s), "solver driver setup");
  KS_CHECK_CUDA(cudaMemcpyAsync(tup, params, (size_t)batch * 24, cudaMemcpyDeviceToDevice, s), "solver driver setup");
  int rc = eval_eqn(modes, params, Fy, cost_out);
  if (rc != 0) return rc;
  KS_LAUNCH(nk_init_kernel, (batch + 255) / 256, 256, 0, s, status_out, nit_out, live, searching, cost_out, tol, batch);

  for (int outer = 0; outer <= maxiter; ++outer) {
    KS_CHECK_CUDA(cudaMemsetAsync(counters, 0, 16, s), "solver driver setup");
    KS_LAUNCH(count_flags_kernel, (batch + 255) / 256, 256, 0, s, live, batch, counters + 2);
    int n_live = 0;
    if (!poll(2, n_live)) return fail(KS_ERR_LAUNCH, "newton-krylov: device error while polling");
    if (n_live == 0) break;
    ++n_outer;
    // right-hand side: -F  (and A^T(-F) for the normal equations)
    if ((rc = eval_eqn(modes, params, Fy, nullptr, opc)) != 0) return rc;
    if (method == 3) {
      if ((rc = dense_step(n_live)) != 0) return rc;
    } else if (method == 0) {
      if (square) {  // every parameter constrained: GMRES on J dx = -F itself                     (:1803-1850)
        KS_LAUNCH(kskry::negate_kernel, 296, 256, 0, s, Fy, rhs, (size_t)batch * nm);
      } else {       // normal equations A^T A dx = A^T (-F)                                        (:1764-1797)
        KS_LAUNCH(kskry::negate_kernel, 296, 256, 0, s, Fy, Fy, (size_t)batch * nm);
        if ((rc = op_rmatvec(Fy, rhs)) != 0) return rc;
      }
      kskry::GmresState g;
      g.V = reinterpret_cast<double*>(at(15)); g.H = reinterpret_cast<double*>(at(16));
      g.G = reinterpret_cast<double*>(at(17)); g.S = reinterpret_cast<double*>(at(18));
      g.x = xsol; g.b = rhs; g.vin = vin; g.w = wout;
      g.sc = reinterpret_cast<double*>(at(13)); g.ic = reinterpret_cast<int*>(at(14));
      g.params = params; g.live = live; g.n_active = counters;
      g.B = batch; g.N = N; g.R = restart; g.maxiter = inner_maxiter; g.n = n; g.m = m; g.cols = mode_cols(cls, m);
      g.precond = preconditioning; g.cmask = cmask; g.rtol = rtol; g.atol_in = atol;
      KS_LAUNCH(kskry::gmres_start_kernel, batch, kskry::kThreads, 0, s, g);
      // GMRES step: vectors longer than 8192 are split over a cluster of 2 / 4 / 8 CTAs per orbit, each holding its slice
      // in registers; shorter ones (and N > 65536, and the debug switch) use the one-CTA-per-orbit kernel
      int gm_cluster = 0;
      // mode 1 (auto): register-resident vector for N <= 4096 (one CTA, 8 or 16 elements per thread) and N > 8192 (clusters);
      // in between (and mode 0) the basis-in-memory kernel
      if (g_gmres_cluster > 1 || (g_gmres_cluster == 1 && (N > kskry::kSlice || N <= 16 * kskry::kThreads))) {
        gm_cluster = g_gmres_cluster;
        while (gm_cluster < kskry::kMaxCluster && (size_t)gm_cluster * kskry::kSlice < (size_t)N) gm_cluster *= 2;
        if ((size_t)gm_cluster * kskry::kSlice < (size_t)N) gm_cluster = 0;
      }
      const long max_steps = (long)inner_maxiter * (restart + 1) + 2;
      for (long st = 0; st < max_steps; ++st) {
        if (st % check_every == 0) {
          int active = 0;
          if (!poll(0, active)) return fail(KS_ERR_LAUNCH, "newton-krylov: device error while polling");
          if (active == 0) break;
        }
        if (square) {
          if ((rc = op_matvec(vin, wout)) != 0) return rc;
        } else {
          if ((rc = op_matvec(vin, Fy)) != 0) return rc;
          if ((rc = op_rmatvec(Fy, wout)) != 0) return rc;
        }
        KS_CHECK_CUDA(cudaMemsetAsync(counters, 0, 4, s), "solver driver setup");
        if (gm_cluster > 0) {
          cudaError_t ce;
          if (gm_cluster == 1 && N <= 8 * kskry::kThreads)
            ce = KS_LAUNCH_CLUSTER(kskry::gmres_step_cluster_kernel<8>, (unsigned)batch, (unsigned)kskry::kThreads, 1u,
                                   (size_t)kskry::kClusterSmemBytes, s, g);
          else if (gm_cluster == 1 && N <= 16 * kskry::kThreads)
            ce = KS_LAUNCH_CLUSTER(kskry::gmres_step_cluster_kernel<16>, (unsigned)batch, (unsigned)kskry::kThreads, 1u,
                                   (size_t)kskry::kClusterSmemBytes, s, g);
          else
            ce = KS_LAUNCH_CLUSTER(kskry::gmres_step_cluster_kernel<32>, (unsigned)batch * gm_cluster, (unsigned)kskry::kThreads,
                                   (unsigned)gm_cluster, (size_t)kskry::kClusterSmemBytes, s, g);
          if (ce != cudaSuccess) return fail(KS_ERR_LAUNCH, "newton-krylov: cluster launch of the GMRES step failed");
        } else {
          KS_LAUNCH(kskry::gmres_step_kernel, batch, kskry::kThreads, 0, s, g);
        }
        ++n_steps;
      }
    } else {
      KS_LAUNCH(kskry::negate_kernel, 296, 256, 0, s, Fy, rhs, (size_t)batch * nm);
      kskry::LsqrState g;
      g.u = reinterpret_cast<double*>(at(15)); g.v = reinterpret_cast<double*>(at(16));
      g.wv = reinterpret_cast<double*>(at(17)); g.hbar = reinterpret_cast<double*>(at(18)); g.lsmr = (method == 2) ? 1 : 0;
      g.x = xsol; g.b = rhs; g.vin = vin; g.w = wout;
      g.sc = reinterpret_cast<double*>(at(13)); g.ic = reinterpret_cast<int*>(at(14));
      g.live = live; g.n_active = counters;
      g.B = batch; g.M = nm; g.N = N; g.ld = N; g.iter_lim = inner_maxiter; g.atol = atol; g.btol = btol; g.conlim = conlim;
      KS_LAUNCH(kskry::lsqr_start_kernel, batch, kskry::kThreads, 0, s, g);
      const long max_steps = 2L * inner_maxiter + 2;
      for (long st = 0; st < max_steps; ++st) {
        if (st % check_every == 0) {
          int active = 0;
          if (!poll(0, active)) return fail(KS_ERR_LAUNCH, "newton-krylov: device error while polling");
          if (active == 0) break;
        }
        // steps alternate: A^T u (even), A v (odd); vin / wout are strided by N in both cases
        if (st % 2 == 0) {
          KsEvalArgs a;
          std::memset(&a, 0, sizeof(a));
          a.modes = modes; a.params = params; a.vmodes = vin; a.out_state = wout; a.batch = batch; a.cmask = cmask;
          a.v_ld = N; a.out_ld = N; a.out_tail = 1;
          a.n_cache = n_cache;
          ++n_ops;
          rc = dispatch<KS_OP_RMATVEC>(cls, n, m, a, eval_ws, eval_bytes, stream);
        } else {
          KsEvalArgs a;
          std::memset(&a, 0, sizeof(a));
          a.modes = modes; a.params = params; a.vmodes = vin; a.out_state = wout; a.batch = batch; a.cmask = cmask;
          a.v_ld = N; a.v_tail = 1; a.out_ld = N;
          a.n_cache = n_cache;
          ++n_ops;
          rc = dispatch<KS_OP_MATVEC>(cls, n, m, a, eval_ws, eval_bytes, stream);
        }
        if (rc != 0) return rc;
        KS_CHECK_CUDA(cudaMemsetAsync(counters, 0, 4, s), "solver driver setup");
        KS_LAUNCH(kskry::lsqr_step_kernel, batch, kskry::kThreads, 0, s, g);
        ++n_steps;
      }
    }
    // line search on u + step * dx                                                       (:1373-1399)
    kskry::NewtonState ns;
    ns.u = modes; ns.up = params; ns.dx = xsol; ns.tu = tu; ns.tup = tup; ns.tcost = tcost; ns.cost = cost_out;
    ns.step = step; ns.status = status_out; ns.nit = nit_out; ns.live = live; ns.searching = searching;
    ns.n_searching = counters + 1; ns.B = batch; ns.nm = nm; ns.N = N; ns.maxiter = maxiter; ns.cmask = cmask;
    ns.tol = tol; ns.ftol = ftol; ns.min_step = min_step;
    ns.mode = 0;
    KS_CHECK_CUDA(cudaMemsetAsync(counters + 1, 0, 4, s), "solver driver setup");
    KS_LAUNCH(kskry::newton_step_kernel, batch, kskry::kThreads, 0, s, ns);
    ns.mode = 1;
    for (int ls = 0; ls < 1200; ++ls) {
      if ((rc = eval_eqn(tu, tup, Fy, tcost)) != 0) return rc;
      KS_CHECK_CUDA(cudaMemsetAsync(counters + 1, 0, 4, s), "solver driver setup");
      KS_LAUNCH(kskry::newton_step_kernel, batch, kskry::kThreads, 0, s, ns);
      ++n_ls;
      int still = 0;
      if (!poll(1, still)) return fail(KS_ERR_LAUNCH, "newton-krylov: device error while polling");
      if (still == 0) break;
    }
  }
  if (counters_out) {
    counters_out[0] = n_ops; counters_out[1] = n_steps; counters_out[2] = n_ls; counters_out[3] = n_outer;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "newton-krylov launch failed");
}
}  // namespace

extern "C" {
int ks_newton_krylov(int cls, int n, int m, int batch, double* modes, double* params, uint32_t cmask, int method,
                     double tol, double ftol, int maxiter, double min_step, int preconditioning, int restart,
                     int inner_maxiter, double rtol, double atol, double btol, double conlim, int32_t* status_out,
                     int32_t* nit_out, double* cost_out, long long* counters_out, int check_every, void* ws,
                     size_t ws_bytes, void* stream) {
  if (method < 0 || method > 2) return fail(KS_ERR_ARG, "bad method (0 gmres, 1 lsqr, 2 lsmr)");
  return newton_driver(cls, n, m, batch, modes, params, cmask, method, tol, ftol, maxiter, min_step, preconditioning, restart,
                       inner_maxiter, rtol, atol, btol, conlim, status_out, nit_out, cost_out, counters_out, check_every, ws, ws_bytes,
                       stream);
}

size_t ks_newton_dense_workspace_bytes(int cls, int n, int m, int batch, int chunk) {
  if (!shape_ok(cls, n, m) || batch <= 0 || chunk <= 0) return 0;
  return nk_layout(cls, n, m, batch, 3, chunk, 0).total;
}

int ks_newton_dense(int cls, int n, int m, int batch, double* modes, double* params, uint32_t cmask, double tol, double ftol,
                    int maxiter, double min_step, int chunk, int refine, int32_t* status_out, int32_t* nit_out,
                    double* cost_out, long long* counters_out, void* ws, size_t ws_bytes, void* stream) {
  if (chunk <= 0 || refine < 0) return fail(KS_ERR_ARG, "dense Newton step: chunk must be positive and refine non-negative");
  if (shape_ok(cls, n, m) && (size_t)(n - 1) * mode_cols(cls, m) > 5000)
    return fail(KS_ERR_SHAPE, "dense Newton step: more than 5000 modes per orbit (use the matrix-free methods)");
  return newton_driver(cls, n, m, batch, modes, params, cmask, 3, tol, ftol, maxiter, min_step, 0, chunk, refine, 0.0, 0.0, 0.0, 0.0,
                       status_out, nit_out, cost_out, counters_out, 8, ws, ws_bytes, stream);
}

}  // extern "C"
