// Launch code of the tile path for ONE symmetry class (KS_TILE_CLS), included by ks_tile_c{0..3}.cu so the four classes
// compile in parallel.  Everything here has internal linkage except ks_tile_run_c<CLS>.
#pragma once
#include <cstring>

#include "ks_host.cuh"
#include "ks_tile.cuh"
#include "ks_tile64.cuh"
#include "ks_tile_tma.cuh"

#ifndef KS_TILE_CLS
#error "define KS_TILE_CLS (0..3) before including ks_tile_host.cuh"
#endif

namespace {
using namespace ksh;
constexpr int kCls = KS_TILE_CLS;
template <class K>
int tile_ctas_per_sm(K kern, size_t smem) {
#ifdef KS_HOST_EMUL
  (void)kern; (void)smem;
  return 2;
#else
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kst::kThreads, smem) != cudaSuccess || nb < 1) nb = 1;
  return nb;
#endif
}
template <class K>
int tile_launch(K kern, size_t smem, int ntiles, const kst::TileArgs& ta, cudaStream_t s) {
  KS_LAUNCH(kern, ntiles, kst::kThreads, smem, s, ta);  // one CTA per tile
  return 0;
}
int tile_ensure_twiddles(cudaStream_t s) {
#ifdef KS_HOST_EMUL
  static bool done = false;
  if (!done) { KS_LAUNCH(kst::tw_init_kernel, 2, 256, 0, s, 0); done = true; }
#else
  static bool done[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return fail(KS_ERR_LAUNCH, "cudaGetDevice failed");
  if (!done[dev]) {
    KS_LAUNCH(kst::tw_init_kernel, 2, 256, 0, s, 0);
    if (cudaStreamSynchronize(s) != cudaSuccess) return fail(KS_ERR_LAUNCH, "twiddle table kernel failed");
    done[dev] = true;
  }
#endif
  return 0;
}
#ifndef KS_HOST_EMUL
// ---- TMA-staged persistent pass kernels (ks_tile_tma.cuh) ----
typedef CUresult (*KsEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
KsEncodeTiledFn tile_encode_fn() {
  static KsEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<KsEncodeTiledFn>(p);
  }
  return fn;
}
// XC[chunk][K][n] double2 as a 2-D float64 tensor: inner 2 n doubles (one line), outer chunk * K lines; box {2 TB, K}
bool tile_xc_tensor_map(CUtensorMap* tm, double2* XC, int chunk, int K, int n, int TB) {
  KsEncodeTiledFn enc = tile_encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)2 * n, (cuuint64_t)chunk * K};
  const cuuint64_t strides[1] = {(cuuint64_t)n * 16};
  const cuuint32_t box[2] = {(cuuint32_t)2 * TB, (cuuint32_t)K};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, XC, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
template <class K>
int tile_persistent_grid(K kern, size_t smem, int ntiles, DynSmemOptIn& optin) {
  if (!optin.ensure(kern, (int)smem)) return -1;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kst::kThreads, smem) != cudaSuccess || nb < 1) nb = 1;
  const int cap = nb * sm_count();
  return ntiles < cap ? ntiles : cap;
}
template <int M, int MODE, int STAGES>
int tile_row_tma_ms(const kst::TileArgs& ta, cudaStream_t s) {
  using P = kstma::RowPlan<M, STAGES>;
  auto kern = kstma::tile_row_tma_kernel<M, MODE, STAGES>;
  static DynSmemOptIn optin;
  const int grid = tile_persistent_grid(kern, P::bytes, ta.e.batch * ta.ntb, optin);
  if (grid < 0) return fail(KS_ERR_LAUNCH, "cudaFuncSetAttribute(row tma) failed");
  CUtensorMap tm;
  if (!tile_xc_tensor_map(&tm, ta.XC, ta.e.batch, P::K, ta.g.n, P::TB)) return fail(KS_ERR_LAUNCH, "cuTensorMapEncodeTiled failed");
  KS_LAUNCH(kern, grid, kst::kThreads, P::bytes, s, ta, tm);
  return 0;
}
template <int M, int MODE>
int tile_row_tma_m(const kst::TileArgs& ta, cudaStream_t s) {
  return tile_row_tma_ms<M, MODE, 2>(ta, s);
}
template <int MODE>
int tile_row_tma(const kst::TileArgs& ta, cudaStream_t s) {
  switch (ta.g.m) {
    case 64: return tile_row_tma_m<64, MODE>(ta, s);
    case 128: return tile_row_tma_m<128, MODE>(ta, s);
    default: return tile_row_tma_m<256, MODE>(ta, s);
  }
}
template <int CLS, int OP, int N, int STAGE, int STAGES>
int tile_colfwd_tma_s(const kst::TileArgs& ta, cudaStream_t s) {
  using P = kstma::ColPlan<N, STAGES>;
  auto kern = kstma::tile_colfwd_tma_kernel<CLS, OP, N, STAGE, STAGES>;
  static DynSmemOptIn optin;
  const int grid = tile_persistent_grid(kern, P::bytes, ta.e.batch * ta.nkb, optin);
  if (grid < 0) return fail(KS_ERR_LAUNCH, "cudaFuncSetAttribute(col tma) failed");
  KS_LAUNCH(kern, grid, kst::kThreads, P::bytes, s, ta);
  return 0;
}
template <int CLS, int OP, int N, int STAGE>
int tile_colfwd_tma(const kst::TileArgs& ta, cudaStream_t s) {
  return tile_colfwd_tma_s<CLS, OP, N, STAGE, 2>(ta, s);
}
#endif

template <int CLS, int N>
int tile_colinv(const kst::TileArgs& ta, cudaStream_t s) {
  return tile_launch(kst::tile_colinv_kernel<CLS, N>, kst::tile_smem_bytes<N>(), ta.e.batch * ta.nkb, ta, s);
}
template <int CLS, int OP, int N, int STAGE>
int tile_colfwd(const kst::TileArgs& ta, cudaStream_t s) {
  return tile_launch(kst::tile_colfwd_kernel<CLS, OP, N, STAGE>, kst::tile_smem_bytes<N>(), ta.e.batch * ta.nkb, ta, s);
}
template <int MODE>
int tile_row(const kst::TileArgs& ta, cudaStream_t s) {
  const int nt = ta.e.batch * ta.ntb;
  switch (ta.g.m) {
    case 64: return tile_launch(kst::tile_row_kernel<64, MODE>, kst::tile_smem_bytes<64>(), nt, ta, s);
    case 128: return tile_launch(kst::tile_row_kernel<128, MODE>, kst::tile_smem_bytes<128>(), nt, ta, s);
    default: return tile_launch(kst::tile_row_kernel<256, MODE>, kst::tile_smem_bytes<256>(), nt, ta, s);
  }
}
template <int CLS, int OP, int N>
int tile_chunk_run(kst::TileArgs& ta, cudaStream_t s) {
  const bool nuu = (OP == KS_OP_EQN || OP == KS_OP_COSTGRAD) || !(ta.e.cmask & KS_FIX_X);
  tile_colinv<CLS, N>(ta, s);
  ta.forward = nuu ? 1 : 0;
  ta.zero_nuu = nuu ? 0 : 1;
#ifndef KS_HOST_EMUL
  const int tma = ks_tile_tma_mode();
  const bool tma_rows = tma == 1 || tma == 2, tma_cols = tma == 1 || tma == 3;
#else
  const bool tma_rows = false, tma_cols = false;
#endif
  int rc = 0;
  auto row = [&](auto mode_c) {
    constexpr int MODE = decltype(mode_c)::value;
#ifndef KS_HOST_EMUL
    if (tma_rows) return tile_row_tma<MODE>(ta, s);
#endif
    return tile_row<MODE>(ta, s);
  };
  auto col = [&](auto stage_c) {
    constexpr int STAGE = decltype(stage_c)::value;
#ifndef KS_HOST_EMUL
    if (tma_cols) return tile_colfwd_tma<CLS, OP, N, STAGE>(ta, s);
#endif
    return tile_colfwd<CLS, OP, N, STAGE>(ta, s);
  };
  if ((rc = row(std::integral_constant<int, kst::MODE_SQ>{})) != 0) return rc;
  if ((rc = col(std::integral_constant<int, 1>{})) != 0) return rc;
  if constexpr (OP == KS_OP_EQN) {
    KS_LAUNCH((kst::tile_finalize_kernel<CLS, OP>), (ta.e.batch + 127) / 128, 128, 0, s, ta);
  }
  if constexpr (OP != KS_OP_EQN) {
    ta.forward = 1;
    if ((rc = row(std::integral_constant<int, kst::MODE_MUL>{})) != 0) return rc;
    if ((rc = col(std::integral_constant<int, 2>{})) != 0) return rc;
  }
  return 0;
}
// 64x64: one CTA per orbit, all passes on chip (ks_tile64.cuh); `chunk` < 0 selects it
template <int CLS, int OP>
int launch_fused64(const KsEvalArgs& a, cudaStream_t s) {
  auto kern = kst::tile_fused64_kernel<CLS, OP>;
  static DynSmemOptIn optin;
  if (!optin.ensure(kern, (int)kst::F64::smem_bytes)) return fail(KS_ERR_LAUNCH, "cudaFuncSetAttribute(fused64) failed");
  kst::TileArgs ta;
  std::memset(&ta, 0, sizeof(ta));
  ta.e = a;
  ta.g = make_geom(CLS, 64, 64);
  ta.TU = a.scratch;  // one stage-1 slot (Nm doubles) per CTA
  const int cap = 2 * sm_count();
  KS_LAUNCH(kern, a.batch < cap ? a.batch : cap, kst::kF64Threads, kst::F64::smem_bytes, s, ta);
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "fused 64x64 kernel launch failed");
}
template <int CLS, int OP>
int launch_tile(int n, int m, int chunk, const KsEvalArgs& a, cudaStream_t s) {
  if (int rc = tile_ensure_twiddles(s)) return rc;
  if (chunk < 0) return launch_fused64<CLS, OP>(a, s);
  const kspw::Geom g = make_geom(CLS, n, m);
  const int KB = kst::col_lines(n), TB = kst::col_lines(m);
  const int nkb = (g.k + KB - 1) / KB, ntb = (n / 2) / TB;
  char* base = reinterpret_cast<char*>(a.scratch);
  auto take = [&](size_t doubles) {
    char* p = base;
    base += (doubles * 8 + 255) & ~(size_t)255;
    return reinterpret_cast<double*>(p);
  };
  kst::TileArgs ta;
  std::memset(&ta, 0, sizeof(ta));
  ta.g = g;
  ta.XC = reinterpret_cast<double2*>(take((size_t)chunk * g.k * n * 2));
  ta.FU2 = reinterpret_cast<double2*>(take((size_t)chunk * n * m));
  ta.TU = take((size_t)chunk * g.nm);
  ta.partial = take((size_t)chunk * nkb * 4);
  ta.nkb = nkb;
  ta.ntb = ntb;
  const size_t v_ld = a.v_ld ? a.v_ld : g.nm, out_ld = a.out_ld ? a.out_ld : g.nm;
  for (int b0 = 0; b0 < a.batch; b0 += chunk) {
    KsEvalArgs e = a;
    e.batch = a.batch - b0 < chunk ? a.batch - b0 : chunk;
    e.modes += (size_t)b0 * g.nm;
    e.params += (size_t)b0 * 3;
    if (e.vmodes) e.vmodes += (size_t)b0 * v_ld;
    if (e.vparams) e.vparams += (size_t)b0 * 3;
    if (e.out_state) e.out_state += (size_t)b0 * out_ld;
    if (e.out_params) e.out_params += (size_t)b0 * 3;
    if (e.out_cost) e.out_cost += b0;
    if (e.f_out) e.f_out += (size_t)b0 * g.nm;
    ta.e = e;
    int rc = 0;
    switch (n) {
      case 64: rc = tile_chunk_run<CLS, OP, 64>(ta, s); break;
      case 128: rc = tile_chunk_run<CLS, OP, 128>(ta, s); break;
      default: rc = tile_chunk_run<CLS, OP, 256>(ta, s); break;
    }
    if (rc != 0) return rc;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "tile kernel launch failed");
}


int tile_run(int op, int n, int m, int chunk, const KsEvalArgs& a, cudaStream_t s) {
  switch (op) {
    case KS_OP_EQN: return launch_tile<kCls, KS_OP_EQN>(n, m, chunk, a, s);
    case KS_OP_COSTGRAD: return launch_tile<kCls, KS_OP_COSTGRAD>(n, m, chunk, a, s);
    case KS_OP_MATVEC: return launch_tile<kCls, KS_OP_MATVEC>(n, m, chunk, a, s);
    case KS_OP_RMATVEC: return launch_tile<kCls, KS_OP_RMATVEC>(n, m, chunk, a, s);
    default: return fail(KS_ERR_ARG, "tile path: unknown op");
  }
}
}  // namespace

#define KS_TILE_CAT_(a, b) a##b
#define KS_TILE_CAT(a, b) KS_TILE_CAT_(a, b)
int KS_TILE_CAT(ks_tile_run_c, KS_TILE_CLS)(int op, int n, int m, int chunk, const KsEvalArgs& a, cudaStream_t s) {
  return tile_run(op, n, m, chunk, a, s);
}
