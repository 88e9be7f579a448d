// TMA-staged variants of the large-grid pass kernels (ks_tile.cuh): the same five passes and the same per-tile arithmetic,
// but every CTA is PERSISTENT and fetches the tile after the one it is transforming with the Tensor Memory Accelerator
// (cp.async.bulk[.tensor] into a two-stage shared-memory ring, completion on an mbarrier), so the global/L2 latency of a
// tile is hidden behind the FFTs of the previous one instead of being paid once per short-lived CTA.
//
// MEASURED (profiles/r02q_tile_tma.jsonl): bit-identical results, but SLOWER than the plain kernels -- 256x256: 0.158 ms
// against 0.127 ms per 64 evaluations, 128x128: 0.144 against 0.119 ms.  The ring costs 65 KB of shared memory per CTA, which
// leaves two resident CTAs per SM where the plain kernels have three, and the passes are bound by the latency of their own
// shared-memory exchange / barrier chain, which only more resident tiles hide; a one-stage ring at three CTAs per SM merely
// drew level with the plain rows (0.133 ms).  The plain kernels therefore stay the default (ks_debug_tile_tma(0)); these are
// selectable with ks_debug_tile_tma(1 | 2 | 3) and are kept parity-tested.
//
//   rows   (passes B, D): the tile's input is the box XC[kp = 0..K-1][t0 .. t0+TB) plus its partner box at t + n/2 --
//          two 2-D tensor boxes (inner 16 TB bytes, K rows of stride 16 n bytes) loaded with cp.async.bulk.tensor.2d; the
//          results go back through the same staging buffer with cp.async.bulk.tensor.2d stores (bulk async-group).
//   columns (passes C, E): the tile's input is KB consecutive lines of XC = one contiguous block, loaded with the
//          non-tensor cp.async.bulk; u's / F's tuples still come through ordinary loads issued before the wait.
// Pass A (modes -> XC) keeps the plain kernel: its input is the strided real-packed modes array, read once.
//
// Reference path replaced: as ks_tile.cuh (orbithunter/ks/orbits.py transforms :2212-2420 and the eqn / matvec / rmatvec /
// costgrad bodies).  Not built for the CPU emulation (KS_HOST_EMUL): the plain kernels are the cross-check.
#pragma once
#ifndef KS_HOST_EMUL
#include <cuda.h>

#include "ks_tile.cuh"

namespace kstma {
using namespace kst;

// ---- PTX wrappers ------------------------------------------------------------------------------------------------------
KS_DEVI unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
KS_DEVI void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
KS_DEVI void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
KS_DEVI void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
KS_DEVI void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
KS_DEVI void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<unsigned long long>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
KS_DEVI void tma_store_2d(const CUtensorMap* tm, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<unsigned long long>(tm)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
KS_DEVI void bulk_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
KS_DEVI void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
KS_DEVI void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
KS_DEVI void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
KS_DEVI void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- shared-memory plans -------------------------------------------------------------------------------------------------
// rows: [STAGES][2 boxes][K][TB] double2 | line buffers | twiddles | 2 mbarriers
template <int M, int STAGES>
struct RowPlan {
  using F = LineFft<M>;
  static constexpr int TB = kThreads / F::T, K = M / 2 - 1;
  static constexpr int box = K * TB;                 // double2 per box
  static constexpr int stage = 2 * box;              // double2 per stage
  static constexpr size_t off_lines = (size_t)STAGES * stage * 16;
  static constexpr size_t off_tab = off_lines + (size_t)TB * F::LS * 16;
  static constexpr size_t off_bar = off_tab + (size_t)F::TW * 16;
  static constexpr size_t bytes = off_bar + 16;
};
// columns: [2 stages][KB][N] double2 | line buffers | twiddles | reduction scratch | 2 mbarriers
template <int N, int STAGES>
struct ColPlan {
  using F = LineFft<N>;
  static constexpr int KB = kThreads / F::T;
  static constexpr int stage = KB * N;  // double2 per stage
  static constexpr size_t off_lines = (size_t)STAGES * stage * 16;
  static constexpr size_t off_tab = off_lines + (size_t)KB * F::LS * 16;
  static constexpr size_t off_red = off_tab + (size_t)F::TW * 16;
  static constexpr size_t off_bar = off_red + kWarps * 4 * 8;
  static constexpr size_t bytes = off_bar + 16;
};

// ---- passes B / D: rows, persistent, TMA boxes in and out ----------------------------------------------------------------
// tmXC: 2-D tensor map over XC viewed as doubles: inner dimension 2 n (one line), outer dimension chunk * K lines; box
// {2 TB, K}.  Tile `tile` = (orbit, tb): boxes at coordinates (2 t0, orbit K) and (2 (t0 + n/2), orbit K).
template <int M, int MODE, int STAGES = 2>
__global__ void __launch_bounds__(kThreads, 2) tile_row_tma_kernel(const TileArgs ta, const __grid_constant__ CUtensorMap tmXC) {
  using P = RowPlan<M, STAGES>;
  using F = LineFft<M>;
  constexpr int TB = P::TB, K = P::K, NIT = (K * TB + kThreads - 1) / kThreads;
  constexpr unsigned kStageBytes = (unsigned)P::stage * 16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* stages = reinterpret_cast<double2*>(smem_raw);
  double2* lines = reinterpret_cast<double2*>(smem_raw + P::off_lines);
  double2* tab = reinterpret_cast<double2*>(smem_raw + P::off_tab);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + P::off_bar);
  const Geom g = ta.g;
  const int n = g.n, n2 = n / 2;
  const int ll = threadIdx.x / F::T, gi = threadIdx.x % F::T;
  const int tl = threadIdx.x % TB, kq = threadIdx.x / TB;
  const int ntiles = ta.e.batch * ta.ntb;
  F::fill_table(tab);
  F f;
  f.init(gi, tab);
  double2* ln = lines + ll * F::LS;
  double2* l2 = lines + tl * F::LS;
  if (threadIdx.x == 0) {
    mbar_init(bar + 0, 1);
    mbar_init(bar + 1, 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int tile, int s) {  // one thread: arm the stage's barrier and start both box loads
    const int orbit = tile / ta.ntb, t0 = (tile - orbit * ta.ntb) * TB;
    mbar_expect_tx(bar + s, kStageBytes);
    tma_load_2d(stages + s * P::stage, &tmXC, 2 * t0, orbit * K, bar + s);
    tma_load_2d(stages + s * P::stage + P::box, &tmXC, 2 * (t0 + n2), orbit * K, bar + s);
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  int it = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    const int orbit = tile / ta.ntb, t0 = (tile - orbit * ta.ntb) * TB;
    if (threadIdx.x == 0 && tile + (int)gridDim.x < ntiles) {
      bulk_wait_read0();  // the other stage's stores (previous tile) have been read out of shared memory
      issue(tile + gridDim.x, s ^ 1);
    }
    double2* sA = stages + s * P::stage;
    double2* sB = sA + P::box;
    double2* fu = ta.FU2 + ((size_t)orbit * n2 + t0 + ll) * M + gi;
    double2 uf[MODE == MODE_MUL ? F::PT : 1];
    if constexpr (MODE == MODE_MUL) {  // the stored field: issued before the wait, needed after the inverse transform
#pragma unroll
      for (int sI = 0; sI < F::PT; ++sI) uf[sI] = fu[F::post_index(sI)];
    }
    mbar_wait(bar + s, (it >> 1) & 1);
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int kp = kq + i * (kThreads / TB);
      if (kp < K) {
        const double2 A = sA[kp * TB + tl], B = sB[kp * TB + tl];
        l2[kp + 1] = make_double2(A.x - B.y, A.y + B.x);
        l2[M - kp - 1] = make_double2(A.x + B.y, B.x - A.y);
      }
    }
    if (threadIdx.x < TB) {
      lines[threadIdx.x * F::LS] = make_double2(0.0, 0.0);
      lines[threadIdx.x * F::LS + M / 2] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    double vr[F::PT], vi[F::PT];
#pragma unroll
    for (int sI = 0; sI < F::PT; ++sI) {
      const double2 v = ln[gi + F::pre_index(sI)];
      vr[sI] = v.x;
      vi[sI] = v.y;
    }
    f.template transform<-1>(vr, vi, ln);
#pragma unroll
    for (int sI = 0; sI < F::PT; ++sI) {
      const double ua = vr[sI] * g.cx, ub = vi[sI] * g.cx;
      if constexpr (MODE == MODE_SQ) {
        fu[F::post_index(sI)] = make_double2(ua, ub);
        vr[sI] = ua * ua;
        vi[sI] = ub * ub;
      } else {
        vr[sI] = ua * uf[sI].x;
        vi[sI] = ub * uf[sI].y;
      }
    }
    if (ta.forward) {
      F::post_to_pre(vr, vi);
      f.template transform<1>(vr, vi, ln);
#pragma unroll
      for (int sI = 0; sI < F::PT; ++sI) ln[gi + F::post_index(sI)] = make_double2(vr[sI], vi[sI]);
      __syncthreads();
      // results back into this tile's (consumed) staging boxes, then out with two bulk tensor stores
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const int kp = kq + i * (kThreads / TB);
        if (kp < K) {
          const double2 y = l2[kp + 1], z = l2[M - kp - 1];
          sA[kp * TB + tl] = make_double2((y.x + z.x) * g.cx, (y.y - z.y) * g.cx);
          sB[kp * TB + tl] = make_double2((y.y + z.y) * g.cx, (z.x - y.x) * g.cx);
        }
      }
      fence_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        tma_store_2d(&tmXC, 2 * t0, orbit * K, sA);
        tma_store_2d(&tmXC, 2 * (t0 + n2), orbit * K, sB);
        bulk_commit();
      }
    } else {
      __syncthreads();  // every line buffer has been read before the next tile overwrites it
    }
  }
  if (threadIdx.x == 0) bulk_wait_all0();
}

// ---- passes C / E: columns, persistent, the tile's lines arrive as one bulk copy -----------------------------------------
template <int CLS, int OP, int N, int STAGE, int STAGES = 2>
__global__ void __launch_bounds__(kThreads, 2) tile_colfwd_tma_kernel(const TileArgs ta) {
  using P = ColPlan<N, STAGES>;
  using F = LineFft<N>;
  constexpr int KB = P::KB, H1 = N / 2, NIT = H1 * KB / kThreads, JS = kThreads / KB;
  constexpr bool REINV = (STAGE == 1 && OP != KS_OP_EQN);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* stages = reinterpret_cast<double2*>(smem_raw);
  double2* lines = reinterpret_cast<double2*>(smem_raw + P::off_lines);
  double2* tab = reinterpret_cast<double2*>(smem_raw + P::off_tab);
  double* red = reinterpret_cast<double*>(smem_raw + P::off_red);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + P::off_bar);
  const Geom g = ta.g;
  const KsEvalArgs& a = ta.e;
  const int ll = threadIdx.x / F::T, gi = threadIdx.x % F::T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kl = threadIdx.x % KB, j0 = threadIdx.x / KB;
  const int ntiles = a.batch * ta.nkb;
  const bool skip_input = (STAGE == 1 && ta.zero_nuu);
  F::fill_table(tab);
  F f;
  f.init(gi, tab);
  double2* ln = lines + ll * F::LS;
  double2* l2 = lines + kl * F::LS;
  if (threadIdx.x == 0) {
    mbar_init(bar + 0, 1);
    mbar_init(bar + 1, 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int tile, int s) {
    const int orbit = tile / ta.nkb, k0 = (tile - orbit * ta.nkb) * KB;
    const int nl = g.k - k0 < KB ? g.k - k0 : KB;
    const unsigned bytes = (unsigned)nl * N * 16;
    mbar_expect_tx(bar + s, bytes);
    bulk_load_1d(stages + s * P::stage, ta.XC + ((size_t)orbit * g.k + k0) * N, bytes, bar + s);
  };
  if (!skip_input && threadIdx.x == 0 && (int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  int it = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    const int orbit = tile / ta.nkb, kb = tile - orbit * ta.nkb, k0 = kb * KB;
    double2* XC = ta.XC + (size_t)orbit * g.k * N;
    const kspw::OrbitCtx c = kspw::make_ctx<CLS, OP>(a, g, orbit, ta.TU + (size_t)orbit * g.nm);
    if (!skip_input && threadIdx.x == 0 && tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x, s ^ 1);
    // the tuples of u / v / F this tile's algebra needs: ordinary loads, in flight during the wait and the transform
    const int kpw = k0 + kl;
    const bool act = kpw < g.k;
    Tup Mt[NIT], Vt[(STAGE == 1 && (OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC)) ? NIT : 1];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      if (!act) continue;
      if constexpr (STAGE == 1) {
        if constexpr (OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) kspw::stage1_loads<CLS, OP>(c, g, j0 + i * JS, kpw, Mt[i], Vt[i]);
        else kspw::stage1_loads<CLS, OP>(c, g, j0 + i * JS, kpw, Mt[i], Vt[0]);
      } else {
        Mt[i] = kspw::stage2_load<CLS, OP>(c, g, j0 + i * JS, kpw);
      }
    }
    double vr[F::PT], vi[F::PT];
    {
      const int kp = k0 + ll;
      const bool live = kp < g.k && !skip_input;
      if (!skip_input) mbar_wait(bar + s, (it >> 1) & 1);
      const double2* src = stages + s * P::stage + ll * N + gi;
#pragma unroll
      for (int sI = 0; sI < F::PT; ++sI) {
        const double2 v = live ? src[F::pre_index(sI)] : make_double2(0.0, 0.0);
        vr[sI] = v.x;
        vi[sI] = v.y;
      }
    }
    f.template transform<1>(vr, vi, ln);
#pragma unroll
    for (int sI = 0; sI < F::PT; ++sI) ln[gi + F::post_index(sI)] = make_double2(vr[sI], vi[sI]);
    __syncthreads();
    kspw::Sums sm;
    if (act) {
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const int j = j0 + i * JS;
        Tup Pt;
        if (j == 0) {
          const double c0 = ks::kSqrt2 * g.ct;
          Pt = Tup{l2[0].x * c0, 0.0, l2[0].y * c0, 0.0};
        } else {
          const double2 y = l2[j], z = l2[N - j];
          Pt = Tup{(y.x + z.x) * g.ct, (y.y - z.y) * g.ct, (y.y + z.y) * g.ct, (z.x - y.x) * g.ct};
        }
        if constexpr (STAGE == 1) {
          Tup W{0.0, 0.0, 0.0, 0.0};
          kspw::stage1_point<CLS, OP>(c, g, j, kpw, Pt, Mt[i], Vt[(OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) ? i : 0], W, sm);
          if constexpr (REINV) spec_store(l2, j, N, W);
        } else {
          kspw::stage2_point<CLS, OP>(c, g, j, kpw, Pt, Mt[i]);
        }
      }
    }
    if constexpr (STAGE == 1) {
      const double r0 = warp_sum(sm.ff), r1 = warp_sum(sm.e), r2 = warp_sum(fma(-c.sig, sm.d, sm.x)), r3 = warp_sum(sm.d);
      if (lane == 0) {
        red[warp * 4 + 0] = r0;
        red[warp * 4 + 1] = r1;
        red[warp * 4 + 2] = r2;
        red[warp * 4 + 3] = r3;
      }
      __syncthreads();
      if (threadIdx.x < 4) {
        double t = 0.0;
        for (int wq = 0; wq < kWarps; ++wq) t += red[wq * 4 + threadIdx.x];
        ta.partial[((size_t)orbit * ta.nkb + kb) * 4 + threadIdx.x] = t;
      }
    } else {
      if (OP != KS_OP_MATVEC && kb == 0 && threadIdx.x == 0) finalize_orbit<CLS, OP>(ta, c, orbit);
    }
    if constexpr (REINV) {
#pragma unroll
      for (int sI = 0; sI < F::PT; ++sI) {
        const double2 v = ln[gi + F::pre_index(sI)];
        vr[sI] = v.x;
        vi[sI] = v.y;
      }
      f.template transform<-1>(vr, vi, ln);
      const int kp = k0 + ll;
      if (kp < g.k) {
        double2* o = XC + (size_t)kp * N + gi;
#pragma unroll
        for (int sI = 0; sI < F::PT; ++sI) o[F::post_index(sI)] = make_double2(vr[sI] * g.ct, vi[sI] * g.ct);
      }
    }
    __syncthreads();  // line buffers, reduction scratch and this stage are free for the tile after next
  }
}

}  // namespace kstma
#endif  // KS_HOST_EMUL
