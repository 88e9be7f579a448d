// Large-grid fused spectral kernels (n, m in {64, 128, 256}, any mix): multi-pass, line-block tiles, register radix FFTs.
//
// Same mathematics as ks_generic.cuh / ks_warp32.cuh (tuple algebra in ks_pointwise.cuh).  One evaluation is five passes
// over the orbit's complex spatial-mode array XC[kp][t] (kept in a caller-provided workspace, L2-resident for a chunk of
// orbits), each pass one kernel whose CTAs own a block of lines of ONE orbit, so the grid is (orbits x line blocks) and all
// 148 SMs are busy even for a handful of 256x256 orbits:
//   A  colinv          modes u -> XC                       (k complex length-n inverse FFTs, KB wavenumbers per tile)
//   B  row<MODE_SQ>    XC -> field u (stored) ; u*u -> XC   (n/2 two-for-one complex length-m inverse+forward FFTs)
//   C  colfwd<stage 1> XC -> tuples -> F, scalar partial sums, second spectrum -> XC (forward + inverse FFT per line)
//   D  row<MODE_MUL>   XC -> second field * u -> XC
//   E  colfwd<stage 2> XC -> tuples -> output
// The per-tile partial sums behind cost and parameter components are added in a fixed order (deterministic) by pass E
// (a one-thread-per-orbit finalize kernel for eqn, which has no pass E).  A length-N line FFT is done by T = 8 or 16 adjacent lanes: each lane holds N/T points in
// registers, does radix-8/16 sub-FFTs there (ks_fft_reg.cuh) and exchanges once through a padded shared-memory line;
// between a row pass' inverse and forward transforms the data never leaves registers (the lane-to-point assignment of the
// output distribution equals that of the input distribution).  Global accesses are 64-256 byte contiguous segments.
//
// Reference path replaced: orbithunter/ks/orbits.py transforms :2212-2420, :3472-3579, :3869-4005; eqn :526-555;
// matvec :659-709, :2623-2660; rmatvec :711-755, :1984-2045, :3059-3115; costgrad :757-791; precondition :1034-1089.
#pragma once
#include "ks_common.cuh"
#include "ks_fft_reg.cuh"
#include "ks_pointwise.cuh"

namespace kst {
using ks::static_for;
using kspw::Geom;
using kspw::Tup;
using kspw::gload;

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;

// N = N1 * N2 (Cooley-Tukey, two register stages), T lanes per line.
template <int N>
struct Plan;
template <>
struct Plan<64> {
  static constexpr int N1 = 8, N2 = 8, T = 8;
};
template <>
struct Plan<128> {
  static constexpr int N1 = 16, N2 = 8, T = 8;
};
template <>
struct Plan<256> {
  static constexpr int N1 = 16, N2 = 16, T = 16;
};
KS_HD constexpr bool size_ok(int v) { return v == 64 || v == 128 || v == 256; }

// exp(-2 pi i j / N) at g_tw[N + j], N = 16 .. 256 (filled once per device by tw_init_kernel)
#ifdef KS_HOST_EMUL
// one table for all translation units: the host linker keeps a single copy of the inline member functions that read it
inline double2 g_tw[512];
#else
static __device__ double2 g_tw[512];  // per translation unit (no relocatable device code); each unit fills its own
#endif
static __global__ void tw_init_kernel(int) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < 16 || idx >= 512) return;
  int N = 16;
  while (2 * N <= idx) N *= 2;
  double s, c;
  sincospi(2.0 * (idx - N) / N, &s, &c);
  g_tw[idx] = make_double2(c, -s);
}

template <int N>
struct LineFft {
  using P = Plan<N>;
  static constexpr int N1 = P::N1, N2 = P::N2, T = P::T, C1 = N2 / T, C2 = N1 / T, PT = N / T;
  static constexpr int LS = ((N + N1 + 7) / 8) * 8 + 1;  // line stride (double2): >= N1 * (N2 + 1), = 1 mod 8
  static constexpr int TW = C1 * N1 * T;                 // shared twiddle table entries (double2)
  // register layouts (element index minus the lane's g):
  //   "pre"  slot c * N1 + n1   <-> N2 * n1 + T * c      (input of a transform)
  //   "post" slot c2 * N2 + k2  <-> T * c2 + N1 * k2     (output of a transform)
  KS_HD static constexpr int pre_index(int s) { return N2 * (s % N1) + T * (s / N1); }
  KS_HD static constexpr int post_index(int s) { return T * (s / N2) + N1 * (s % N2); }
  KS_HD static constexpr int pre_slot_of(int off) {
    for (int s = 0; s < PT; ++s)
      if (pre_index(s) == off) return s;
    return -1;
  }
  const double2* tws;  // tws[(c * N1 + k1) * T + g] = W_N^{(g + T c) k1}, forward sign (conflict-free, shared by all lines)
  int g;
  // whole CTA: fill the twiddle table (visible after the next __syncthreads)
  KS_DEVI static void fill_table(double2* tab) {
    for (int i = threadIdx.x; i < TW; i += blockDim.x) {
      const int gg = i % T, ck = i / T, k1 = ck % N1, c = ck / N1;
      tab[i] = g_tw[N + (gg + T * c) * k1];
    }
  }
  KS_DEVI void init(int g_, const double2* tab) {
    g = g_;
    tws = tab + g_;
  }
  // pre layout in, post layout out; X[k] = sum_n x[n] exp(-i SIGN 2 pi k n / N), unnormalised.  `line` is this line's
  // shared-memory buffer (LS double2), used for the one exchange.  All lanes of the warp must call.
  template <int SIGN>
  KS_DEVI void transform(double (&vr)[PT], double (&vi)[PT], double2* line) {
    __syncwarp();
    static_for<0, C1>([&](auto Cc) {
      constexpr int c = decltype(Cc)::value;
      double xr[N1], xi[N1];
#pragma unroll
      for (int n1 = 0; n1 < N1; ++n1) {
        xr[n1] = vr[c * N1 + n1];
        xi[n1] = vi[c * N1 + n1];
      }
      ksfft::fft<N1, SIGN>(xr, xi);
      double2* e = line + g + T * c;
      e[0] = make_double2(xr[0], xi[0]);
#pragma unroll
      for (int k1 = 1; k1 < N1; ++k1) {
        const double2 w = tws[(c * N1 + k1) * T];
        const double wr = w.x, wi = SIGN > 0 ? w.y : -w.y;
        e[k1 * (N2 + 1)] = make_double2(fma(xr[k1], wr, -xi[k1] * wi), fma(xr[k1], wi, xi[k1] * wr));
      }
    });
    __syncwarp();
    static_for<0, C2>([&](auto Cc) {
      constexpr int c2 = decltype(Cc)::value;
      double yr[N2], yi[N2];
      const double2* e = line + (g + T * c2) * (N2 + 1);
#pragma unroll
      for (int n2 = 0; n2 < N2; ++n2) {
        const double2 v = e[n2];
        yr[n2] = v.x;
        yi[n2] = v.y;
      }
      ksfft::fft<N2, SIGN>(yr, yi);
#pragma unroll
      for (int k2 = 0; k2 < N2; ++k2) {
        vr[c2 * N2 + k2] = yr[k2];
        vi[c2 * N2 + k2] = yi[k2];
      }
    });
    __syncwarp();
  }
  // post layout -> pre layout (a compile-time permutation of registers)
  KS_DEVI static void post_to_pre(double (&vr)[PT], double (&vi)[PT]) {
    double tr[PT], ti[PT];
    static_for<0, PT>([&](auto Sc) {
      constexpr int s = decltype(Sc)::value;
      constexpr int d = pre_slot_of(post_index(s));
      static_assert(d >= 0, "distributions must coincide");
      tr[d] = vr[s];
      ti[d] = vi[s];
    });
#pragma unroll
    for (int s = 0; s < PT; ++s) {
      vr[s] = tr[s];
      vi[s] = ti[s];
    }
  }
};

struct TileArgs {
  KsEvalArgs e;  // pointers already offset to the chunk's first orbit; e.batch = orbits in this chunk
  Geom g;
  double2* XC;      // [chunk][k][n]
  double2* FU2;     // [chunk][n/2][m]  field rows (t, t + n/2) as (x, y)
  double* TU;       // [chunk][nm]      stage-1 state when the caller does not want F
  double* partial;  // [chunk][nkb][4]
  int nkb, ntb;     // line blocks per orbit in column / row passes
  int forward;      // row pass: transform the product forward again (0: only the field is wanted)
  int zero_nuu;     // stage 1: the u*u transform was skipped, use P = 0
};

KS_HD constexpr int col_lines(int N) { return kThreads / (N == 256 ? 16 : 8); }
template <int N>
KS_HD constexpr size_t tile_smem_bytes() {
  return ((size_t)(kThreads / LineFft<N>::T) * LineFft<N>::LS + LineFft<N>::TW) * 16 + kWarps * 4 * 8;
}

KS_DEVI void spec_store(double2* ln, int j, int N, const Tup& t) {
  if (j == 0) {
    ln[0] = make_double2(ks::kSqrt2 * t.a1, ks::kSqrt2 * t.b1);
    ln[N / 2] = make_double2(0.0, 0.0);
  } else {
    ln[j] = make_double2(t.a1 - t.b2, t.a2 + t.b1);
    ln[N - j] = make_double2(t.a1 + t.b2, t.b1 - t.a2);
  }
}

// One CTA = one tile (a block of lines of one orbit); the hardware scheduler balances tiles over SMs and co-resident CTAs
// sit in different phases (global loads / FFT / stores), which is what hides the memory latency.  Every phase first issues
// all of its global loads into registers and only then touches shared memory.

// ---- pass A: modes -> XC ----------------------------------------------------------------------------------------------
template <int CLS, int N>
__global__ void __launch_bounds__(kThreads) tile_colinv_kernel(const TileArgs ta) {
  using F = LineFft<N>;
  constexpr int KB = kThreads / F::T, H1 = N / 2, NIT = H1 * KB / kThreads;
  KS_DYN_SMEM(double2, smem);
  double2* tab = smem + KB * F::LS;
  const Geom g = ta.g;
  const int ll = threadIdx.x / F::T, gi = threadIdx.x % F::T;
  const int orbit = blockIdx.x / ta.nkb, k0 = (blockIdx.x - orbit * ta.nkb) * KB;
  const double* Mo = ta.e.modes + (size_t)orbit * g.nm;
  double2* XC = ta.XC + (size_t)orbit * g.k * N;
  F::fill_table(tab);
  F f;
  f.init(gi, tab);
  double2* ln = smem + ll * F::LS;
  {
    Tup t[NIT];
    const int kl = threadIdx.x % KB, j0 = threadIdx.x / KB, kp = k0 + kl;
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      t[i] = Tup{0.0, 0.0, 0.0, 0.0};
      if (kp < g.k) t[i] = gload<CLS>(Mo, j0 + i * (kThreads / KB), kp, g);
    }
#pragma unroll
    for (int i = 0; i < NIT; ++i) spec_store(smem + kl * F::LS, j0 + i * (kThreads / KB), N, t[i]);
  }
  __syncthreads();
  double vr[F::PT], vi[F::PT];
#pragma unroll
  for (int s = 0; s < F::PT; ++s) {
    const double2 v = ln[gi + F::pre_index(s)];
    vr[s] = v.x;
    vi[s] = v.y;
  }
  f.template transform<-1>(vr, vi, ln);
  const int kp = k0 + ll;
  if (kp < g.k) {
    double2* o = XC + (size_t)kp * N + gi;
#pragma unroll
    for (int s = 0; s < F::PT; ++s) o[F::post_index(s)] = make_double2(vr[s] * g.ct, vi[s] * g.ct);
  }
}

// ---- passes B / D: rows ------------------------------------------------------------------------------------------------
enum { MODE_SQ = 0, MODE_MUL = 1 };
template <int M, int MODE>
__global__ void __launch_bounds__(kThreads) tile_row_kernel(const TileArgs ta) {
  using F = LineFft<M>;
  constexpr int TB = kThreads / F::T, K = M / 2 - 1, NIT = (K * TB + kThreads - 1) / kThreads;
  KS_DYN_SMEM(double2, smem);
  double2* tab = smem + TB * F::LS;
  const Geom g = ta.g;
  const int n = g.n, n2 = n / 2;
  const int ll = threadIdx.x / F::T, gi = threadIdx.x % F::T;
  const int orbit = blockIdx.x / ta.ntb, t0 = (blockIdx.x - orbit * ta.ntb) * TB;
  double2* XC = ta.XC + (size_t)orbit * K * n;
  F::fill_table(tab);
  F f;
  f.init(gi, tab);
  double2* ln = smem + ll * F::LS;
  const int tl = threadIdx.x % TB, kq = threadIdx.x / TB;  // this thread's column of the tile, first wavenumber
  {
    double2 A[NIT], B[NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int kp = kq + i * (kThreads / TB);
      if (kp < K) {
        A[i] = XC[(size_t)kp * n + t0 + tl];
        B[i] = XC[(size_t)kp * n + t0 + tl + n2];
      }
    }
    double2* l2 = smem + tl * F::LS;
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int kp = kq + i * (kThreads / TB);
      if (kp < K) {
        l2[kp + 1] = make_double2(A[i].x - B[i].y, A[i].y + B[i].x);
        l2[M - kp - 1] = make_double2(A[i].x + B[i].y, B[i].x - A[i].y);
      }
    }
    if (threadIdx.x < TB) {
      smem[threadIdx.x * F::LS] = make_double2(0.0, 0.0);
      smem[threadIdx.x * F::LS + M / 2] = make_double2(0.0, 0.0);
    }
  }
  double2* fu = ta.FU2 + ((size_t)orbit * n2 + t0 + ll) * M + gi;
  double2 uf[MODE == MODE_MUL ? F::PT : 1];
  if constexpr (MODE == MODE_MUL) {  // the stored field, issued before the inverse transform needs it
#pragma unroll
    for (int s = 0; s < F::PT; ++s) uf[s] = fu[F::post_index(s)];
  }
  __syncthreads();
  double vr[F::PT], vi[F::PT];
#pragma unroll
  for (int s = 0; s < F::PT; ++s) {
    const double2 v = ln[gi + F::pre_index(s)];
    vr[s] = v.x;
    vi[s] = v.y;
  }
  f.template transform<-1>(vr, vi, ln);
#pragma unroll
  for (int s = 0; s < F::PT; ++s) {
    const double ua = vr[s] * g.cx, ub = vi[s] * g.cx;
    if constexpr (MODE == MODE_SQ) {
      fu[F::post_index(s)] = make_double2(ua, ub);
      vr[s] = ua * ua;
      vi[s] = ub * ub;
    } else {
      vr[s] = ua * uf[s].x;
      vi[s] = ub * uf[s].y;
    }
  }
  if (!ta.forward) return;
  F::post_to_pre(vr, vi);
  f.template transform<1>(vr, vi, ln);
#pragma unroll
  for (int s = 0; s < F::PT; ++s) ln[gi + F::post_index(s)] = make_double2(vr[s], vi[s]);
  __syncthreads();
  const double2* l2 = smem + tl * F::LS;
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    const int kp = kq + i * (kThreads / TB);
    if (kp < K) {
      const double2 y = l2[kp + 1], z = l2[M - kp - 1];
      XC[(size_t)kp * n + t0 + tl] = make_double2((y.x + z.x) * g.cx, (y.y - z.y) * g.cx);
      XC[(size_t)kp * n + t0 + tl + n2] = make_double2((y.y + z.y) * g.cx, (z.x - y.x) * g.cx);
    }
  }
}

KS_DEVI double warp_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// fixed-order sum of the per-tile partials of one orbit -> cost and parameter components
template <int CLS, int OP>
KS_DEVI void finalize_orbit(const TileArgs& ta, const kspw::OrbitCtx& c, int orbit) {
  double t[4] = {0.0, 0.0, 0.0, 0.0};
  for (int kb = 0; kb < ta.nkb; ++kb)
    for (int i = 0; i < 4; ++i) t[i] += ta.partial[((size_t)orbit * ta.nkb + kb) * 4 + i];
  kspw::finish_scalars<CLS, OP>(ta.e, ta.g, c, orbit, t[0], t[1], t[2], t[3]);
}

// ---- passes C / E: columns forward, tuple algebra, (C) second spectrum inverse ----------------------------------------
template <int CLS, int OP, int N, int STAGE>
__global__ void __launch_bounds__(kThreads) tile_colfwd_kernel(const TileArgs ta) {
  using F = LineFft<N>;
  constexpr int KB = kThreads / F::T, H1 = N / 2, NIT = H1 * KB / kThreads, JS = kThreads / KB;
  constexpr bool REINV = (STAGE == 1 && OP != KS_OP_EQN);
  KS_DYN_SMEM(double2, smem);
  double2* tab = smem + KB * F::LS;
  double* red = reinterpret_cast<double*>(tab + F::TW);  // [kWarps][4]
  const Geom g = ta.g;
  const KsEvalArgs& a = ta.e;
  const int ll = threadIdx.x / F::T, gi = threadIdx.x % F::T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int orbit = blockIdx.x / ta.nkb, kb = blockIdx.x - orbit * ta.nkb, k0 = kb * KB;
  double2* XC = ta.XC + (size_t)orbit * g.k * N;
  const kspw::OrbitCtx c = kspw::make_ctx<CLS, OP>(a, g, orbit, ta.TU + (size_t)orbit * g.nm);
  F::fill_table(tab);
  F f;
  f.init(gi, tab);
  double2* ln = smem + ll * F::LS;
  double vr[F::PT], vi[F::PT];
  {
    const int kp = k0 + ll;
    const bool live = kp < g.k && !(STAGE == 1 && ta.zero_nuu);
    const double2* src = XC + (size_t)(live ? kp : 0) * N + gi;
#pragma unroll
    for (int s = 0; s < F::PT; ++s) {
      const double2 v = live ? src[F::pre_index(s)] : make_double2(0.0, 0.0);
      vr[s] = v.x;
      vi[s] = v.y;
    }
  }
  // the tuples of u / v / F this tile's algebra needs: in flight during the forward transform
  const int kl = threadIdx.x % KB, j0 = threadIdx.x / KB, kpw = k0 + kl;
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
  __syncthreads();  // twiddle table
  f.template transform<1>(vr, vi, ln);
#pragma unroll
  for (int s = 0; s < F::PT; ++s) ln[gi + F::post_index(s)] = make_double2(vr[s], vi[s]);
  __syncthreads();
  kspw::Sums sm;
  if (act) {
    double2* l2 = smem + kl * F::LS;
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int j = j0 + i * JS;
      Tup P;
      if (j == 0) {
        const double c0 = ks::kSqrt2 * g.ct;
        P = Tup{l2[0].x * c0, 0.0, l2[0].y * c0, 0.0};
      } else {
        const double2 y = l2[j], z = l2[N - j];
        P = Tup{(y.x + z.x) * g.ct, (y.y - z.y) * g.ct, (y.y + z.y) * g.ct, (z.x - y.x) * g.ct};
      }
      if constexpr (STAGE == 1) {
        Tup W{0.0, 0.0, 0.0, 0.0};
        kspw::stage1_point<CLS, OP>(c, g, j, kpw, P, Mt[i], Vt[(OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) ? i : 0], W, sm);
        if constexpr (REINV) spec_store(l2, j, N, W);
      } else {
        kspw::stage2_point<CLS, OP>(c, g, j, kpw, P, Mt[i]);
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
    // the per-tile partial sums of stage 1 are complete (previous launches): fold them here, once per orbit
    if (OP != KS_OP_MATVEC && kb == 0 && threadIdx.x == 0) finalize_orbit<CLS, OP>(ta, c, orbit);
  }
  if constexpr (REINV) {
#pragma unroll
    for (int s = 0; s < F::PT; ++s) {
      const double2 v = ln[gi + F::pre_index(s)];
      vr[s] = v.x;
      vi[s] = v.y;
    }
    f.template transform<-1>(vr, vi, ln);
    const int kp = k0 + ll;
    if (kp < g.k) {
      double2* o = XC + (size_t)kp * N + gi;
#pragma unroll
      for (int s = 0; s < F::PT; ++s) o[F::post_index(s)] = make_double2(vr[s] * g.ct, vi[s] * g.ct);
    }
  }
}

// eqn only (no stage 2): one thread per orbit
template <int CLS, int OP>
__global__ void tile_finalize_kernel(const TileArgs ta) {
  const int orbit = blockIdx.x * blockDim.x + threadIdx.x;
  if (orbit >= ta.e.batch) return;
  const kspw::OrbitCtx c = kspw::make_ctx<CLS, OP>(ta.e, ta.g, orbit, ta.TU + (size_t)orbit * ta.g.nm);
  finalize_orbit<CLS, OP>(ta, c, orbit);
}

// workspace doubles per orbit of a chunk
KS_HD size_t orbit_doubles(int n, int m, int cols) {
  const size_t k = m / 2 - 1;
  const size_t nkb = (k + col_lines(n) - 1) / col_lines(n);
  size_t d = 2 * k * n + (size_t)n * m + (size_t)(n - 1) * cols + nkb * 4;
  return (d + 31) & ~(size_t)31;
}

}  // namespace kst
