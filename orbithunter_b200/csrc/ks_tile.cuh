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
// plus a one-thread-per-orbit finalize kernel that adds the per-tile partial sums in a fixed order (deterministic cost
// and parameter components).  A length-N line FFT is done by T = 8 or 16 adjacent lanes: each lane holds N/T points in
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
static __device__ double2 g_tw[512];
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
  double twr[C1 * N1], twi[C1 * N1];  // W_N^{n2 k1}, forward sign
  int g;
  KS_DEVI void init(int g_) {
    g = g_;
#pragma unroll
    for (int c = 0; c < C1; ++c)
#pragma unroll
      for (int k1 = 0; k1 < N1; ++k1) {
        const double2 w = g_tw[N + (g_ + T * c) * k1];
        twr[c * N1 + k1] = w.x;
        twi[c * N1 + k1] = w.y;
      }
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
        const double wr = twr[c * N1 + k1], wi = SIGN > 0 ? twi[c * N1 + k1] : -twi[c * N1 + k1];
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
  return (size_t)(kThreads / LineFft<N>::T) * LineFft<N>::LS * 16 + kWarps * 4 * 8;
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

// ---- pass A: modes -> XC ----------------------------------------------------------------------------------------------
template <int CLS, int N>
__global__ void __launch_bounds__(kThreads) tile_colinv_kernel(const TileArgs ta) {
  using F = LineFft<N>;
  constexpr int KB = kThreads / F::T;
  KS_DYN_SMEM(double2, smem);
  const Geom g = ta.g;
  const int ll = threadIdx.x / F::T, gi = threadIdx.x % F::T;
  F f;
  f.init(gi);
  double2* ln = smem + ll * F::LS;
  const int ntiles = ta.e.batch * ta.nkb;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int orbit = tile / ta.nkb, k0 = (tile - orbit * ta.nkb) * KB;
    const double* Mo = ta.e.modes + (size_t)orbit * g.nm;
    double2* XC = ta.XC + (size_t)orbit * g.k * N;
    for (int idx = threadIdx.x; idx < (g.h + 1) * KB; idx += kThreads) {
      const int kl = idx % KB, j = idx / KB, kp = k0 + kl;
      Tup t{0.0, 0.0, 0.0, 0.0};
      if (kp < g.k) t = gload<CLS>(Mo, j, kp, g);
      spec_store(smem + kl * F::LS, j, N, t);
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
    __syncthreads();
  }
}

// ---- passes B / D: rows ------------------------------------------------------------------------------------------------
enum { MODE_SQ = 0, MODE_MUL = 1 };
template <int M, int MODE>
__global__ void __launch_bounds__(kThreads) tile_row_kernel(const TileArgs ta) {
  using F = LineFft<M>;
  constexpr int TB = kThreads / F::T;
  KS_DYN_SMEM(double2, smem);
  const Geom g = ta.g;
  const int n = g.n, n2 = n / 2;
  const int ll = threadIdx.x / F::T, gi = threadIdx.x % F::T;
  F f;
  f.init(gi);
  double2* ln = smem + ll * F::LS;
  const int ntiles = ta.e.batch * ta.ntb;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int orbit = tile / ta.ntb, t0 = (tile - orbit * ta.ntb) * TB;
    double2* XC = ta.XC + (size_t)orbit * g.k * n;
    for (int idx = threadIdx.x; idx < g.k * TB; idx += kThreads) {
      const int tl = idx % TB, kp = idx / TB;
      const double2 A = XC[(size_t)kp * n + t0 + tl], B = XC[(size_t)kp * n + t0 + tl + n2];
      double2* l2 = smem + tl * F::LS;
      l2[kp + 1] = make_double2(A.x - B.y, A.y + B.x);
      l2[M - kp - 1] = make_double2(A.x + B.y, B.x - A.y);
    }
    if (threadIdx.x < TB) {
      smem[threadIdx.x * F::LS] = make_double2(0.0, 0.0);
      smem[threadIdx.x * F::LS + M / 2] = make_double2(0.0, 0.0);
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
    double2* fu = ta.FU2 + ((size_t)orbit * n2 + t0 + ll) * M + gi;
#pragma unroll
    for (int s = 0; s < F::PT; ++s) {
      const double ua = vr[s] * g.cx, ub = vi[s] * g.cx;
      if constexpr (MODE == MODE_SQ) {
        fu[F::post_index(s)] = make_double2(ua, ub);
        vr[s] = ua * ua;
        vi[s] = ub * ub;
      } else {
        const double2 u = fu[F::post_index(s)];
        vr[s] = ua * u.x;
        vi[s] = ub * u.y;
      }
    }
    if (ta.forward) {
      F::post_to_pre(vr, vi);
      f.template transform<1>(vr, vi, ln);
#pragma unroll
      for (int s = 0; s < F::PT; ++s) ln[gi + F::post_index(s)] = make_double2(vr[s], vi[s]);
      __syncthreads();
      for (int idx = threadIdx.x; idx < g.k * TB; idx += kThreads) {
        const int tl = idx % TB, kp = idx / TB;
        const double2* l2 = smem + tl * F::LS;
        const double2 y = l2[kp + 1], z = l2[M - kp - 1];
        XC[(size_t)kp * n + t0 + tl] = make_double2((y.x + z.x) * g.cx, (y.y - z.y) * g.cx);
        XC[(size_t)kp * n + t0 + tl + n2] = make_double2((y.y + z.y) * g.cx, (z.x - y.x) * g.cx);
      }
    }
    __syncthreads();
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

// ---- passes C / E: columns forward, tuple algebra, (C) second spectrum inverse ----------------------------------------
template <int CLS, int OP, int N, int STAGE>
__global__ void __launch_bounds__(kThreads) tile_colfwd_kernel(const TileArgs ta) {
  using F = LineFft<N>;
  constexpr int KB = kThreads / F::T;
  constexpr bool REINV = (STAGE == 1 && OP != KS_OP_EQN);
  KS_DYN_SMEM(double2, smem);
  double* red = reinterpret_cast<double*>(smem + KB * F::LS);  // [kWarps][4]
  const Geom g = ta.g;
  const KsEvalArgs& a = ta.e;
  const int ll = threadIdx.x / F::T, gi = threadIdx.x % F::T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  F f;
  f.init(gi);
  double2* ln = smem + ll * F::LS;
  const int ntiles = a.batch * ta.nkb;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int orbit = tile / ta.nkb, kb = tile - orbit * ta.nkb, k0 = kb * KB;
    double2* XC = ta.XC + (size_t)orbit * g.k * N;
    const kspw::OrbitCtx c = kspw::make_ctx<CLS, OP>(a, g, orbit, ta.TU + (size_t)orbit * g.nm);
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
    f.template transform<1>(vr, vi, ln);
#pragma unroll
    for (int s = 0; s < F::PT; ++s) ln[gi + F::post_index(s)] = make_double2(vr[s], vi[s]);
    __syncthreads();
    kspw::Sums sm;
    for (int idx = threadIdx.x; idx < (g.h + 1) * KB; idx += kThreads) {
      const int kl = idx % KB, j = idx / KB, kp = k0 + kl;
      if (kp >= g.k) continue;
      double2* l2 = smem + kl * F::LS;
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
        kspw::stage1_point<CLS, OP>(c, g, j, kp, P, W, sm);
        if constexpr (REINV) spec_store(l2, j, N, W);
      } else {
        kspw::stage2_point<CLS, OP>(c, g, j, kp, P);
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
    }
    __syncthreads();
    if constexpr (STAGE == 1) {
      if (threadIdx.x < 4) {
        double t = 0.0;
        for (int wq = 0; wq < kWarps; ++wq) t += red[wq * 4 + threadIdx.x];
        ta.partial[((size_t)orbit * ta.nkb + kb) * 4 + threadIdx.x] = t;
      }
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
    __syncthreads();
  }
}

// one thread per orbit: fixed-order sum of the per-tile partials -> cost and parameter components
template <int CLS, int OP>
__global__ void tile_finalize_kernel(const TileArgs ta) {
  const int orbit = blockIdx.x * blockDim.x + threadIdx.x;
  if (orbit >= ta.e.batch) return;
  const Geom g = ta.g;
  const kspw::OrbitCtx c = kspw::make_ctx<CLS, OP>(ta.e, g, orbit, ta.TU + (size_t)orbit * g.nm);
  double t[4] = {0.0, 0.0, 0.0, 0.0};
  for (int kb = 0; kb < ta.nkb; ++kb)
    for (int i = 0; i < 4; ++i) t[i] += ta.partial[((size_t)orbit * ta.nkb + kb) * 4 + i];
  kspw::finish_scalars<CLS, OP>(ta.e, g, c, orbit, t[0], t[1], t[2], t[3]);
}

// workspace doubles per orbit of a chunk
KS_HD size_t orbit_doubles(int n, int m, int cols) {
  const size_t k = m / 2 - 1;
  const size_t nkb = (k + col_lines(n) - 1) / col_lines(n);
  size_t d = 2 * k * n + (size_t)n * m + (size_t)(n - 1) * cols + nkb * 4;
  return (d + 31) & ~(size_t)31;
}

}  // namespace kst
