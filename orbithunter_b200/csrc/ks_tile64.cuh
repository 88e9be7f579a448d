// 64x64 fused spectral kernel: ONE CTA = ONE orbit, all five passes of ks_tile.cuh inside the CTA.
//
// The complex spatial-mode array XC[kp][t] (31 x 64 double2 = 31 KB) and the field pairs FU2[t][x] (32 x 64 double2 = 32 KB)
// never leave shared memory; with the 32 padded line buffers (36.5 KB) a CTA uses 101 KB, so two CTAs (16 warps) are
// resident per SM.  256 threads = 32 lines x 8 lanes: every pass transforms all of its lines (31 wavenumbers or 32 row
// pairs) at once with the radix-8 x radix-8 register FFT of ks_tile.cuh (LineFft<64>).  Global traffic per evaluation is the
// algorithmic minimum (u's modes in, the result out) plus one L2-resident round trip of the stage-1 state vector.
// This is the BASELINE configs[2] shape (4096 orbits of each symmetry class at 64x64).  Same tuple algebra
// (ks_pointwise.cuh), same reference citations as ks_tile.cuh.
#pragma once
#include "ks_tile.cuh"

namespace kst {

constexpr int kF64Threads = 256;
constexpr int kF64Warps = kF64Threads / 32;
struct F64 {
  static constexpr int N = 64, K = 31, N2 = 32, LINES = 32;
  using F = LineFft<64>;
  static constexpr int kXc = K * N, kFu = N2 * N, kLines = LINES * F::LS, kTab = F::TW;  // double2 counts
  static constexpr size_t smem_bytes = (size_t)(kXc + kFu + kLines + kTab) * 16 + kF64Warps * 4 * 8;
};

template <int CLS, int OP>
__global__ void __launch_bounds__(kF64Threads, 2) tile_fused64_kernel(const TileArgs ta) {
  using F = F64::F;
  constexpr int N = F64::N, K = F64::K, NH = F64::N2, LINES = F64::LINES;
  constexpr int NIT = NH * LINES / kF64Threads;  // tuples per thread in the column passes: 4
  constexpr int JS = kF64Threads / LINES;        // 8
  constexpr int RIT = (K + JS - 1) / JS;         // wavenumbers per thread in the row gathers: 4
  constexpr bool REINV = (OP != KS_OP_EQN);
  KS_DYN_SMEM(double2, smem);
  double2* XC = smem;                  // [K][N]
  double2* FU = XC + F64::kXc;         // [NH][N]  rows (t, t + 32) of the field as (x, y)
  double2* lines = FU + F64::kFu;      // [LINES][LS]
  double2* tab = lines + F64::kLines;  // twiddles
  double* red = reinterpret_cast<double*>(tab + F64::kTab);
  const Geom g = ta.g;
  const KsEvalArgs& a = ta.e;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ll = tid / F::T, gi = tid % F::T;  // FFT phases: line, lane within the line
  const int kl = tid % LINES, j0 = tid / LINES;  // tuple / gather phases: line, first index
  F::fill_table(tab);
  F f;
  f.init(gi, tab);
  double2* ln = lines + ll * F::LS;
  double2* lk = lines + kl * F::LS;
  const bool nuu = kspw::need_nuu<OP>(a);
  __syncthreads();

  auto load_pre = [&](double (&vr)[F::PT], double (&vi)[F::PT], const double2* src) {
#pragma unroll
    for (int s = 0; s < F::PT; ++s) {
      const double2 v = src[gi + F::pre_index(s)];
      vr[s] = v.x;
      vi[s] = v.y;
    }
  };
  // lines (spectra) -> inverse time transform -> XC
  auto col_inverse = [&]() {
    double vr[F::PT], vi[F::PT];
    load_pre(vr, vi, ln);
    f.template transform<-1>(vr, vi, ln);
    if (ll < K) {
      double2* o = XC + ll * N + gi;
#pragma unroll
      for (int s = 0; s < F::PT; ++s) o[F::post_index(s)] = make_double2(vr[s] * g.ct, vi[s] * g.ct);
    }
    __syncthreads();
  };
  // XC -> forward time transform -> lines (natural order)
  auto col_forward = [&](bool zero) {
    double vr[F::PT], vi[F::PT];
    if (ll < K && !zero) {
      load_pre(vr, vi, XC + ll * N);
    } else {
#pragma unroll
      for (int s = 0; s < F::PT; ++s) vr[s] = vi[s] = 0.0;
    }
    f.template transform<1>(vr, vi, ln);
#pragma unroll
    for (int s = 0; s < F::PT; ++s) ln[gi + F::post_index(s)] = make_double2(vr[s], vi[s]);
    __syncthreads();
  };
  // XC -> rows: inverse space transform -> field; MODE_SQ keeps the field and squares it, MODE_MUL multiplies by the kept
  // field; forward space transform -> XC
  auto row_pass = [&](auto mode_c, bool forward) {
    constexpr int MODE = decltype(mode_c)::value;
#pragma unroll
    for (int i = 0; i < RIT; ++i) {
      const int kp = j0 + i * JS;
      if (kp < K) {
        const double2 A = XC[kp * N + kl], B = XC[kp * N + kl + NH];
        lk[kp + 1] = make_double2(A.x - B.y, A.y + B.x);
        lk[N - kp - 1] = make_double2(A.x + B.y, B.x - A.y);
      }
    }
    if (tid < LINES) {
      lines[tid * F::LS] = make_double2(0.0, 0.0);
      lines[tid * F::LS + N / 2] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    double vr[F::PT], vi[F::PT];
    load_pre(vr, vi, ln);
    f.template transform<-1>(vr, vi, ln);
    double2* fu = FU + ll * N + gi;
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
    if (forward) {
      F::post_to_pre(vr, vi);
      f.template transform<1>(vr, vi, ln);
#pragma unroll
      for (int s = 0; s < F::PT; ++s) ln[gi + F::post_index(s)] = make_double2(vr[s], vi[s]);
    }
    __syncthreads();
    if (forward) {
#pragma unroll
      for (int i = 0; i < RIT; ++i) {
        const int kp = j0 + i * JS;
        if (kp < K) {
          const double2 y = lk[kp + 1], z = lk[N - kp - 1];
          XC[kp * N + kl] = make_double2((y.x + z.x) * g.cx, (y.y - z.y) * g.cx);
          XC[kp * N + kl + NH] = make_double2((y.y + z.y) * g.cx, (z.x - y.x) * g.cx);
        }
      }
      __syncthreads();
    }
  };
  auto tuple_of_line = [&](int j) {
    if (j == 0) {
      const double c0 = ks::kSqrt2 * g.ct;
      return Tup{lk[0].x * c0, 0.0, lk[0].y * c0, 0.0};
    }
    const double2 y = lk[j], z = lk[N - j];
    return Tup{(y.x + z.x) * g.ct, (y.y - z.y) * g.ct, (y.y + z.y) * g.ct, (z.x - y.x) * g.ct};
  };

#pragma unroll 1
  for (int orbit = blockIdx.x; orbit < a.batch; orbit += gridDim.x) {
    const kspw::OrbitCtx c = kspw::make_ctx<CLS, OP>(a, g, orbit, ta.TU + (size_t)blockIdx.x * g.nm);
    const bool act = kl < K;
    // operator cache (Krylov drivers): N(u, u) was stored by the eqn evaluation that opened the outer iteration, so the
    // forward transforms of u u (the second half of phase B and the first half of phase C: 2 of the 8 line sweeps) are
    // skipped.  (Caching u's field as well and skipping phases A and B altogether was measured and dropped: reading 32 KB of
    // field per orbit from L2 / HBM is not overlapped with anything in this one-orbit-at-a-time kernel -- 235 us per 4096
    // operator applications against 260 us uncached, profiles/r02oc_*.)
    const bool cached = (OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) && a.n_cache != nullptr;
    // A: u's tuples (kept in registers for stage 1) -> spectra -> XC
    Tup Mt[NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      Mt[i] = Tup{0.0, 0.0, 0.0, 0.0};
      if (act) Mt[i] = gload<CLS>(c.Mo, j0 + i * JS, kl, g);
    }
    Tup Vt[(OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) ? NIT : 1];
    if constexpr (OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) {
#pragma unroll
      for (int i = 0; i < NIT; ++i)
        if (act) Vt[i] = gload<CLS>(c.Vo, j0 + i * JS, kl, g);
    }
    Tup Nt[(OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) ? NIT : 1];
    if constexpr (OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) {
      if (cached) {
        const double* Nc = a.n_cache + (size_t)orbit * g.nm;
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          Nt[i] = Tup{0.0, 0.0, 0.0, 0.0};
          if (act) Nt[i] = gload<CLS>(Nc, j0 + i * JS, kl, g);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NIT; ++i) spec_store(lk, j0 + i * JS, N, Mt[i]);
    __syncthreads();
    col_inverse();
    // B: field of u, u*u (the latter only when its transform is wanted)
    row_pass(std::integral_constant<int, MODE_SQ>{}, nuu && !cached);
    // C: stage 1
    if (!cached) col_forward(!nuu);
    kspw::Sums sm;
    if (act) {
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const int j = j0 + i * JS;
        Tup W{0.0, 0.0, 0.0, 0.0};
        if (cached) {
          if constexpr (OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) kspw::stage1_point_n<CLS, OP>(c, g, j, kl, Nt[i], Mt[i], Vt[i], W, sm);
        } else {
          const Tup Nn = kspw::n_from_p<CLS>(c, g, j, kl, tuple_of_line(j));
          if constexpr (OP == KS_OP_EQN) {
            if (a.n_out) kspw::gstore<CLS>(a.n_out + (size_t)orbit * g.nm, j, kl, g, Nn);  // fills the operator cache
          }
          kspw::stage1_point_n<CLS, OP>(c, g, j, kl, Nn, Mt[i], Vt[(OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) ? i : 0], W, sm);
        }
        if constexpr (REINV) spec_store(lk, j, N, W);
      }
    }
    {
      const double r0 = warp_sum(sm.ff), r1 = warp_sum(sm.e), r2 = warp_sum(fma(-c.sig, sm.d, sm.x)), r3 = warp_sum(sm.d);
      if (lane == 0) {
        red[warp * 4 + 0] = r0;
        red[warp * 4 + 1] = r1;
        red[warp * 4 + 2] = r2;
        red[warp * 4 + 3] = r3;
      }
      __syncthreads();
      if (tid == 0 && OP != KS_OP_MATVEC) {
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        for (int wq = 0; wq < kF64Warps; ++wq)
          for (int i = 0; i < 4; ++i) t[i] += red[wq * 4 + i];
        kspw::finish_scalars<CLS, OP>(a, g, c, orbit, t[0], t[1], t[2], t[3]);
      }
    }
    if constexpr (REINV) {
      col_inverse();
      // D: second field times u
      row_pass(std::integral_constant<int, MODE_MUL>{}, true);
      // E: stage 2 (the stage-1 state vector F / J v's first part comes back from its L2-resident slot)
      col_forward(false);
      if (act) {
        Tup V0[NIT];
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          if constexpr (OP == KS_OP_RMATVEC) V0[i] = Vt[i];
          else V0[i] = kspw::stage2_load<CLS, OP>(c, g, j0 + i * JS, kl);
        }
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
          const int j = j0 + i * JS;
          kspw::stage2_point<CLS, OP>(c, g, j, kl, tuple_of_line(j), V0[i]);
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace kst
