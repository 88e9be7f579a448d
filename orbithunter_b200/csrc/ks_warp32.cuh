// 32x32 fused spectral kernels: ONE WARP evaluates TWO orbits, all transforms in registers.
//
// Reference path replaced (orbithunter/ks/orbits.py): eqn :526-555, matvec :659-709 (Relative :2623-2660),
// rmatvec :711-755 with _rmatvec_parameters :1984-2024 (Relative :3059-3094), costgrad :757-791,
// precondition :1034-1089, the transforms :2212-2420 and the SR/Anti projections :3472-3579, :3869-4005.
//
// Data layout.  The reference stores modes as a real (n-1, cols) array: rows [Re_t j=0..15 ; Im_t j=1..15],
// columns [Re_x k=1..15 | Im_x k=1..15] (15 columns for SR/Anti).  For a fixed spatial wavenumber k the four reals
//   (a1, a2, b1, b2) = (M[j,k'], M[15+j,k'], M[j,15+k'], M[15+j,15+k'])
// are one "tuple": C_A[j] = (a1 + i a2)/sqrt2, C_B[j] = (b1 + i b2)/sqrt2 are the time-rfft coefficients of the real
// and imaginary parts of the spatial mode X[t,k] = A + iB.  The complex time spectrum of X is
//   Y[j] = C_A[j] + i C_B[j],   Y[32-j] = conj(C_A[j]) + i conj(C_B[j]),   Y[16] = 0 (Nyquist dropped).
// A 2-D transform is therefore: 15 complex length-32 FFTs over t (one per wavenumber, "COL" phase, lane = wavenumber)
// and 32 real length-32 FFTs over x done as 16 complex FFTs by the two-for-one trick ("ROW" phase, lane = rows
// (t, t+16)).  Two orbits share a warp so that both phases keep 30-32 lanes busy.  The two phases exchange the
// spatial modes X[t,k] (480 complex per orbit) through a per-warp shared-memory tile; only __syncwarp is needed.
// All normalisations are folded into one constant per product (n = m = 32 makes them powers of two).
//
// Per warp: XB tile 2 x 480 double2 (15 KB) + field stash 2 x 1024 doubles (16 KB) => 7 resident warps / SM.
// The kernel is persistent: each warp strides over orbit pairs.
#pragma once
#include "ks_common.cuh"
#include "ks_fft_reg.cuh"

namespace ksw32 {
using ks::static_for;

// Two build variants of the kernel (template parameter TS = "tuple stash in shared memory"):
//   TS = false: XB + field stash per warp (31.7 KB) -> 7 warps / SM; the stage-1 state vector (F) round-trips
//               through an L2-resident global scratch slot and the u modes are re-read from global in phase C.
//   TS = true : + a lane-private tuple stash (16 KB) holding u's tuples from phase A to C and F from C to E
//               -> 4 warps / SM, but no global/L2 round trips inside an evaluation.
// The library instantiates TS = false with AS = true (all global reads staged with cp.async one transform ahead), for
// matvec / rmatvec and as the cross-check of ks_warp32x1.cuh; the other combinations are kept for the measurements in
// DESIGN.md section 4 (profiles/r01b, r01e).
constexpr int kXbDoubles = 2 * 480 * 2, kSuDoubles = 2 * 1024, kTsDoubles = 2 * 1024;
template <bool TS>
struct Cfg {
  static constexpr int warp_smem_doubles = kXbDoubles + kSuDoubles + (TS ? kTsDoubles : 0);
  static constexpr int warps_per_cta = TS ? 4 : 7;
  static constexpr int smem_bytes = warps_per_cta * warp_smem_doubles * 8;
};
constexpr int kMaxWarpsPerCta = 7;

template <int CLS>
struct Lay {
  static constexpr bool full = (CLS == KS_ORBIT || CLS == KS_RELATIVE);
  static constexpr int cols = full ? 30 : 15;
  static constexpr int nm = 31 * cols;
  // which of the (a = Re_x, b = Im_x) halves survive the class projection at time index j
  KS_HD static constexpr bool has_a(int j) { return full || (CLS == KS_SHIFT_REFLECTION && (j & 1)); }
  KS_HD static constexpr bool has_b(int j) { return full || CLS == KS_ANTISYMMETRIC || (CLS == KS_SHIFT_REFLECTION && !(j & 1)); }
};

struct Tup {
  double a1, a2, b1, b2;
};

// p points at (orbit base + k'), k' = wavenumber - 1
template <int CLS, int J>
KS_DEVI Tup load_tup(const double* __restrict__ p) {
  using L = Lay<CLS>;
  Tup t{0.0, 0.0, 0.0, 0.0};
  if constexpr (L::full) {
    t.a1 = p[J * 30];
    t.b1 = p[J * 30 + 15];
    if constexpr (J > 0) {
      t.a2 = p[(15 + J) * 30];
      t.b2 = p[(15 + J) * 30 + 15];
    }
  } else if constexpr (L::has_a(J)) {
    t.a1 = p[J * 15];
    if constexpr (J > 0) t.a2 = p[(15 + J) * 15];
  } else {
    t.b1 = p[J * 15];
    if constexpr (J > 0) t.b2 = p[(15 + J) * 15];
  }
  return t;
}
template <int CLS, int J>
KS_DEVI void store_tup(double* __restrict__ p, const Tup& t) {
  using L = Lay<CLS>;
  if constexpr (L::full) {
    p[J * 30] = t.a1;
    p[J * 30 + 15] = t.b1;
    if constexpr (J > 0) {
      p[(15 + J) * 30] = t.a2;
      p[(15 + J) * 30 + 15] = t.b2;
    }
  } else if constexpr (L::has_a(J)) {
    p[J * 15] = t.a1;
    if constexpr (J > 0) p[(15 + J) * 15] = t.a2;
  } else {
    p[J * 15] = t.b1;
    if constexpr (J > 0) p[(15 + J) * 15] = t.b2;
  }
}
template <int CLS, int J>
KS_DEVI void project(Tup& t) {
  if constexpr (!Lay<CLS>::has_a(J)) t.a1 = t.a2 = 0.0;
  if constexpr (!Lay<CLS>::has_b(J)) t.b1 = t.b2 = 0.0;
  if constexpr (J == 0) t.a2 = t.b2 = 0.0;
}

// tuple -> entries j and 32-j of the (sqrt2-scaled) complex time spectrum fed to the inverse transform
template <int J>
KS_DEVI void tup_to_spec(const Tup& t, double (&yr)[32], double (&yi)[32]) {
  if constexpr (J == 0) {
    yr[0] = ks::kSqrt2 * t.a1;
    yi[0] = ks::kSqrt2 * t.b1;
    yr[16] = 0.0;
    yi[16] = 0.0;
  } else {
    yr[J] = t.a1 - t.b2;
    yi[J] = t.a2 + t.b1;
    yr[32 - J] = t.a1 + t.b2;
    yi[32 - J] = t.b1 - t.a2;
  }
}
// forward complex time spectrum -> unscaled tuple (true tuple = result / (64*gamma), extra sqrt2 for J = 0)
template <int J>
KS_DEVI Tup spec_to_tup(const double (&yr)[32], const double (&yi)[32]) {
  Tup t;
  if constexpr (J == 0) {
    t.a1 = yr[0];
    t.b1 = yi[0];
    t.a2 = 0.0;
    t.b2 = 0.0;
  } else {
    t.a1 = yr[J] + yr[32 - J];
    t.a2 = yi[J] - yi[32 - J];
    t.b1 = yi[J] + yi[32 - J];
    t.b2 = yr[32 - J] - yr[J];
  }
  return t;
}

// lane-private tuple stash: slot (J, half) of this lane lives at ts[(2*J + half) * 32 + lane]  (conflict-free LDS/STS.128)
template <int J>
KS_DEVI void ts_store(double2* __restrict__ ts, const Tup& t) {
  ts[(2 * J) * 32] = make_double2(t.a1, t.a2);
  ts[(2 * J + 1) * 32] = make_double2(t.b1, t.b2);
}
template <int J>
KS_DEVI Tup ts_load(const double2* __restrict__ ts) {
  const double2 a = ts[(2 * J) * 32], b = ts[(2 * J + 1) * 32];
  return Tup{a.x, a.y, b.x, b.y};
}

// d/dt and d/dx acting on a tuple (reference dt :423-430, dx :499-516; frequency tables :5239-5307)
KS_DEVI Tup dt_tup(const Tup& t, double w) { return Tup{-w * t.a2, w * t.a1, -w * t.b2, w * t.b1}; }
KS_DEVI Tup dx_tup(const Tup& t, double q) { return Tup{-q * t.b1, -q * t.b2, q * t.a1, q * t.a2}; }
KS_DEVI double dot_tup(const Tup& x, const Tup& y) { return fma(x.a1, y.a1, fma(x.a2, y.a2, fma(x.b1, y.b1, x.b2 * y.b2))); }
// <dt-rotation of m, f> / w  and  <dx-rotation of m, f> / q
KS_DEVI double rot_t(const Tup& m, const Tup& f) { return fma(m.a1, f.a2, fma(-m.a2, f.a1, fma(m.b1, f.b2, -m.b2 * f.b1))); }
KS_DEVI double rot_x(const Tup& m, const Tup& f) { return fma(m.a1, f.b1, fma(m.a2, f.b2, fma(-m.b1, f.a1, -m.b2 * f.a2))); }

KS_DEVI double half_reduce(double v) {  // sum over the 16 lanes that share an orbit
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// ---- asynchronous global -> shared staging (cp.async / LDGSTS): latency of the u-mode and stage-1 reloads is hidden
// behind a transform instead of being paid tuple by tuple.  Destinations are lane-private, so no barrier is needed.
KS_DEVI void cp_async8(double* smem_dst, const double* gsrc) {
#ifdef KS_HOST_EMUL
  *smem_dst = *gsrc;
#else
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
#endif
}
KS_DEVI void cp_async_commit() {
#ifndef KS_HOST_EMUL
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
KS_DEVI void cp_async_wait() {
#ifndef KS_HOST_EMUL
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}
// slot numbering of a lane's 62 staged values: tuple J occupies slots base(J) .. base(J)+3 (J = 0: 2 slots)
KS_HD constexpr int slot_base(int J) { return J == 0 ? 0 : 4 * J - 2; }
// own-XB-column staging: slot s lives in component (s & 1) of the double2 xcol[(s >> 1) * 15]
template <int CLS, int J>
KS_DEVI void stage_tup_col(const double* __restrict__ p, double2* xcol) {
  using L = Lay<CLS>;
  double* d = reinterpret_cast<double*>(xcol + (slot_base(J) >> 1) * 15);  // slot_base is even
  if constexpr (L::full) {
    if constexpr (J == 0) {
      cp_async8(d, p);
      cp_async8(d + 1, p + 15);
    } else {
      cp_async8(d, p + J * 30);
      cp_async8(d + 1, p + (15 + J) * 30);
      cp_async8(d + 30, p + J * 30 + 15);
      cp_async8(d + 31, p + (15 + J) * 30 + 15);
    }
  } else {
    cp_async8(d, p + J * 15);
    if constexpr (J > 0) cp_async8(d + 1, p + (15 + J) * 15);
  }
}
template <int CLS, int J>
KS_DEVI Tup unstage_tup_col(const double2* xcol) {
  using L = Lay<CLS>;
  const double2 lo = xcol[(slot_base(J) >> 1) * 15];
  if constexpr (L::full) {
    if constexpr (J == 0) return Tup{lo.x, 0.0, lo.y, 0.0};
    const double2 hi = xcol[((slot_base(J) >> 1) + 1) * 15];
    return Tup{lo.x, lo.y, hi.x, hi.y};
  } else if constexpr (L::has_a(J)) {
    return Tup{lo.x, J > 0 ? lo.y : 0.0, 0.0, 0.0};
  } else {
    return Tup{0.0, 0.0, lo.x, J > 0 ? lo.y : 0.0};
  }
}
// field-stash staging (the u stash is dead between phase D and the next phase B): slot s at su[s * 32]
template <int CLS, int J>
KS_DEVI void stage_tup_su(const double* __restrict__ p, double* su) {
  using L = Lay<CLS>;
  double* d = su + slot_base(J) * 32;
  if constexpr (L::full) {
    if constexpr (J == 0) {
      cp_async8(d, p);
      cp_async8(d + 32, p + 15);
    } else {
      cp_async8(d, p + J * 30);
      cp_async8(d + 32, p + (15 + J) * 30);
      cp_async8(d + 64, p + J * 30 + 15);
      cp_async8(d + 96, p + (15 + J) * 30 + 15);
    }
  } else {
    cp_async8(d, p + J * 15);
    if constexpr (J > 0) cp_async8(d + 32, p + (15 + J) * 15);
  }
}
template <int CLS, int J>
KS_DEVI Tup unstage_tup_su(const double* su) {
  using L = Lay<CLS>;
  const double* d = su + slot_base(J) * 32;
  if constexpr (L::full) {
    if constexpr (J == 0) return Tup{d[0], 0.0, d[32], 0.0};
    return Tup{d[0], d[32], d[64], d[96]};
  } else if constexpr (L::has_a(J)) {
    return Tup{d[0], J > 0 ? d[32] : 0.0, 0.0, 0.0};
  } else {
    return Tup{0.0, 0.0, d[0], J > 0 ? d[32] : 0.0};
  }
}

// ---- phase helpers --------------------------------------------------------------------------------------------
// COL: inverse time transform of the spectrum in (yr, yi); column k' of the XB tile <- X[t], t = 0..31
KS_DEVI void col_inverse_store(double (&yr)[32], double (&yi)[32], double2* __restrict__ xcol, bool act) {
  ksfft::fft<32, -1>(yr, yi);
  if (act) {
#pragma unroll
    for (int t = 0; t < 32; ++t) xcol[t * 15] = make_double2(yr[t], yi[t]);
  }
}
KS_DEVI void col_load(const double2* __restrict__ xcol, double (&yr)[32], double (&yi)[32]) {
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    const double2 v = xcol[t * 15];
    yr[t] = v.x;
    yi[t] = v.y;
  }
}
KS_DEVI void col_load_forward(const double2* __restrict__ xcol, double (&yr)[32], double (&yi)[32]) {
  col_load(xcol, yr, yi);
  ksfft::fft<32, 1>(yr, yi);
}
// ROW: rows (ta, tb) of XB -> 64 * (u[ta, x] + i u[tb, x]) by one complex inverse transform (two-for-one)
KS_DEVI void row_load_inverse(const double2* __restrict__ xa, const double2* __restrict__ xb, double (&zr)[32],
                              double (&zi)[32]) {
  zr[0] = zi[0] = zr[16] = zi[16] = 0.0;
#pragma unroll
  for (int k = 1; k < 16; ++k) {
    const double2 A = xa[k - 1], B = xb[k - 1];
    zr[k] = A.x - B.y;
    zi[k] = A.y + B.x;
    zr[32 - k] = A.x + B.y;
    zi[32 - k] = B.x - A.y;
  }
  ksfft::fft<32, -1>(zr, zi);
}
// ROW: forward transform of (row ta) + i (row tb), split into the two rows' spatial modes (x2), store to XB rows
KS_DEVI void row_forward_store(double (&zr)[32], double (&zi)[32], double2* __restrict__ xa, double2* __restrict__ xb) {
  ksfft::fft<32, 1>(zr, zi);
#pragma unroll
  for (int k = 1; k < 16; ++k) {
    xa[k - 1] = make_double2(zr[k] + zr[32 - k], zi[k] - zi[32 - k]);
    xb[k - 1] = make_double2(zi[k] + zi[32 - k], zr[32 - k] - zr[k]);
  }
}

// ---- the kernel -----------------------------------------------------------------------------------------------
template <int CLS, int OP, bool TS, bool AS = false, bool PC = false, bool PC_RUNTIME = true>
__global__ void __launch_bounds__(Cfg<TS>::warps_per_cta * 32, 1) ks_warp32_kernel(const KsEvalArgs a) {
  static_assert(!(TS && AS), "async staging is the no-tuple-stash variant");
  constexpr int kWarpSmemDoubles = Cfg<TS>::warp_smem_doubles, kWarpsPerCta = Cfg<TS>::warps_per_cta;
  using L = Lay<CLS>;
  constexpr int NM = L::nm;
  constexpr bool REL = (CLS == KS_RELATIVE);
  KS_DYN_SMEM(double, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = lane >> 4, l16 = lane & 15;
  const bool colact = l16 < 15;
  const int kk = colact ? l16 : 14;  // wavenumber index k' = k - 1 of this lane in COL phases
  double* wsm = smem + warp * kWarpSmemDoubles;
  double2* xb = reinterpret_cast<double2*>(wsm);
  double* su = wsm + kXbDoubles;
  double2* ts = reinterpret_cast<double2*>(wsm + kXbDoubles + kSuDoubles) + lane;  // only dereferenced when TS
  double2* xcol = xb + o * 480 + kk;
  double2* xrow_a = xb + o * 480 + l16 * 15;
  double2* xrow_b = xb + o * 480 + (l16 + 16) * 15;
  // warp-major numbering of the persistent warps: consecutive orbits go to different SMs, so the last (partial) round of
  // a batch is spread over all SMs with few warps each instead of filling some SMs and leaving the others idle
  const int gw = warp * gridDim.x + blockIdx.x, nw = gridDim.x * kWarpsPerCta;
  const int npairs = (a.batch + 1) >> 1;
  const bool fix_t = a.cmask & KS_FIX_T, fix_x = a.cmask & KS_FIX_X, fix_s = (a.cmask & KS_FIX_S) || !REL;
  // does this op need N(u,u) = 1/2 Dx FFT2(u^2)?  eqn/costgrad always; matvec/rmatvec only for the dF/dL column.
  const bool need_nuu = (OP == KS_OP_EQN || OP == KS_OP_COSTGRAD) || !fix_x;

  // programmatic dependent launch: the set-up above overlapped the previous kernel's tail; no global access before this
  ks_pdl_launch_dependents();
  ks_pdl_wait();
  if constexpr (AS) {  // stage the first pair's u modes into the (still unused) field stash
    if (gw < npairs) {
      const int ob0 = (2 * gw + o < a.batch) ? 2 * gw + o : a.batch - 1;
      const double* M0 = a.modes + (size_t)ob0 * NM + kk;
      static_for<0, 16>([&](auto Jc) { stage_tup_su<CLS, decltype(Jc)::value>(M0, su + lane); });
    }
    cp_async_commit();
  }
#pragma unroll 1
  for (int pair = gw; pair < npairs; pair += nw) {
    const int orbit = 2 * pair + o;
    const bool valid = orbit < a.batch;
    const int ob = valid ? orbit : a.batch - 1;
    const bool st = valid && colact;  // this lane owns a column of a real orbit
    const double* Mp = a.modes + (size_t)ob * NM + kk;
    const double T = a.params[ob * 3 + 0], Lx = a.params[ob * 3 + 1], S = a.params[ob * 3 + 2];
    // frequency tables, same expression order as the reference (:5267, :5303)
    const double wbase = -1.0 * (ks::kTwoPi * 32.0 / T);
    const double q = (ks::kTwoPi * 32.0 / Lx) * ((kk + 1) * (1.0 / 32.0));
    const double q2 = q * q, q4 = q2 * q2, d = q4 - q2;
    const double sig = REL ? S / T : 0.0;
    // scale constants: field arrays hold 64*u, products 4096*u*v, forward output 64*(...)
    constexpr double cN = 0.5 / (64.0 * 4096.0);   // 1/2 (u^2)_x
    constexpr double cP = 1.0 / (64.0 * 4096.0);   // u*v, u*w products

    // ---------------- phase A (COL): u modes -> X_u ----------------
    {
      double yr[32], yi[32];
      if constexpr (AS) cp_async_wait<0>();
      static_for<0, 16>([&](auto Jc) {
        constexpr int J = decltype(Jc)::value;
        Tup t;
        if constexpr (AS) t = unstage_tup_su<CLS, J>(su + lane); else t = load_tup<CLS, J>(Mp);
        if constexpr (TS) ts_store<J>(ts, t);
        tup_to_spec<J>(t, yr, yi);
      });
      col_inverse_store(yr, yi, xcol, colact);
    }
    __syncwarp();
    // ---------------- phase B (ROW): field u (x64) -> stash; u^2 -> X_s ----------------
    {
      double zr[32], zi[32];
      row_load_inverse(xrow_a, xrow_b, zr, zi);
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        su[x * 32 + lane] = zr[x];
        su[(32 + x) * 32 + lane] = zi[x];
      }
      if (need_nuu) {
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          zr[x] *= zr[x];
          zi[x] *= zi[x];
        }
        row_forward_store(zr, zi, xrow_a, xrow_b);
      }
    }
    __syncwarp();
    // ---------------- phase C (COL): tuple stage 1, then second inverse transform ----------------
    const int v_ld = a.v_ld ? a.v_ld : NM, out_ld = a.out_ld ? a.out_ld : NM;
    double* Fp = nullptr;  // where the stage-1 state vector (F or partial matvec) lives until stage 2
    if constexpr (OP == KS_OP_EQN) Fp = a.out_state + (size_t)ob * out_ld + kk;
    if constexpr (OP == KS_OP_COSTGRAD)
      Fp = (a.f_out ? a.f_out + (size_t)ob * NM : a.scratch + (size_t)(2 * gw + o) * NM) + kk;
    if constexpr (OP == KS_OP_MATVEC) Fp = a.out_state + (size_t)ob * out_ld + kk;
    const double* Vp = (OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) ? a.vmodes + (size_t)ob * v_ld + kk : nullptr;
    {
      double yr[32], yi[32];
      if (need_nuu) {
        col_load(xcol, yr, yi);
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) yr[t] = yi[t] = 0.0;
      }
      if constexpr (AS) {  // the column is in registers now: its slots stage u's tuples while the transform runs
        // lanes 15/31 mirror a neighbour's column: they must not write into it (their results are discarded anyway)
        if (colact) static_for<0, 16>([&](auto Jc) { stage_tup_col<CLS, decltype(Jc)::value>(Mp, xcol); });
        cp_async_commit();
      }
      if (need_nuu) ksfft::fft<32, 1>(yr, yi);
      if constexpr (AS) cp_async_wait<0>();
      double s_ff = 0.0, s_mf = 0.0, s_x = 0.0, s_d = 0.0, s_nv = 0.0;
      double vt = 0.0, vx = 0.0, vs = 0.0;
      if constexpr (OP == KS_OP_MATVEC) {
        double v3[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int slot = ks::free_slot(a.cmask, i);
          v3[i] = a.v_tail ? (slot < 0 ? 0.0 : a.vmodes[(size_t)ob * v_ld + NM + slot]) : a.vparams[ob * 3 + i];
        }
        vt = fix_t ? 0.0 : v3[0] * (-1.0 / T);
        vx = fix_x ? 0.0 : v3[1] * (1.0 / Lx);
        vs = fix_s ? 0.0 : v3[2] * (-1.0 / T);
      }
      static_for<0, 16>([&](auto Jc) {
        constexpr int J = decltype(Jc)::value;
        const double w = (J == 0) ? 0.0 : wbase * (J * (1.0 / 32.0));
        const Tup P = spec_to_tup<J>(yr, yi);
        const double qn = (J == 0 ? ks::kSqrt2 * cN : cN) * q;
        Tup N{-qn * P.b1, -qn * P.b2, qn * P.a1, qn * P.a2};
        project<CLS, J>(N);
        Tup Mt;
        if constexpr (TS) Mt = ts_load<J>(ts);
        else if constexpr (AS) Mt = unstage_tup_col<CLS, J>(xcol);
        else Mt = load_tup<CLS, J>(Mp);
        if constexpr (OP == KS_OP_EQN || OP == KS_OP_COSTGRAD) {
          // F = u_t + u_xx + u_xxxx [- (S/T) u_x] + 1/2 (u^2)_x        (:1918-1922, :3052-3057, :1950)
          Tup F;
          F.a1 = fma(d, Mt.a1, fma(-w, Mt.a2, N.a1));
          F.a2 = fma(d, Mt.a2, fma(w, Mt.a1, N.a2));
          F.b1 = fma(d, Mt.b1, fma(-w, Mt.b2, N.b1));
          F.b2 = fma(d, Mt.b2, fma(w, Mt.b1, N.b2));
          if constexpr (REL) {
            const double sq = sig * q;
            F.a1 = fma(sq, Mt.b1, F.a1);
            F.a2 = fma(sq, Mt.b2, F.a2);
            F.b1 = fma(-sq, Mt.a1, F.b1);
            F.b2 = fma(-sq, Mt.a2, F.b2);
          }
          project<CLS, J>(F);
          s_ff += dot_tup(F, F);
          if constexpr (TS && OP == KS_OP_COSTGRAD) {
            ts_store<J>(ts, F);
            if (st && a.f_out) store_tup<CLS, J>(Fp, F);
          } else {
            if (st) store_tup<CLS, J>(Fp, F);
          }
          if constexpr (OP == KS_OP_COSTGRAD) {
            s_mf += dot_tup(Mt, F);
            s_x = fma(w, rot_t(Mt, F), s_x);
            if constexpr (REL) s_d += rot_x(Mt, F);
            tup_to_spec<J>(dx_tup(F, q), yr, yi);  // W = Dx F feeds the adjoint nonlinear term (:1977)
          }
        } else if constexpr (OP == KS_OP_RMATVEC) {
          const Tup V = load_tup<CLS, J>(Vp);
          s_mf += dot_tup(Mt, V);
          s_x = fma(w, rot_t(Mt, V), s_x);
          s_nv += dot_tup(N, V);
          if constexpr (REL) s_d += rot_x(Mt, V);
          tup_to_spec<J>(dx_tup(V, q), yr, yi);
        } else {  // MATVEC: everything except the u*v product                      (:690-707, :2648-2658)
          const Tup V = load_tup<CLS, J>(Vp);
          const double e = 2.0 * q2 - 4.0 * q4;
          // coefficient of dx-rotation(M): vt*(-1/T)*(-S/T) + vx*(-1/L)*(-S/T) + vs*(-1/T); vt,vx,vs pre-scaled
          const double cdx = REL ? (-sig * vt + sig * vx + vs) * q : 0.0;
          const double cdt = vt * w;
          Tup R;
          R.a1 = fma(d, V.a1, -w * V.a2);
          R.a2 = fma(d, V.a2, w * V.a1);
          R.b1 = fma(d, V.b1, -w * V.b2);
          R.b2 = fma(d, V.b2, w * V.b1);
          if constexpr (REL) {
            const double sq = sig * q;
            R.a1 = fma(sq, V.b1, R.a1);
            R.a2 = fma(sq, V.b2, R.a2);
            R.b1 = fma(-sq, V.a1, R.b1);
            R.b2 = fma(-sq, V.a2, R.b2);
          }
          const double ce = vx * e;
          R.a1 += fma(ce, Mt.a1, fma(-cdt, Mt.a2, fma(-cdx, Mt.b1, -vx * N.a1)));
          R.a2 += fma(ce, Mt.a2, fma(cdt, Mt.a1, fma(-cdx, Mt.b2, -vx * N.a2)));
          R.b1 += fma(ce, Mt.b1, fma(-cdt, Mt.b2, fma(cdx, Mt.a1, -vx * N.b1)));
          R.b2 += fma(ce, Mt.b2, fma(cdt, Mt.b1, fma(cdx, Mt.a2, -vx * N.b2)));
          project<CLS, J>(R);
          if constexpr (TS) ts_store<J>(ts, R); else { if (st) store_tup<CLS, J>(Fp, R); }
          tup_to_spec<J>(V, yr, yi);
        }
      });
      // per-orbit reductions: cost and the parameter components of J^T v                (:2004-2020, :3064-3092)
      if constexpr (OP == KS_OP_EQN || OP == KS_OP_COSTGRAD) {
        if (!colact) s_ff = 0.0;
        const double ff = half_reduce(s_ff);
        if (valid && l16 == 0 && a.out_cost) a.out_cost[orbit] = 0.5 * ff;
      }
      if constexpr (OP == KS_OP_COSTGRAD || OP == KS_OP_RMATVEC) {
        // r_t = -1/T <dt u - sig dx u, v>;  r_x = 1/L <(2q^2 - 4q^4) u - N + sig dx u, v>;  r_s = -1/T <dx u, v>
        // for v = F the identity N = F - lin(u) gives <N, F> = <F,F> - d<u,F> - <dt u,F> + sig <dx u,F>.
        const double e = 2.0 * q2 - 4.0 * q4;
        double px;
        if constexpr (OP == KS_OP_COSTGRAD)
          px = fma(e + d, s_mf, s_x - s_ff);
        else
          px = fma(e, s_mf, fma(sig * q, s_d, -s_nv));
        double pt = fma(-sig * q, s_d, s_x), ps = q * s_d;
        if (!colact) px = pt = ps = 0.0;
        px = half_reduce(px);
        pt = half_reduce(pt);
        if constexpr (REL) ps = half_reduce(ps);
        if (valid && l16 == 0 && (a.out_params || a.out_tail)) {
          double gt = fix_t ? 0.0 : (-1.0 / T) * pt, gx = fix_x ? 0.0 : (1.0 / Lx) * px, gs = fix_s ? 0.0 : (-1.0 / T) * ps;
          if (OP == KS_OP_COSTGRAD && (PC_RUNTIME ? a.precond != 0 : PC)) {  // t <- t / T, x <- x / L^4   (:1069-1078)
            if (!fix_t) gt *= 1.0 / T;
            if (!fix_x) gx *= 1.0 / ((Lx * Lx) * (Lx * Lx));
          }
          if (a.out_tail) {
            const double g3[3] = {gt, gx, gs};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const int slot = ks::free_slot(a.cmask, i);
              if (slot >= 0) a.out_state[(size_t)orbit * out_ld + NM + slot] = g3[i];
            }
          } else {
            a.out_params[orbit * 3 + 0] = gt;
            a.out_params[orbit * 3 + 1] = gx;
            a.out_params[orbit * 3 + 2] = gs;
          }
        }
      }
      if constexpr (OP == KS_OP_EQN) {
        if constexpr (AS) {  // eqn ends here: the next pair's u modes go to the field stash now
          if (pair + nw < npairs) {
            const int obn = (2 * (pair + nw) + o < a.batch) ? 2 * (pair + nw) + o : a.batch - 1;
            const double* Mn = a.modes + (size_t)obn * NM + kk;
            static_for<0, 16>([&](auto Jc) { stage_tup_su<CLS, decltype(Jc)::value>(Mn, su + lane); });
          }
          cp_async_commit();
        }
        continue;
      }
      col_inverse_store(yr, yi, xcol, colact);
    }
    __syncwarp();
    // ---------------- phase D (ROW): second field (x64) times u (x64) -> forward ----------------
    {
      double zr[32], zi[32];
      row_load_inverse(xrow_a, xrow_b, zr, zi);
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        zr[x] *= su[x * 32 + lane];
        zi[x] *= su[(32 + x) * 32 + lane];
      }
      row_forward_store(zr, zi, xrow_a, xrow_b);
    }
    __syncwarp();
    // ---------------- phase E (COL): tuple stage 2 ----------------
    {
      double yr[32], yi[32];
      col_load(xcol, yr, yi);
      if constexpr (AS) {
        if constexpr (OP == KS_OP_COSTGRAD || OP == KS_OP_MATVEC) {  // stage-1 vector back from its L2 slot
          if (colact) static_for<0, 16>([&](auto Jc) { stage_tup_col<CLS, decltype(Jc)::value>(Fp, xcol); });
        }
        cp_async_commit();
        if (pair + nw < npairs) {  // next pair's u modes into the field stash (dead from here to the next phase B)
          const int obn = (2 * (pair + nw) + o < a.batch) ? 2 * (pair + nw) + o : a.batch - 1;
          const double* Mn = a.modes + (size_t)obn * NM + kk;
          static_for<0, 16>([&](auto Jc) { stage_tup_su<CLS, decltype(Jc)::value>(Mn, su + lane); });
        }
        cp_async_commit();
      }
      ksfft::fft<32, 1>(yr, yi);
      if constexpr (AS) cp_async_wait<1>();
      double* Op = a.out_state + (size_t)ob * out_ld + kk;
      static_for<0, 16>([&](auto Jc) {
        constexpr int J = decltype(Jc)::value;
        const double w = (J == 0) ? 0.0 : wbase * (J * (1.0 / 32.0));
        const Tup P = spec_to_tup<J>(yr, yi);
        const double cp = (J == 0 ? ks::kSqrt2 * cP : cP);
        Tup R;
        if constexpr (OP == KS_OP_MATVEC) {
          // + 2 * 1/2 Dx(u v)                                                        (:692)
          Tup A0;
          if constexpr (TS) A0 = ts_load<J>(ts);
          else if constexpr (AS) A0 = unstage_tup_col<CLS, J>(xcol);
          else A0 = load_tup<CLS, J>(Fp);
          const double qp = cp * q;
          R = Tup{fma(-qp, P.b1, A0.a1), fma(-qp, P.b2, A0.a2), fma(qp, P.a1, A0.b1), fma(qp, P.a2, A0.b2)};
        } else {
          // J^T v = -v_t + v_xx + v_xxxx [+ (S/T) v_x] - FFT2(u .* IFFT2(Dx v))       (:2041-2045, :3110-3115, :1977)
          Tup V;
          if constexpr (OP == KS_OP_COSTGRAD) {
            if constexpr (TS) V = ts_load<J>(ts);
            else if constexpr (AS) V = unstage_tup_col<CLS, J>(xcol);
            else V = load_tup<CLS, J>(Fp);
          } else {
            V = load_tup<CLS, J>(Vp);
          }
          R.a1 = fma(d, V.a1, fma(w, V.a2, -cp * P.a1));
          R.a2 = fma(d, V.a2, fma(-w, V.a1, -cp * P.a2));
          R.b1 = fma(d, V.b1, fma(w, V.b2, -cp * P.b1));
          R.b2 = fma(d, V.b2, fma(-w, V.b1, -cp * P.b2));
          if constexpr (REL) {
            const double sq = sig * q;
            R.a1 = fma(-sq, V.b1, R.a1);
            R.a2 = fma(-sq, V.b2, R.a2);
            R.b1 = fma(sq, V.a1, R.b1);
            R.b2 = fma(sq, V.a2, R.b2);
          }
          if (OP == KS_OP_COSTGRAD && (PC_RUNTIME ? a.precond != 0 : PC)) {  // 1 / (|w_j| + q^2 + q^4)     (:1061-1065)
            const double pm = 1.0 / ((w < 0 ? -w : w) + q2 + q4);
            R.a1 *= pm; R.a2 *= pm; R.b1 *= pm; R.b2 *= pm;
          }
        }
        project<CLS, J>(R);
        if (st) store_tup<CLS, J>(Op, R);
      });
    }
  }
}

}  // namespace ksw32
