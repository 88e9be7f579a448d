// 32x32 eqn / costgrad, "one orbit per warp, three warps per scheduler" form (variant 5).
//
// Same mathematics, lane mapping, tile layout and scalings as ks_warp32x1.cuh (reference orbithunter/ks/orbits.py: eqn
// :526-555, costgrad :757-791, rmatvec :711-755 with _rmatvec_parameters :1984-2024 / Relative :3059-3094, precondition
// :1034-1089), but built for 12 resident warps per SM instead of 8.  profiles/r01s_ncu_full_x1_costgrad32_summary.txt
// shows why: with two warps per scheduler the fp64 pipe is 57-63 % busy -- each warp spends half of an evaluation in
// phases that do not feed the pipe (shared-memory exchanges, scoreboard waits), and a single partner warp cannot cover
// them.  A third warp per scheduler needs <= 168 registers per thread, so
//   * the field row u(t, .) is kept in a lane-private shared-memory stash between phases B and D (16-byte accesses)
//     instead of 64 registers,
//   * u's modes are read straight from global memory at the start of phase A (the other warps of the scheduler cover the
//     latency; the next orbit's lines are prefetched into L2) -- no cp.async landing zone, which also takes a quarter of
//     the shared-memory traffic away (the crossbar is the second busiest unit of the x1 kernel).
// Per warp: XB tile 7.5 KB + field stash 8 KB = 15.5 KB -> 12 warps = 186 KB per SM.
#pragma once
#include "ks_warp32x1.cuh"

namespace ksx3 {
using ks::static_for;
using ksx1::irfft32;
using ksx1::rfft32_fwd;
using ksx1::warp_sum;

constexpr int kWarpsPerCta = 12;
constexpr int kTileDoubles = ksx1::kTileDoubles;  // 960
constexpr int kStashDoubles = 32 * 32;             // field row of every lane, as 16 double2 per lane
constexpr int kWarpSmemDoubles = kTileDoubles + kStashDoubles;
constexpr int kSmemBytes = kWarpsPerCta * kWarpSmemDoubles * 8;

KS_DEVI void prefetch_l2(const void* p) {
#ifndef KS_HOST_EMUL
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

template <int CLS, int OP, bool PC>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 1) ks_warp32x3_kernel(const KsEvalArgs a) {
  static_assert(OP == KS_OP_EQN || OP == KS_OP_COSTGRAD, "eqn and costgrad");
  constexpr bool GRAD = (OP == KS_OP_COSTGRAD);
  constexpr bool FULL = (CLS == KS_ORBIT || CLS == KS_RELATIVE);
  constexpr bool REL = (CLS == KS_RELATIVE);
  constexpr int COLS = FULL ? 30 : 15, NM = 31 * COLS;
  KS_DYN_SMEM(double, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* tile = smem + warp * kWarpSmemDoubles;  // tile[(t * 15 + c) * 2 + component]
  double2* stash = reinterpret_cast<double2*>(tile + kTileDoubles) + lane;  // pair pr of this lane's row at stash[pr * 32]
  const int c = lane >> 1, h = lane & 1;
  const bool colact = c < 15;
  const int cc = colact ? c : 14;
  const bool presE = FULL || h == 1;                                   // half tuple present at even time index
  const bool presO = FULL || (CLS == KS_ANTISYMMETRIC ? h == 1 : h == 0);  // ... at odd time index
  const int gw = warp * gridDim.x + blockIdx.x, nw = gridDim.x * kWarpsPerCta;  // warp-major (see ks_warp32x1.cuh)
  const bool fix_t = a.cmask & KS_FIX_T, fix_x = a.cmask & KS_FIX_X, fix_s = (a.cmask & KS_FIX_S) || !REL;
  constexpr double cN = 0.5 / (64.0 * 4096.0), cP = 1.0 / (64.0 * 4096.0);
  const int coff = (FULL ? h * 15 : 0) + cc;
  const double q_unit = (cc + 1) * (1.0 / 32.0);

  ks_pdl_launch_dependents();
  ks_pdl_wait();
  if (gw < a.batch) prefetch_l2(reinterpret_cast<const char*>(a.modes + (size_t)gw * NM) + lane * 256);
#pragma unroll 1
  for (int orbit = gw; orbit < a.batch; orbit += nw) {
    // ---------------- phase A (COL): half tuples of u -> A[t] or B[t] -> tile component h ----------------
    double m1[16], m2[16];  // half tuples (x1, x2) of u; overwritten by F in stage 1
    {
      const double* Mn = a.modes + (size_t)orbit * NM + coff;
      static_for<0, 16>([&](auto Jc) {
        constexpr int J = decltype(Jc)::value;
        const bool pres = (J & 1) ? presO : presE;
        m1[J] = pres ? Mn[J * COLS] : 0.0;
        if constexpr (J > 0) m2[J] = pres ? Mn[(15 + J) * COLS] : 0.0; else m2[J] = 0.0;
      });
      if (orbit + nw < a.batch && lane * 256 < NM * 8)
        prefetch_l2(reinterpret_cast<const char*>(a.modes + (size_t)(orbit + nw) * NM) + lane * 256);
    }
    {
      double x[32];
      double hr[16], hi[16];
#pragma unroll
      for (int J = 1; J < 16; ++J) {
        hr[J] = m1[J];
        hi[J] = m2[J];
      }
      hr[0] = hi[0] = 0.0;
      irfft32(ks::kSqrt2 * m1[0], hr, hi, x);
      if (colact) {
#pragma unroll
        for (int t = 0; t < 32; ++t) tile[(t * 15 + c) * 2 + h] = x[t];
      }
    }
    __syncwarp();
    // ---------------- phase B (ROW): row `lane` -> 64 u -> stash; square -> forward -> tile ----------------
    {
      double hr[16], hi[16];
      const double2* row = reinterpret_cast<const double2*>(tile) + lane * 15;
#pragma unroll
      for (int k = 1; k < 16; ++k) {
        const double2 v = row[k - 1];
        hr[k] = v.x;
        hi[k] = v.y;
      }
      hr[0] = hi[0] = 0.0;
      double U[32];
      irfft32(0.0, hr, hi, U);
#pragma unroll
      for (int pr = 0; pr < 16; ++pr) stash[pr * 32] = make_double2(U[2 * pr], U[2 * pr + 1]);
#pragma unroll
      for (int x = 0; x < 32; ++x) U[x] = U[x] * U[x];
      double c0;
      rfft32_fwd(U, c0, hr, hi);
      double2* wrow = reinterpret_cast<double2*>(tile) + lane * 15;
#pragma unroll
      for (int k = 1; k < 16; ++k) wrow[k - 1] = make_double2(hr[k], hi[k]);
    }
    __syncwarp();
    // ---------------- phase C (COL): N = 1/2 Dx(.) via the other component; stage 1; W = Dx F via the store ----------------
    {
      // per-orbit scalars are recomputed in the phases that use them (C and E) from a fresh load of (T, L, S): nothing
      // but the half tuples stays live across the row phases, which is what keeps the kernel inside 168 registers
      const double T = a.params[orbit * 3 + 0], Lx = a.params[orbit * 3 + 1], S = a.params[orbit * 3 + 2];
      const double wbase = -1.0 * (ks::kTwoPi * 32.0 / T);
      const double q = (ks::kTwoPi * 32.0 / Lx) * q_unit;
      const double q2 = q * q, q4 = q2 * q2, d = q4 - q2;
      const double sgq = h ? q : -q;  // d/dx: a' = -q b, b' = +q a
      const double sig = REL ? S / T : 0.0;
      double x[32];
      const double scl = 4.0 * sgq * cN;
#pragma unroll
      for (int t = 0; t < 32; ++t) x[t] = scl * tile[(t * 15 + cc) * 2 + (1 - h)];
      double n0, nr[16], ni[16];
      rfft32_fwd(x, n0, nr, ni);
      nr[0] = (0.5 * ks::kSqrt2) * n0;
      ni[0] = 0.0;
      double s_ff = 0.0, s_mf = 0.0, s_x = 0.0, s_d = 0.0;
      static_for<0, 16>([&](auto Jc) {
        constexpr int J = decltype(Jc)::value;
        const bool pres = (J & 1) ? presO : presE;
        const double w = (J == 0) ? 0.0 : wbase * (J * (1.0 / 32.0));
        const double u1 = m1[J], u2 = m2[J];
        double f1 = fma(d, u1, fma(-w, u2, nr[J]));
        double f2 = (J == 0) ? 0.0 : fma(d, u2, fma(w, u1, ni[J]));
        if constexpr (REL) {  // - (S/T) dx(u): a += sig q b, b -= sig q a  (partner's half via shuffle)
          const double p1 = __shfl_xor_sync(0xffffffffu, u1, 1), p2 = __shfl_xor_sync(0xffffffffu, u2, 1);
          f1 = fma(-sig * sgq, p1, f1);
          if constexpr (J > 0) f2 = fma(-sig * sgq, p2, f2);
          if constexpr (GRAD) s_d += sgq * (p1 * f1 + p2 * f2);  // <dx u, F> pieces
        }
        if (!pres) {
          f1 = 0.0;
          f2 = 0.0;
        }
        s_ff = fma(f1, f1, fma(f2, f2, s_ff));
        if constexpr (GRAD) {
          s_mf = fma(u1, f1, fma(u2, f2, s_mf));
          s_x = fma(w, fma(u1, f2, -u2 * f1), s_x);
        }
        m1[J] = f1;
        m2[J] = f2;
      });
      double* Fp = (OP == KS_OP_EQN) ? a.out_state : a.f_out;
      if (Fp) {
        Fp += (size_t)orbit * NM + coff;
        static_for<0, 16>([&](auto Jc) {
          constexpr int J = decltype(Jc)::value;
          const bool pres = (J & 1) ? presO : presE;
          if (colact && pres) {
            Fp[J * COLS] = m1[J];
            if constexpr (J > 0) Fp[(15 + J) * COLS] = m2[J];
          }
        });
      }
      const double lw = colact ? 1.0 : 0.0;
      const double ff = warp_sum(s_ff * lw);
      if (lane == 0 && a.out_cost) a.out_cost[orbit] = 0.5 * ff;
      if constexpr (GRAD) {
        const double e = 2.0 * q2 - 4.0 * q4;
        double px = warp_sum(fma(e + d, s_mf, s_x - s_ff) * lw);
        double pt = warp_sum(fma(-sig, s_d, s_x) * lw);
        double ps = REL ? warp_sum(s_d * lw) : 0.0;
        if (lane == 0 && (a.out_params || a.out_tail)) {
          double gt = fix_t ? 0.0 : (-1.0 / T) * pt, gx = fix_x ? 0.0 : (1.0 / Lx) * px, gs = fix_s ? 0.0 : (-1.0 / T) * ps;
          if constexpr (PC) {
            if (!fix_t) gt *= 1.0 / T;
            if (!fix_x) gx *= 1.0 / ((Lx * Lx) * (Lx * Lx));
          }
          if (a.out_tail) {
            const int out_ld = a.out_ld ? a.out_ld : NM;
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
        __syncwarp();
        continue;
      }
      double hr[16], hi[16];
#pragma unroll
      for (int J = 1; J < 16; ++J) {
        hr[J] = m1[J];
        hi[J] = m2[J];
      }
      hr[0] = hi[0] = 0.0;
      irfft32(ks::kSqrt2 * m1[0], hr, hi, x);
      // Dx F = iq (A_F + i B_F): h0 writes q A_F to .y, h1 writes -q B_F to .x.  The power-of-two normalisation of the
      // second product (4 cP, applied by ks_warp32x1.cuh when phase E reads the tile) rides on this multiplication.
      const double sw = (h ? -q : q) * (4.0 * cP);
      if (colact) {
#pragma unroll
        for (int t = 0; t < 32; ++t) tile[(t * 15 + c) * 2 + (1 - h)] = sw * x[t];
      }
    }
    __syncwarp();
    // ---------------- phase D (ROW): second field times u (from the stash) -> forward -> tile ----------------
    {
      double hr[16], hi[16];
      const double2* row = reinterpret_cast<const double2*>(tile) + lane * 15;
#pragma unroll
      for (int k = 1; k < 16; ++k) {
        const double2 v = row[k - 1];
        hr[k] = v.x;
        hi[k] = v.y;
      }
      hr[0] = hi[0] = 0.0;
      double w[32];
      irfft32(0.0, hr, hi, w);
#pragma unroll
      for (int pr = 0; pr < 16; ++pr) {
        const double2 uu = stash[pr * 32];
        w[2 * pr] *= uu.x;
        w[2 * pr + 1] *= uu.y;
      }
      double c0;
      rfft32_fwd(w, c0, hr, hi);
      double2* wrow = reinterpret_cast<double2*>(tile) + lane * 15;
#pragma unroll
      for (int k = 1; k < 16; ++k) wrow[k - 1] = make_double2(hr[k], hi[k]);
    }
    __syncwarp();
    // ---------------- phase E (COL): stage 2: grad = -F_t + F_xx + F_xxxx [+ (S/T) F_x] - T(Fx(u w)) ----------------
    {
      const double T = a.params[orbit * 3 + 0], Lx = a.params[orbit * 3 + 1], S = a.params[orbit * 3 + 2];
      const double wbase = -1.0 * (ks::kTwoPi * 32.0 / T);
      const double q = (ks::kTwoPi * 32.0 / Lx) * q_unit;
      const double q2 = q * q, q4 = q2 * q2, d = q4 - q2;
      const double sgq = h ? q : -q;  // d/dx: a' = -q b, b' = +q a
      const double sig = REL ? S / T : 0.0;
      double x[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) x[t] = tile[(t * 15 + cc) * 2 + h];
      double p0, pr[16], pi[16];
      rfft32_fwd(x, p0, pr, pi);
      pr[0] = (0.5 * ks::kSqrt2) * p0;
      pi[0] = 0.0;
      const int out_ld = a.out_ld ? a.out_ld : NM;
      double* Op = a.out_state + (size_t)orbit * out_ld + coff;
      static_for<0, 16>([&](auto Jc) {
        constexpr int J = decltype(Jc)::value;
        const bool pres = (J & 1) ? presO : presE;
        const double w = (J == 0) ? 0.0 : wbase * (J * (1.0 / 32.0));
        const double f1 = m1[J], f2 = m2[J];
        double g1 = fma(d, f1, fma(w, f2, -pr[J]));
        double g2 = (J == 0) ? 0.0 : fma(d, f2, fma(-w, f1, -pi[J]));
        if constexpr (REL) {  // + (S/T) dx(F)
          const double p1 = __shfl_xor_sync(0xffffffffu, f1, 1), p2 = __shfl_xor_sync(0xffffffffu, f2, 1);
          g1 = fma(sig * sgq, p1, g1);
          if constexpr (J > 0) g2 = fma(sig * sgq, p2, g2);
        }
        if constexpr (PC) {
          const double pm = 1.0 / ((w < 0 ? -w : w) + q2 + q4);
          g1 *= pm;
          g2 *= pm;
        }
        if (colact && pres) {
          Op[J * COLS] = g1;
          if constexpr (J > 0) Op[(15 + J) * COLS] = g2;
        }
      });
    }
    __syncwarp();  // the tile is rewritten by the next orbit's phase A (other lanes' columns are read in phase E)
  }
}

}  // namespace ksx3
