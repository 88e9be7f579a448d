// 32x32 eqn / costgrad, "ping-pong" form (variant 6): one orbit per warp with all state in registers, 8 warps per SM as in
// ks_warp32x1.cuh, but the two warps of every SM scheduler take turns on the fp64 pipe.
//
// Why (profiles/r02e_fft_stream_code_size.jsonl, r02d_x3_knockout_decomposition.jsonl, r02b ncu of variant 5): one fully
// unrolled evaluation is 2 800 instructions = 45 KB of code, more than the SM's instruction cache serves at full rate.  The
// same register-resident transforms plus shared-memory exchange reach 98 % of the measured fp64 peak as a 7 KB or 36 KB loop
// but 64 % as a 57 KB loop (tools/fft_stream.cu) -- exactly where ks_warp32x1.cuh (8 warps) and ks_warp32x3.cuh (12 warps)
// are stuck: free-running warps drift apart and together walk over the whole 45 KB all the time.  Rolling the phases into
// loops costs more in register shuffles than it saves (round 1, and variant "p" of this round).
//
// Here the eight warps form two groups (warps 0-3 and 4-7: one warp of each group per scheduler).  An evaluation is a
// chain of "compute chunks" (the register-resident transforms and the pointwise algebra: phases A..E) separated by
// "exchange chunks" (tile stores, __syncwarp, tile loads, global stores).  A token passed through two named barriers lets
// exactly one group run its compute chunk while the other group runs its exchange chunk:
//   group 0:                compute; bar.arrive(1); exchange; bar.sync(2); compute; ...
//   group 1:  bar.sync(1);  compute; bar.arrive(2); exchange; bar.sync(1); compute; ...
// so (i) the fp64 pipe of a scheduler always has one warp streaming FMAs at full rate (a single warp's FFT stream reaches
// 99 % of the pipe, tools/fft_stream.cu kind 0 at 4 warps/SM) while the partner's shared-memory traffic is hidden behind
// it, and (ii) the groups stay one chunk apart in the code, so the instruction working set is two chunks (< 16 KB).
// The barriers carry no data dependence (tiles are private to a warp); every warp of a CTA runs the same number of
// loop trips so that the barrier counts match (idle trips only pass the token on).
// Mathematics, scalings and the lane mapping are those of ks_warp32x1.cuh (reference orbithunter/ks/orbits.py: eqn
// :526-555, costgrad :757-791, rmatvec :711-755, _rmatvec_parameters :1984-2024 / Relative :3059-3094, precondition
// :1034-1089).
#pragma once
#include "ks_warp32x1.cuh"

namespace kspp {
using ks::static_for;
using ksx1::irfft32;
using ksx1::rfft32_fwd;
using ksx1::warp_sum;

constexpr int kWarpsPerCta = 8;
constexpr int kTileDoubles = ksx1::kTileDoubles;    // 960
constexpr int kStageDoubles = ksx1::kStageDoubles;  // cp.async landing zone of the next orbit's half tuples
constexpr int kWarpSmemDoubles = kTileDoubles + kStageDoubles;
constexpr int kSmemBytes = kWarpsPerCta * kWarpSmemDoubles * 8;

KS_DEVI void bar_sync(int id) {
#ifndef KS_HOST_EMUL
  asm volatile("barrier.sync.aligned %0, 256;" ::"r"(id) : "memory");
#else
  (void)id;
#endif
}
KS_DEVI void bar_arrive(int id) {
#ifndef KS_HOST_EMUL
  asm volatile("barrier.arrive.aligned %0, 256;" ::"r"(id) : "memory");
#else
  (void)id;
#endif
}

template <int CLS, int OP, bool PC>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 1) ks_warp32pp_kernel(const KsEvalArgs a) {
  static_assert(OP == KS_OP_EQN || OP == KS_OP_COSTGRAD, "eqn and costgrad");
  constexpr bool GRAD = (OP == KS_OP_COSTGRAD);
  constexpr bool FULL = (CLS == KS_ORBIT || CLS == KS_RELATIVE);
  constexpr bool REL = (CLS == KS_RELATIVE);
  constexpr int COLS = FULL ? 30 : 15, NM = 31 * COLS;
  KS_DYN_SMEM(double, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* tile = smem + warp * kWarpSmemDoubles;  // tile[(t * 15 + c) * 2 + component]
  double* stage = tile + kTileDoubles + lane;     // slot s of this lane at stage[s * 32]
  const int c = lane >> 1, h = lane & 1;
  const bool colact = c < 15;
  const int cc = colact ? c : 14;
  const bool presE = FULL || h == 1;                                        // half tuple present at even time index
  const bool presO = FULL || (CLS == KS_ANTISYMMETRIC ? h == 1 : h == 0);  // ... at odd time index
  const int gw = warp * gridDim.x + blockIdx.x, nw = gridDim.x * kWarpsPerCta;  // warp-major (see ks_warp32x1.cuh)
  const bool fix_t = a.cmask & KS_FIX_T, fix_x = a.cmask & KS_FIX_X, fix_s = (a.cmask & KS_FIX_S) || !REL;
  constexpr double cN = 0.5 / (64.0 * 4096.0), cP = 1.0 / (64.0 * 4096.0);
  const int coff = (FULL ? h * 15 : 0) + cc;
  const double q_unit = (cc + 1) * (1.0 / 32.0);
  // ---- the token: group 0 = warps 0-3, group 1 = warps 4-7; barrier 1 releases group 1, barrier 2 releases group 0 ----
  const int group = warp >> 2;
  const int my_bar = group ? 1 : 2, other_bar = group ? 2 : 1;
  const bool pp = a.pingpong == 1;
  const bool pair_lockstep = a.pingpong == 2;  // tuning: the two warps of a scheduler start every compute chunk together
  // warp 0 of the CTA owns the lowest orbit index, so its trip count is the CTA's maximum; everybody runs that many trips
  const int trips = (a.batch - (int)blockIdx.x + nw - 1) / nw;
  constexpr int kChunks = (OP == KS_OP_EQN) ? 3 : 5;
  int chunks_left = trips * kChunks;  // compute chunks this warp still has to take the token for
  auto release = [&]() {  // end of a compute chunk: hand the token to the other group
    if (pp) bar_arrive(other_bar);
  };
  auto acquire = [&]() {  // before a compute chunk; the very last wait of group 1 has no matching release
    --chunks_left;
    if (pp && !(group == 1 && chunks_left == 0)) bar_sync(my_bar);
#ifndef KS_HOST_EMUL
    if (pair_lockstep) asm volatile("barrier.sync.aligned %0, 64;" ::"r"(3 + (warp & 3)) : "memory");
#endif
  };

  auto prefetch = [&](int orb) {  // this lane's half tuples of orbit `orb` -> its stage slots (cp.async, 8 bytes each)
    if (orb < a.batch) {
      const double* Mn = a.modes + (size_t)orb * NM + coff;
      static_for<0, 16>([&](auto Jc) {
        constexpr int J = decltype(Jc)::value;
        const bool pres = (J & 1) ? presO : presE;
        if (pres) {
          ksw32::cp_async8(stage + (J == 0 ? 0 : 2 * J - 1) * 32, Mn + J * COLS);
          if constexpr (J > 0) ksw32::cp_async8(stage + (2 * J) * 32, Mn + (15 + J) * COLS);
        }
      });
    }
    ksw32::cp_async_commit();
  };

  ks_pdl_launch_dependents();
  ks_pdl_wait();
  prefetch(gw);
  if (pp && group == 1) bar_sync(my_bar);  // group 0 computes first
#pragma unroll 1
  for (int trip = 0, orbit = gw; trip < trips; ++trip, orbit += nw) {
    if (orbit >= a.batch) {  // idle trip (uneven tail of the batch): keep the token moving
#pragma unroll 1
      for (int k = 0; k < kChunks; ++k) {
        release();
        acquire();
      }
      continue;
    }
    const double T = a.params[orbit * 3 + 0], Lx = a.params[orbit * 3 + 1], S = a.params[orbit * 3 + 2];
    const double wbase = -1.0 * (ks::kTwoPi * 32.0 / T);
    const double q = (ks::kTwoPi * 32.0 / Lx) * q_unit;
    const double q2 = q * q, q4 = q2 * q2, d = q4 - q2;
    const double sgq = h ? q : -q;  // d/dx: a' = -q b, b' = +q a
    const double sig = REL ? S / T : 0.0;
    // ---------------- phase A (COL): half tuples of u -> A[t] or B[t] -> tile component h ----------------
    double m1[16], m2[16];  // half tuples (x1, x2) of u; overwritten by F in stage 1
    ksw32::cp_async_wait<0>();
    static_for<0, 16>([&](auto Jc) {
      constexpr int J = decltype(Jc)::value;
      const bool pres = (J & 1) ? presO : presE;
      m1[J] = pres ? stage[(J == 0 ? 0 : 2 * J - 1) * 32] : 0.0;
      if constexpr (J > 0) m2[J] = pres ? stage[(2 * J) * 32] : 0.0; else m2[J] = 0.0;
    });
    prefetch(orbit + nw);
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
      release();
      if (colact) {
#pragma unroll
        for (int t = 0; t < 32; ++t) tile[(t * 15 + c) * 2 + h] = x[t];
      }
    }
    __syncwarp();
    // ---------------- phase B (ROW): row `lane` -> 64 u -> keep; square -> forward -> tile ----------------
    double U[32];
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
      acquire();
      irfft32(0.0, hr, hi, U);
      double s[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) s[x] = U[x] * U[x];
      double c0;
      rfft32_fwd(s, c0, hr, hi);
      release();
      double2* wrow = reinterpret_cast<double2*>(tile) + lane * 15;
#pragma unroll
      for (int k = 1; k < 16; ++k) wrow[k - 1] = make_double2(hr[k], hi[k]);
    }
    __syncwarp();
    // ---------------- phase C (COL): N = 1/2 Dx(.) via the other component; stage 1; W = Dx F via the store ----------------
    {
      double x[32];
      const double scl = 4.0 * sgq * cN;
#pragma unroll
      for (int t = 0; t < 32; ++t) x[t] = tile[(t * 15 + cc) * 2 + (1 - h)];
      acquire();
#pragma unroll
      for (int t = 0; t < 32; ++t) x[t] *= scl;
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
        release();
        __syncwarp();
        acquire();  // the next trip's phase A
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
#pragma unroll
      for (int t = 0; t < 32; ++t) x[t] *= sw;
      release();
      if (colact) {
#pragma unroll
        for (int t = 0; t < 32; ++t) tile[(t * 15 + c) * 2 + (1 - h)] = x[t];
      }
    }
    __syncwarp();
    // ---------------- phase D (ROW): second field times u -> forward -> tile ----------------
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
      acquire();
      double w[32];
      irfft32(0.0, hr, hi, w);
#pragma unroll
      for (int x = 0; x < 32; ++x) w[x] *= U[x];
      double c0;
      rfft32_fwd(w, c0, hr, hi);
      release();
      double2* wrow = reinterpret_cast<double2*>(tile) + lane * 15;
#pragma unroll
      for (int k = 1; k < 16; ++k) wrow[k - 1] = make_double2(hr[k], hi[k]);
    }
    __syncwarp();
    // ---------------- phase E (COL): stage 2: grad = -F_t + F_xx + F_xxxx [+ (S/T) F_x] - T(Fx(u w)) ----------------
    {
      double x[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) x[t] = tile[(t * 15 + cc) * 2 + h];
      acquire();
      double p0, pr[16], pi[16];
      rfft32_fwd(x, p0, pr, pi);
      pr[0] = (0.5 * ks::kSqrt2) * p0;
      pi[0] = 0.0;
      double g1v[16], g2v[16];
      static_for<0, 16>([&](auto Jc) {
        constexpr int J = decltype(Jc)::value;
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
        g1v[J] = g1;
        g2v[J] = g2;
      });
      release();
      const int out_ld = a.out_ld ? a.out_ld : NM;
      double* Op = a.out_state + (size_t)orbit * out_ld + coff;
      static_for<0, 16>([&](auto Jc) {
        constexpr int J = decltype(Jc)::value;
        const bool pres = (J & 1) ? presO : presE;
        if (colact && pres) {
          Op[J * COLS] = g1v[J];
          if constexpr (J > 0) Op[(15 + J) * COLS] = g2v[J];
        }
      });
    }
    __syncwarp();  // the tile is rewritten by the next orbit's phase A (other lanes' columns are read in phase E)
    acquire();     // the next trip's phase A (or the end of the token chain)
  }
}

}  // namespace kspp
