// 32x32 eqn / costgrad / matvec / rmatvec and the fused adjoint-descent passes, "one orbit per warp, state in registers" form
// (variant 4, the default for every 32x32 operation).  matvec / rmatvec also have COLUMN MODES for the dense Newton step
// (KsEvalArgs::u_div, unit_v): evaluation e works on orbit e / u_div and, with unit_v, on the unit vector of cdof slot
// e % u_div made in registers -- the Jacobian (matvec over unit vectors) and, the route ks_newton_dense takes, J J^T column by
// column as J (J^T e_j) without a Jacobian in memory.
//
// Same mathematics and reference citations as ks_warp32.cuh, different work decomposition, chosen after profiling the
// two-orbits-per-warp kernels (profiles/r01*): those need 128 registers of transform buffer per lane plus 31-47 KB of
// shared memory per warp, which caps residency at 4-7 warps/SM and forces either L2 round trips or one warp per
// scheduler.  Here every 1-D transform is a REAL length-32 transform done as a 16-point complex FFT in registers
// (64 registers of buffer), so the field stash (32 doubles) and the tuple stash (31 doubles) also live in registers,
// the only shared memory is the 7.5 KB spatial-mode tile, and 8 warps (2 per scheduler) are resident per SM.
//
//   ROW phases : lane t owns field row t (32 reals).
//   COL phases : lane 2c+h owns, for wavenumber c+1, the real sequence A[t] = Re X[t,c] (h = 0) or B[t] = Im X[t,c]
//                (h = 1); its spectrum is exactly the a-half (a1 + i a2) or b-half (b1 + i b2) of the reference's
//                tuples, so each lane holds "half tuples" for all 16 time indices.  d/dx (multiplication of X by iq)
//                is folded into which component of the tile a lane reads or writes, so the two halves never have to be
//                exchanged except for RelativeOrbitKS' shift terms (warp shuffles).
// Scalings are those of ks_warp32.cuh: field arrays hold 64 u, forward outputs are doubled, constants cN, cP.
#pragma once
#include "ks_fft32_gen.cuh"
#include "ks_warp32.cuh"

namespace ksx1 {
using ks::static_for;
using ksfft::cos64;
using ksfft::sin64;

#ifdef KS_X1_WARPS_PER_CTA
constexpr int kWarpsPerCta = KS_X1_WARPS_PER_CTA;  // tuning build: 4 -> two CTAs per SM, 2 -> four
#else
constexpr int kWarpsPerCta = 8;
#endif
constexpr int kCtasPerSm = kWarpsPerCta >= 8 ? 1 : 8 / kWarpsPerCta;
constexpr int kTileDoubles = 32 * 15 * 2;  // XB tile of one orbit: 480 double2
constexpr int kPackedDoubles = 16 * 32 * 2;  // per orbit, packed lane-major layout of the fused adjoint pass (below)
constexpr int kStageDoubles = 31 * 32;      // cp.async landing zone for the next orbit's half tuples (lane-private slots)
// -DKS_X1_PARK: u's and F's half tuples wait in shared memory while the row phases run (below) instead of in 62 registers.
// Measured (profiles/r02w_x1_register_parking.jsonl): no gain -- 26.0 vs 25.0 us per 4096-orbit launch, 337 vs 334 us per
// 65 536 -- ptxas still takes all 255 registers and the dependent-issue stalls stay; off by default.
#ifdef KS_X1_PARK
constexpr bool kPark = true;
#else
constexpr bool kPark = false;
#endif
// adj: u and g, packed; otherwise the landing zone of u and -- for matvec / rmatvec (v's landing zone) or kPark -- a second
// lane-private area of the same size
KS_HD constexpr int warp_smem_doubles(bool adj, bool vop = false) {
  return kTileDoubles + (adj ? 2 * kPackedDoubles : ((kPark || vop) ? 2 : 1) * kStageDoubles);
}
KS_HD constexpr int smem_bytes(bool adj, bool vop = false) { return kWarpsPerCta * warp_smem_doubles(adj, vop) * 8; }

// x[32] real -> c0 = DFT32(x)[0], (cr, ci)[k] = DFT32(x)[k], k = 1..15   (forward sign e^{-i...}, unnormalised).
// ks_warp32.cuh's convention has the k >= 1 outputs doubled; here that factor (x2 per forward transform, x4 for the two
// dimensions) is folded into the scale constants applied at the tile loads of the COL phases.
KS_DEVI void rfft32_fwd(const double (&x)[32], double& c0, double (&cr)[16], double (&ci)[16]) {
#ifndef KS_X1_PACKED_IRFFT
  ksg32::rfft32(x, c0, cr, ci);  // generated decimation in time, 163 instructions (tools/gen_fft32.py)
#else
  double yr[17], yi[17];
  ksfft::rfft_dit<32, 1, 0, 32>(x, yr, yi);
  c0 = yr[0];
#pragma unroll
  for (int k = 1; k < 16; ++k) {
    cr[k] = yr[k];
    ci[k] = yi[k];
  }
  cr[0] = ci[0] = 0.0;
#endif
}

// h0 real, (hr, hi)[k], k = 1..15 -> x[n] = h0 + sum_k 2 Re(H[k] e^{+2 pi i k n / 32})     (Nyquist term zero)
// Generated decimation-in-frequency transform with deferred twiddle scales (tools/gen_fft32.py): 162 fp64 instructions.
// -DKS_X1_PACKED_IRFFT selects the earlier route through a packed complex length-16 FFT (233 instructions; cross-check).
KS_DEVI void irfft32(double h0, const double (&hr)[16], const double (&hi)[16], double (&x)[32]) {
#ifndef KS_X1_PACKED_IRFFT
  ksg32::irfft32(h0, hr, hi, x);
#else
  double gr[16], gi[16];
  gr[0] = h0;
  gi[0] = h0;
  static_for<1, 8>([&](auto Kc) {
    constexpr int K = decltype(Kc)::value;
    constexpr double cs = cos64(2 * K), sn = sin64(2 * K);  // i w^K = -sin + i cos
    const double sr = hr[K] + hr[16 - K], si = hi[K] - hi[16 - K];
    const double dr = hr[K] - hr[16 - K], di = hi[K] + hi[16 - K];
    const double tr = -fma(sn, dr, cs * di), ti = fma(cs, dr, -sn * di);
    gr[K] = sr + tr;
    gi[K] = si + ti;
    gr[16 - K] = sr - tr;
    gi[16 - K] = -(si - ti);
  });
  gr[8] = 2.0 * hr[8];
  gi[8] = -2.0 * hi[8];
  ksfft::fft<16, -1>(gr, gi);
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    x[2 * m] = gr[m];
    x[2 * m + 1] = gi[m];
  }
#endif
}

// Packed lane-major layout of the fused adjoint pass' internal orbit / gradient buffers: per orbit 16 pairs x 32 lanes of
// double2; pair pr of lane l holds slots (2 pr, 2 pr + 1) of that lane's 31 half-tuple values (slot 0 = x1(0), slot 2J-1 =
// x1(J), slot 2J = x2(J)).  One 16-byte access per pair instead of two strided 8-byte ones.
KS_DEVI void cp_async16(double* smem_dst, const double* gsrc) {
#ifdef KS_HOST_EMUL
  smem_dst[0] = gsrc[0];
  smem_dst[1] = gsrc[1];
#else
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#endif
}

KS_DEVI double warp_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

template <int CLS, int OP, bool PC>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kCtasPerSm) ks_warp32x1_kernel(const KsEvalArgs a) {
  static_assert(OP == KS_OP_EQN || OP == KS_OP_COSTGRAD || OP == KS_OP_ADJSTEP || OP == KS_OP_ADJMULTI || OP == KS_OP_MATVEC ||
                    OP == KS_OP_RMATVEC,
                "eqn, costgrad, matvec, rmatvec and the fused adjoint steps");
  // ADJ : one adjoint-descent pass per launch (decide on the previous trial, evaluate the next), state in global memory.
  // ADJM: every orbit runs up to adj.inner passes inside one launch; its accepted point u and gradient g stay in the
  //       warp's shared-memory stage (lane-private packed slots), the scalars in registers; the decision on a trial is
  //       taken right after its evaluation.  Same arithmetic, hence the same decisions, as ADJ.
  constexpr bool ADJ = (OP == KS_OP_ADJSTEP);
  constexpr bool ADJM = (OP == KS_OP_ADJMULTI);
  constexpr bool GRAD = (OP == KS_OP_COSTGRAD) || ADJ || ADJM;
  // matvec / rmatvec (ks/orbits.py:659-755): the same five phases; stage 1 works on v's half tuples (landed in the second
  // lane-private area, or generated as a unit vector in Jacobian mode) instead of forming F
  constexpr bool MV = (OP == KS_OP_MATVEC), RMV = (OP == KS_OP_RMATVEC), VOP = MV || RMV;
  constexpr int kWarpSmemDoubles = warp_smem_doubles(ADJ || ADJM, VOP);
  constexpr bool FULL = (CLS == KS_ORBIT || CLS == KS_RELATIVE);
  constexpr bool REL = (CLS == KS_RELATIVE);
  constexpr int COLS = FULL ? 30 : 15, NM = 31 * COLS;
  KS_DYN_SMEM(double, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* tile = smem + warp * kWarpSmemDoubles;  // tile[(t * 15 + c) * 2 + component]
  double* stage = tile + kTileDoubles + lane;     // slot s of this lane at stage[s * 32]
  double* pstage = tile + kTileDoubles + lane * 2;  // packed variant (fused adjoint pass): pair pr at pstage[pr * 64]
  // PARK (eqn / costgrad): the row phases B and D are the register-hungry ones (field row + product + FFT working set), and
  // the 31 half tuples of u (needed again in stage 1) resp. F (needed again in stage 2) used to sit in 62 registers right
  // through them.  Now u is re-read from the landing zone in phase C -- the next orbit's prefetch is issued only then, which
  // still leaves it half an orbit to land -- and F waits in a second lane-private area until phase E.
  constexpr bool PARK = kPark && !(OP == KS_OP_ADJSTEP || OP == KS_OP_ADJMULTI) && !VOP;
  double* park = tile + kTileDoubles + kStageDoubles + lane;  // slot s of this lane at park[s * 32]
  const int c = lane >> 1, h = lane & 1;
  const bool colact = c < 15;
  const int cc = colact ? c : 14;
  // which half-tuples exist for this lane: even / odd time index (class projection)
  const bool presE = FULL || (CLS == KS_ANTISYMMETRIC ? h == 1 : h == 1);
  const bool presO = FULL || (CLS == KS_ANTISYMMETRIC ? h == 1 : h == 0);
  // warp-major numbering of the persistent warps: consecutive orbits go to different SMs, so the last (partial) round of
  // a batch is spread over all SMs with few warps each instead of filling some SMs and leaving the others idle
  const int gw = warp * gridDim.x + blockIdx.x, nw = gridDim.x * kWarpsPerCta;
  const bool fix_t = a.cmask & KS_FIX_T, fix_x = a.cmask & KS_FIX_X, fix_s = (a.cmask & KS_FIX_S) || !REL;
  constexpr double cN = 0.5 / (64.0 * 4096.0), cP = 1.0 / (64.0 * 4096.0);
  // column offset of this lane's half inside a modes row: full classes [Re_x | Im_x], half classes a single block
  const int coff = (FULL ? h * 15 : 0) + cc;
  const int udiv = VOP ? a.u_div : 0;            // column modes: evaluation e works on orbit e / udiv ...
  const bool unit_v = VOP && udiv > 0 && a.unit_v;  // ... with v = the unit vector of cdof slot e % udiv
  const int v_ld = a.v_ld ? a.v_ld : NM;
  auto uidx = [&](int e) { return udiv > 0 ? e / udiv : e; };

  // u's half tuples reach the registers through a cp.async landing zone filled one orbit ahead (the first-touch DRAM
  // latency of the modes was 26 % of the warp-issue samples with plain loads: profiles/r01g_*)
  struct Pending {  // what the next iteration needs to know about its orbit (all lanes hold the same values)
    int live, cur;
    double step, T, L, S;
  };
  // fused adjoint descent: judge the previous trial of orbit `orb` (optimize.py:478-480 backtracking, then
  // _process_correction :2174-2189 and the loop exit fix-up :506-507), publish the scalars, return what to evaluate next
  auto decide = [&](int orb) {
    Pending pd{0, 0, 0.0, 1.0, 1.0, 0.0};
    if (orb >= a.batch) return pd;
    if constexpr (ADJ) {
      const KsAdjFused& A = a.adj;
      int live = A.live[orb], cur = A.cur[orb];
      double step = A.step[orb];
      const double ct = A.tcost[orb];
      double cost = A.cost[orb];
      int status = A.status[orb], nit = A.nit[orb], accept = 0;
      __syncwarp();  // every lane has read the scalars before lane 0 publishes the updated ones
      if (live && !A.first && ct >= 0.0) {  // tcost < 0: no trial has been evaluated yet (pass right after the first)
        if (ct >= cost && step > A.min_step) {
          step *= 0.5;
        } else {
          nit += 1;
          if (ct <= A.tol) {
            status = -1;
            accept = 1;
          } else if (step <= A.min_step) {
            status = 0;
          } else if (nit == A.maxiter) {
            status = 2;
            accept = (ct < cost) ? 1 : 0;
          } else {
            const double big = fmax(fmax(cost, ct), 1.0);
            if ((cost - ct) / big < A.ftol) status = 3; else accept = 1;
          }
          if (accept) {
            cost = ct;
            cur ^= 1;
          }
          live = (cost > A.tol && status == -1) ? 1 : 0;
          if (!live && cost <= A.tol) status = -1;
        }
        if (lane == 0) {
          A.cost[orb] = cost;
          A.step[orb] = step;
          A.status[orb] = status;
          A.nit[orb] = nit;
          A.live[orb] = live;
          A.cur[orb] = cur;
        }
      }
      pd.live = live;
      pd.cur = cur;
      pd.step = A.first ? 0.0 : step;
      if (live) {
        if (lane == 0 && !A.first) atomicAdd(A.n_live, 1);
        const double* pp = A.p[cur] + orb * 3;
        const double* gg = A.gp[cur] + orb * 3;
        pd.T = fma(-pd.step, gg[0], pp[0]);
        pd.L = fma(-pd.step, gg[1], pp[1]);
        pd.S = fma(-pd.step, gg[2], pp[2]);
      }
    } else {
      pd.live = 1;
      pd.T = a.params[uidx(orb) * 3 + 0];
      pd.L = a.params[uidx(orb) * 3 + 1];
      pd.S = a.params[uidx(orb) * 3 + 2];
    }
    return pd;
  };
  auto prefetch = [&](int orb, const Pending& pd) {
    if (orb < a.batch && pd.live) {
      if constexpr (ADJ) {
        const double* Un = a.adj.u[pd.cur] + (size_t)orb * kPackedDoubles + lane * 2;
        const double* Gn = a.adj.g[pd.cur] + (size_t)orb * kPackedDoubles + lane * 2;
#pragma unroll
        for (int pr = 0; pr < 16; ++pr) {
          cp_async16(pstage + pr * 64, Un + pr * 64);
          cp_async16(pstage + kPackedDoubles + pr * 64, Gn + pr * 64);
        }
      } else {
        const double* Mn = a.modes + (size_t)uidx(orb) * NM + coff;
        static_for<0, 16>([&](auto Jc) {
          constexpr int J = decltype(Jc)::value;
          const bool pres = (J & 1) ? presO : presE;
          if (pres) {
            ksw32::cp_async8(stage + (J == 0 ? 0 : 2 * J - 1) * 32, Mn + J * COLS);
            if constexpr (J > 0) ksw32::cp_async8(stage + (2 * J) * 32, Mn + (15 + J) * COLS);
          }
        });
      }
    }
    ksw32::cp_async_commit();
  };
  // matvec / rmatvec: v's half tuples land in the second lane-private area (same slots as u's); issued once phase C has read
  // the current orbit's v.  Unit-vector modes have no v in memory.
  [[maybe_unused]] auto prefetch_v = [&](int orb) {
    if constexpr (VOP) {
      if (orb < a.batch && !unit_v) {
        const double* Vn = a.vmodes + (size_t)orb * v_ld + coff;
        static_for<0, 16>([&](auto Jc) {
          constexpr int J = decltype(Jc)::value;
          const bool pres = (J & 1) ? presO : presE;
          if (pres) {
            ksw32::cp_async8(park + (J == 0 ? 0 : 2 * J - 1) * 32, Vn + J * COLS);
            if constexpr (J > 0) ksw32::cp_async8(park + (2 * J) * 32, Vn + (15 + J) * COLS);
          }
        });
      }
      ksw32::cp_async_commit();
    }
  };
  // programmatic dependent launch: everything above overlapped the previous kernel's tail; nothing below may run before
  // that kernel's writes are visible
  ks_pdl_launch_dependents();
  ks_pdl_wait();
  // the warps of one scheduler (warp index / 4 = 0, 1, 2) start `stagger_ns` apart: identical work per orbit otherwise keeps
  // them in the same phase for the whole launch, so the fp64 pipe idles while all of them exchange data and vice versa
  if (a.stagger_ns > 0 && warp >= 4) ks_nanosleep((unsigned)(a.stagger_ns * (warp >> 2)));
  Pending next{0, 0, 0.0, 1.0, 1.0, 0.0};
  if constexpr (!ADJM) {
    next = decide(gw);
    prefetch(gw, next);
    prefetch_v(gw);
  }
#pragma unroll 1
  for (int orbit = gw; orbit < a.batch; orbit += nw) {
    Pending pd = next;
    // ADJM: this orbit's descent state (identical in every lane)
    [[maybe_unused]] int st_live = 1, st_status = -1, st_nit = 0, st_first = 0;
    [[maybe_unused]] double st_cost = 0.0, st_step = 0.0, st_T = 1.0, st_L = 1.0, st_S = 0.0, st_g0 = 0.0, st_g1 = 0.0, st_g2 = 0.0;
    [[maybe_unused]] double ct = 0.0, tg0 = 0.0, tg1 = 0.0, tg2 = 0.0;
    if constexpr (ADJM) {
      const KsAdjFused& A = a.adj;
      st_first = A.first;
      st_live = A.first ? 1 : A.live[orbit];
      if (!st_live) continue;
      st_step = A.step[orbit];
      st_cost = A.cost[orbit];
      st_status = A.status[orbit];
      st_nit = A.nit[orbit];
      st_T = A.p[0][orbit * 3 + 0];
      st_L = A.p[0][orbit * 3 + 1];
      st_S = A.p[0][orbit * 3 + 2];
      st_g0 = A.gp[0][orbit * 3 + 0];
      st_g1 = A.gp[0][orbit * 3 + 1];
      st_g2 = A.gp[0][orbit * 3 + 2];
      const double* Un = A.u[0] + (size_t)orbit * kPackedDoubles + lane * 2;
      const double* Gn = A.g[0] + (size_t)orbit * kPackedDoubles + lane * 2;
#pragma unroll
      for (int pr = 0; pr < 16; ++pr) {
        cp_async16(pstage + pr * 64, Un + pr * 64);
        cp_async16(pstage + kPackedDoubles + pr * 64, Gn + pr * 64);
      }
      ksw32::cp_async_commit();
      pd.live = 1;
    } else {
      next = decide(orbit + nw);
      if (!pd.live) {  // (fused adjoint descent) this orbit has left the loop: nothing was staged for it
        ksw32::cp_async_wait<0>();
        prefetch(orbit + nw, next);
        continue;
      }
    }
#pragma unroll 1
    for (int it = 0; it < (ADJM ? a.adj.inner : 1); ++it) {
    if constexpr (ADJM) {
      if (!st_live) break;
      pd.step = st_first ? 0.0 : st_step;  // trial point u - step * grad; the first evaluation is at u itself
      pd.T = fma(-pd.step, st_g0, st_T);
      pd.L = fma(-pd.step, st_g1, st_L);
      pd.S = fma(-pd.step, st_g2, st_S);
    }
    const double T = pd.T, Lx = pd.L, S = pd.S;
    const double wbase = -1.0 * (ks::kTwoPi * 32.0 / T);
    const double q = (ks::kTwoPi * 32.0 / Lx) * ((cc + 1) * (1.0 / 32.0));
    const double q2 = q * q, q4 = q2 * q2, d = q4 - q2;
    const double sgq = h ? q : -q;  // d/dx: a' = -q b, b' = +q a
    const double sig = REL ? S / T : 0.0;

    // ---------------- phase A (COL): half tuples of u -> A[t] or B[t] -> tile component h ----------------
    double m1[16], m2[16];  // half tuples (x1, x2) of u; overwritten by F in stage 1
    ksw32::cp_async_wait<0>();
    if constexpr (ADJ || ADJM) {
      // trial point u - step * grad (core.py:1760-1796 increment); slot 2J-1 = x1(J), slot 2J = x2(J), slot 0 = x1(0)
      m2[0] = 0.0;
#pragma unroll
      for (int pr = 0; pr < 16; ++pr) {
        const double2 uu = *reinterpret_cast<const double2*>(pstage + pr * 64);
        const double2 gg = *reinterpret_cast<const double2*>(pstage + kPackedDoubles + pr * 64);
        const double lo = fma(-pd.step, gg.x, uu.x), hi = fma(-pd.step, gg.y, uu.y);
        if (pr == 0) m1[0] = lo; else m2[pr] = lo;
        if (pr < 15) m1[pr + 1] = hi;
      }
    } else {
      static_for<0, 16>([&](auto Jc) {
        constexpr int J = decltype(Jc)::value;
        const bool pres = (J & 1) ? presO : presE;
        m1[J] = pres ? stage[(J == 0 ? 0 : 2 * J - 1) * 32] : 0.0;
        if constexpr (J > 0) m2[J] = pres ? stage[(2 * J) * 32] : 0.0; else m2[J] = 0.0;
      });
    }
    if constexpr (ADJ) {
      if (!a.adj.first) {
        double* Tp = a.adj.u[pd.cur ^ 1] + (size_t)orbit * kPackedDoubles + lane * 2;
#pragma unroll
        for (int pr = 0; pr < 16; ++pr)
          *reinterpret_cast<double2*>(Tp + pr * 64) = make_double2(pr == 0 ? m1[0] : m2[pr], pr < 15 ? m1[pr + 1] : 0.0);
        if (lane == 0) {
          double* tp = a.adj.p[pd.cur ^ 1] + orbit * 3;
          tp[0] = T;
          tp[1] = Lx;
          tp[2] = S;
        }
      }
    }
    if constexpr (!ADJM && !PARK) prefetch(orbit + nw, next);
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
      irfft32(0.0, hr, hi, U);
      double s[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) s[x] = U[x] * U[x];
      double c0;
      rfft32_fwd(s, c0, hr, hi);
      double2* wrow = reinterpret_cast<double2*>(tile) + lane * 15;
#pragma unroll
      for (int k = 1; k < 16; ++k) wrow[k - 1] = make_double2(hr[k], hi[k]);
    }
    __syncwarp();
    // ---------------- phase C (COL): N = 1/2 Dx(.) via the other component; stage 1; W = Dx F via the store ----------------
    double s_ff = 0.0, s_mf = 0.0, s_x = 0.0, s_d = 0.0;
    [[maybe_unused]] double s_nv = 0.0;
    // matvec: the parameter components of v, pre-scaled (-1/T, 1/L, -1/T) as in ks_warp32.cuh; Jacobian mode: unit vector of
    // column e % jac -- state slot (jrow, jcol) of the modes array or parameter slot jpar
    [[maybe_unused]] double vt = 0.0, vx = 0.0, vs = 0.0;
    [[maybe_unused]] int jrow = -1, jcol = -1;
    if constexpr (VOP) {
      int jpar = -1;
      if (unit_v) {
        const int col = orbit % udiv;
        if (col < NM) {
          jrow = col / COLS;
          jcol = col - jrow * COLS;
        } else {
          jpar = col - NM;
        }
      }
      if constexpr (MV) {
        double v3[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int slot = ks::free_slot(a.cmask, i);
          if (unit_v) v3[i] = (slot >= 0 && slot == jpar) ? 1.0 : 0.0;
          else v3[i] = a.v_tail ? (slot < 0 ? 0.0 : a.vmodes[(size_t)orbit * v_ld + NM + slot]) : a.vparams[orbit * 3 + i];
        }
        vt = fix_t ? 0.0 : v3[0] * (-1.0 / T);
        vx = fix_x ? 0.0 : v3[1] * (1.0 / Lx);
        vs = fix_s ? 0.0 : v3[2] * (-1.0 / T);
      }
    }
    // this lane's half tuple J of v
    [[maybe_unused]] auto vhalf = [&](auto Jc, double& v1, double& v2) {
      constexpr int J = decltype(Jc)::value;
      const bool pres = (J & 1) ? presO : presE;
      if (unit_v) {
        const bool mine = pres && colact && jcol == coff;
        v1 = (mine && jrow == J) ? 1.0 : 0.0;
        v2 = (J > 0 && mine && jrow == 15 + J) ? 1.0 : 0.0;
      } else {
        v1 = pres ? park[(J == 0 ? 0 : 2 * J - 1) * 32] : 0.0;
        if constexpr (J > 0) v2 = pres ? park[(2 * J) * 32] : 0.0; else v2 = 0.0;
      }
    };
    {
      double x[32];
      const double scl = 4.0 * sgq * cN;
#pragma unroll
      for (int t = 0; t < 32; ++t) x[t] = scl * tile[(t * 15 + cc) * 2 + (1 - h)];
      double n0, nr[16], ni[16];
      rfft32_fwd(x, n0, nr, ni);
      nr[0] = (0.5 * ks::kSqrt2) * n0;
      ni[0] = 0.0;
      if constexpr (PARK) {  // u's half tuples again (they left the registers after phase A), then release the landing zone
        static_for<0, 16>([&](auto Jc) {
          constexpr int J = decltype(Jc)::value;
          const bool pres = (J & 1) ? presO : presE;
          m1[J] = pres ? stage[(J == 0 ? 0 : 2 * J - 1) * 32] : 0.0;
          if constexpr (J > 0) m2[J] = pres ? stage[(2 * J) * 32] : 0.0; else m2[J] = 0.0;
        });
        prefetch(orbit + nw, next);
      }
      static_for<0, 16>([&](auto Jc) {
        constexpr int J = decltype(Jc)::value;
        const bool pres = (J & 1) ? presO : presE;
        const double w = (J == 0) ? 0.0 : wbase * (J * (1.0 / 32.0));
        const double u1 = m1[J], u2 = m2[J];
        if constexpr (VOP) {
          double v1, v2;
          vhalf(Jc, v1, v2);
          if constexpr (RMV) {  // sums of J^T v's parameter components (ks/orbits.py:2004-2020, :3064-3092); stage 2 acts on v
            s_mf = fma(u1, v1, fma(u2, v2, s_mf));
            s_x = fma(w, fma(u1, v2, -u2 * v1), s_x);
            s_nv = fma(nr[J], v1, fma(ni[J], v2, s_nv));
            if constexpr (REL) {
              const double p1 = __shfl_xor_sync(0xffffffffu, u1, 1), p2 = __shfl_xor_sync(0xffffffffu, u2, 1);
              s_d += sgq * (p1 * v1 + p2 * v2);
            }
            m1[J] = v1;
            m2[J] = v2;
          } else {  // matvec: everything except the u * v product (ks/orbits.py:690-707, :2648-2658)
            const double e = 2.0 * q2 - 4.0 * q4, ce = vx * e, cdt = vt * w;
            double r1 = fma(d, v1, -w * v2), r2 = fma(d, v2, w * v1);
            if constexpr (REL) {
              const double pv1 = __shfl_xor_sync(0xffffffffu, v1, 1), pv2 = __shfl_xor_sync(0xffffffffu, v2, 1);
              r1 = fma(-sig * sgq, pv1, r1);
              r2 = fma(-sig * sgq, pv2, r2);
            }
            r1 += fma(ce, u1, fma(-cdt, u2, -vx * nr[J]));
            r2 += fma(ce, u2, fma(cdt, u1, -vx * ni[J]));
            if constexpr (REL) {
              const double cdxs = (-sig * vt + sig * vx + vs) * sgq;
              const double pu1 = __shfl_xor_sync(0xffffffffu, u1, 1), pu2 = __shfl_xor_sync(0xffffffffu, u2, 1);
              r1 = fma(cdxs, pu1, r1);
              r2 = fma(cdxs, pu2, r2);
            }
            m1[J] = pres ? r1 : 0.0;
            m2[J] = (pres && J > 0) ? r2 : 0.0;
          }
          return;
        }
        double f1 = fma(d, u1, fma(-w, u2, nr[J]));
        double f2 = (J == 0) ? 0.0 : fma(d, u2, fma(w, u1, ni[J]));
        if constexpr (REL) {  // - (S/T) dx(u): a += sig q b, b -= sig q a  (partner's half via shuffle)
          const double p1 = __shfl_xor_sync(0xffffffffu, u1, 1), p2 = __shfl_xor_sync(0xffffffffu, u2, 1);
          f1 = fma(-sig * sgq, p1, f1);
          if constexpr (J > 0) f2 = fma(-sig * sgq, p2, f2);
          if constexpr (GRAD) {  // <dx u, F> pieces: this lane's F half against the partner's u half
            s_d += sgq * (p1 * f1 + p2 * f2);
          }
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
      if constexpr (OP == KS_OP_EQN) {
        double* Fp = a.out_state + (size_t)orbit * NM + coff;
        static_for<0, 16>([&](auto Jc) {
          constexpr int J = decltype(Jc)::value;
          const bool pres = (J & 1) ? presO : presE;
          if (colact && pres) {
            Fp[J * COLS] = m1[J];
            if constexpr (J > 0) Fp[(15 + J) * COLS] = m2[J];
          }
        });
      } else if constexpr (OP == KS_OP_COSTGRAD) {
        if (a.f_out) {
          double* Fp = a.f_out + (size_t)orbit * NM + coff;
          static_for<0, 16>([&](auto Jc) {
            constexpr int J = decltype(Jc)::value;
            const bool pres = (J & 1) ? presO : presE;
            if (colact && pres) {
              Fp[J * COLS] = m1[J];
              if constexpr (J > 0) Fp[(15 + J) * COLS] = m2[J];
            }
          });
        }
      }
      // reductions over the 30 active lanes
      const double lw = colact ? 1.0 : 0.0;
      const double ff = VOP ? 0.0 : warp_sum(s_ff * lw);
      if constexpr (ADJ) {
        if (lane == 0) {
          const KsAdjFused& A = a.adj;
          if (A.first) {  // initial F and cost (optimize.py:462-463); the while-condition cost > tol (:467)
            const double c0 = 0.5 * ff;
            A.cost[orbit] = c0;
            A.tcost[orbit] = -1.0;  // "no trial pending" marker for the next pass
            const int lv = c0 > A.tol ? 1 : 0;
            A.live[orbit] = lv;
            if (lv) atomicAdd(A.n_live, 1);
          } else {
            A.tcost[orbit] = 0.5 * ff;
          }
        }
      } else if constexpr (ADJM) {
        ct = 0.5 * ff;
      } else if constexpr (!VOP) {
        if (lane == 0 && a.out_cost) a.out_cost[orbit] = 0.5 * ff;
      }
      if constexpr (GRAD || RMV) {
        const double e = 2.0 * q2 - 4.0 * q4;
        // with <dx u, F> = s_d (summed over both halves): r_t ~ s_x - sig s_d, r_x ~ (e+d) s_mf + s_x - s_ff, r_s ~ s_d
        // (rmatvec: no identity for <N, v>: r_x ~ e s_mf + sig s_d - <N, v>)
        double px = warp_sum((RMV ? fma(e, s_mf, fma(sig, s_d, -s_nv)) : fma(e + d, s_mf, s_x - s_ff)) * lw);
        double pt = warp_sum(fma(-sig, s_d, s_x) * lw);
        double ps = REL ? warp_sum(s_d * lw) : 0.0;
        if (ADJM || (lane == 0 && (ADJ || a.out_params || a.out_tail))) {
          double gt = fix_t ? 0.0 : (-1.0 / T) * pt, gx = fix_x ? 0.0 : (1.0 / Lx) * px, gs = fix_s ? 0.0 : (-1.0 / T) * ps;
          if constexpr (PC) {
            if (!fix_t) gt *= 1.0 / T;
            if (!fix_x) gx *= 1.0 / ((Lx * Lx) * (Lx * Lx));
          }
          if constexpr (ADJM) {
            tg0 = gt;
            tg1 = gx;
            tg2 = gs;
          } else if constexpr (ADJ) {
            double* gpp = a.adj.gp[a.adj.first ? pd.cur : pd.cur ^ 1] + orbit * 3;
            gpp[0] = gt;
            gpp[1] = gx;
            gpp[2] = gs;
          } else if (a.out_tail) {
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
      if constexpr (OP == KS_OP_EQN) continue;
      if constexpr (PARK) {
#pragma unroll
        for (int J = 0; J < 16; ++J) {
          park[(J == 0 ? 0 : 2 * J - 1) * 32] = m1[J];
          if (J > 0) park[(2 * J) * 32] = m2[J];
        }
      }
      // second inverse: spectrum of F's half -> its field component; Dx F = iq (A_F + i B_F): h0 writes q A_F to .y,
      // h1 writes -q B_F to .x
      double hr[16], hi[16];
      double h0 = m1[0];
      if constexpr (MV) {  // matvec: the second field is v itself (m1 / m2 hold the stage-1 result until phase E)
        double unused;
        vhalf(std::integral_constant<int, 0>{}, h0, unused);
        static_for<1, 16>([&](auto Jc) { vhalf(Jc, hr[decltype(Jc)::value], hi[decltype(Jc)::value]); });
      } else {
#pragma unroll
        for (int J = 1; J < 16; ++J) {
          hr[J] = m1[J];
          hi[J] = m2[J];
        }
      }
      if constexpr (VOP) prefetch_v(orbit + nw);  // v has left its landing zone
      hr[0] = hi[0] = 0.0;
      irfft32(ks::kSqrt2 * h0, hr, hi, x);
      // everything between this store and the load of phase E is linear, so phase E's power-of-two normalisation 4 cP rides
      // along here (exact: same bits as scaling there)
      if constexpr (MV) {
        // this lane read component 1 - h at the top of the phase, which is the component its partner lane stores now
        __syncwarp();
        if (colact) {
#pragma unroll
          for (int t = 0; t < 32; ++t) tile[(t * 15 + c) * 2 + h] = (4.0 * cP) * x[t];
        }
      } else {
        const double sw = (h ? -q : q) * (4.0 * cP);
        if (colact) {
#pragma unroll
          for (int t = 0; t < 32; ++t) tile[(t * 15 + c) * 2 + (1 - h)] = sw * x[t];
        }
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
      double w[32];
      irfft32(0.0, hr, hi, w);
#pragma unroll
      for (int x = 0; x < 32; ++x) w[x] *= U[x];
      double c0;
      rfft32_fwd(w, c0, hr, hi);
      double2* wrow = reinterpret_cast<double2*>(tile) + lane * 15;
#pragma unroll
      for (int k = 1; k < 16; ++k) wrow[k - 1] = make_double2(hr[k], hi[k]);
    }
    __syncwarp();
    // ---------------- phase E (COL): stage 2: grad = -F_t + F_xx + F_xxxx [+ (S/T) F_x] - T(Fx(u w)) ----------------
    {
      double x[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) x[t] = tile[(t * 15 + cc) * 2 + (MV ? 1 - h : h)];  // matvec: Dx(u v) crosses the halves
      double p0, pr[16], pi[16];
      rfft32_fwd(x, p0, pr, pi);
      pr[0] = (0.5 * ks::kSqrt2) * p0;
      pi[0] = 0.0;
      const int out_ld = a.out_ld ? a.out_ld : NM;
      double* Op = (ADJ || ADJM) ? nullptr : a.out_state + (size_t)orbit * out_ld + coff;
      static_for<0, 16>([&](auto Jc) {
        constexpr int J = decltype(Jc)::value;
        const bool pres = (J & 1) ? presO : presE;
        const double w = (J == 0) ? 0.0 : wbase * (J * (1.0 / 32.0));
        double f1, f2;
        if constexpr (PARK) {  // F's half tuples back from their lane-private slots (written by this lane in phase C)
          f1 = park[(J == 0 ? 0 : 2 * J - 1) * 32];
          f2 = (J == 0) ? 0.0 : park[(2 * J) * 32];
        } else {
          f1 = m1[J];
          f2 = m2[J];
        }
        double g1 = fma(d, f1, fma(w, f2, -pr[J]));
        double g2 = (J == 0) ? 0.0 : fma(d, f2, fma(-w, f1, -pi[J]));
        if constexpr (MV) {  // + 2 * 1/2 Dx(u v) (ks/orbits.py:692)
          g1 = fma(sgq, pr[J], f1);
          g2 = (J == 0) ? 0.0 : fma(sgq, pi[J], f2);
        }
        if constexpr (REL && !MV) {  // + (S/T) dx(F)
          const double p1 = __shfl_xor_sync(0xffffffffu, f1, 1), p2 = __shfl_xor_sync(0xffffffffu, f2, 1);
          g1 = fma(sig * sgq, p1, g1);
          if constexpr (J > 0) g2 = fma(sig * sgq, p2, g2);
        }
        if constexpr (PC) {
          const double pm = 1.0 / ((w < 0 ? -w : w) + q2 + q4);
          g1 *= pm;
          g2 *= pm;
        }
        if constexpr (ADJ || ADJM) {  // keep the gradient half tuples (absent halves as zeros) for the packed store below
          m1[J] = pres ? g1 : 0.0;
          m2[J] = pres ? g2 : 0.0;
        } else {
          if (colact && pres) {
            Op[J * COLS] = g1;
            if constexpr (J > 0) Op[(15 + J) * COLS] = g2;
          }
        }
      });
      if constexpr (ADJ) {
        double* Gp = a.adj.g[a.adj.first ? pd.cur : pd.cur ^ 1] + (size_t)orbit * kPackedDoubles + lane * 2;
#pragma unroll
        for (int pr = 0; pr < 16; ++pr)
          *reinterpret_cast<double2*>(Gp + pr * 64) = make_double2(pr == 0 ? m1[0] : m2[pr], pr < 15 ? m1[pr + 1] : 0.0);
      }
      if constexpr (ADJM) {
        // judge this trial now (optimize.py:478-480 backtracking, _process_correction :2174-2189, loop exit :506-507): the
        // same rules as decide() above, on registers; an accepted step rewrites u and g in the warp's stage
        const KsAdjFused& A = a.adj;
        bool take_g = false;
        if (st_first) {  // initial F, cost and gradient (optimize.py:462-467)
          st_first = 0;
          st_cost = ct;
          st_live = ct > A.tol ? 1 : 0;
          take_g = true;
        } else if (ct >= st_cost && st_step > A.min_step) {
          st_step *= 0.5;
        } else {
          int accept = 0;
          st_nit += 1;
          if (ct <= A.tol) {
            st_status = -1;
            accept = 1;
          } else if (st_step <= A.min_step) {
            st_status = 0;
          } else if (st_nit == A.maxiter) {
            st_status = 2;
            accept = (ct < st_cost) ? 1 : 0;
          } else {
            const double big = fmax(fmax(st_cost, ct), 1.0);
            if ((st_cost - ct) / big < A.ftol) st_status = 3; else accept = 1;
          }
          if (accept) {
            st_cost = ct;
#pragma unroll
            for (int pr = 0; pr < 16; ++pr) {
              double2 uu = *reinterpret_cast<const double2*>(pstage + pr * 64);
              const double2 gg = *reinterpret_cast<const double2*>(pstage + kPackedDoubles + pr * 64);
              uu.x = fma(-pd.step, gg.x, uu.x);
              uu.y = fma(-pd.step, gg.y, uu.y);
              *reinterpret_cast<double2*>(pstage + pr * 64) = uu;
            }
            st_T = T;
            st_L = Lx;
            st_S = S;
            take_g = true;
          }
          st_live = (st_cost > A.tol && st_status == -1) ? 1 : 0;
          if (!st_live && st_cost <= A.tol) st_status = -1;
        }
        if (take_g) {
#pragma unroll
          for (int pr = 0; pr < 16; ++pr)
            *reinterpret_cast<double2*>(pstage + kPackedDoubles + pr * 64) =
                make_double2(pr == 0 ? m1[0] : m2[pr], pr < 15 ? m1[pr + 1] : 0.0);
          st_g0 = tg0;
          st_g1 = tg1;
          st_g2 = tg2;
        }
      }
    }
    __syncwarp();  // the tile is rewritten by the next orbit's phase A (other lanes' columns are read in phase E)
    }  // passes of this orbit (one unless ADJM)
    if constexpr (ADJM) {
      const KsAdjFused& A = a.adj;
      double* Uo = A.u[0] + (size_t)orbit * kPackedDoubles + lane * 2;
      double* Go = A.g[0] + (size_t)orbit * kPackedDoubles + lane * 2;
#pragma unroll
      for (int pr = 0; pr < 16; ++pr) {
        *reinterpret_cast<double2*>(Uo + pr * 64) = *reinterpret_cast<const double2*>(pstage + pr * 64);
        *reinterpret_cast<double2*>(Go + pr * 64) = *reinterpret_cast<const double2*>(pstage + kPackedDoubles + pr * 64);
      }
      if (lane == 0) {
        A.cost[orbit] = st_cost;
        A.step[orbit] = st_step;
        A.status[orbit] = st_status;
        A.nit[orbit] = st_nit;
        A.live[orbit] = st_live;
        A.p[0][orbit * 3 + 0] = st_T;
        A.p[0][orbit * 3 + 1] = st_L;
        A.p[0][orbit * 3 + 2] = st_S;
        A.gp[0][orbit * 3 + 0] = st_g0;
        A.gp[0][orbit * 3 + 1] = st_g1;
        A.gp[0][orbit * 3 + 2] = st_g2;
        if (st_live) atomicAdd(A.n_live, 1);
      }
    }
  }
}

}  // namespace ksx1
