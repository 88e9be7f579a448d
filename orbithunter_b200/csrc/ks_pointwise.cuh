// Per-tuple algebra of the fused evaluations, shared by the generic (any even size) and tile (power-of-two, multi-pass)
// kernels.  A tuple (a1, a2, b1, b2) at time index j and wavenumber kp+1 is defined in ks_warp32.cuh.
//
//   stage 1 (after the forward transform of u*u):  F = L(u) + N(u,u) (eqn, ks/orbits.py:526-555, :1899-1950), the scalar
//     sums behind cost and the parameter components of J^T v (:1984-2024, Relative :3059-3094), J v's linear and
//     parameter columns (:659-709, :2623-2660), and the spectrum W whose inverse transform is the second field.
//   stage 2 (after the forward transform of second field * u):  J^T v = L^T(v) - (...) (:711-755, :2026-2045,
//     Relative :3096-3115) [+ preconditioner :1034-1089], or J v's 2 N(u, v) term.
#pragma once
#include "ks_common.cuh"
#include "ks_warp32.cuh"  // Tup algebra

namespace kspw {
using ksw32::Tup;
using ksw32::dot_tup;
using ksw32::dx_tup;
using ksw32::rot_t;
using ksw32::rot_x;

struct Geom {
  int n, m, h, k, cols, nm, nmax;
  int log_n, log_m;  // -1 when not a power of two
  double ct, cx;     // 1/sqrt(2n), 1/sqrt(2m)
};

template <int CLS>
KS_DEVI bool has_a(int j) { return CLS == KS_ORBIT || CLS == KS_RELATIVE || (CLS == KS_SHIFT_REFLECTION && (j & 1)); }
template <int CLS>
KS_DEVI bool has_b(int j) {
  return CLS == KS_ORBIT || CLS == KS_RELATIVE || CLS == KS_ANTISYMMETRIC || (CLS == KS_SHIFT_REFLECTION && !(j & 1));
}
template <int CLS>
KS_DEVI Tup gload(const double* __restrict__ p, int j, int kp, const Geom& g) {
  Tup t{0.0, 0.0, 0.0, 0.0};
  if (CLS == KS_ORBIT || CLS == KS_RELATIVE) {
    t.a1 = p[j * g.cols + kp];
    t.b1 = p[j * g.cols + g.k + kp];
    if (j > 0) {
      t.a2 = p[(g.h + j) * g.cols + kp];
      t.b2 = p[(g.h + j) * g.cols + g.k + kp];
    }
  } else if (has_a<CLS>(j)) {
    t.a1 = p[j * g.cols + kp];
    if (j > 0) t.a2 = p[(g.h + j) * g.cols + kp];
  } else {
    t.b1 = p[j * g.cols + kp];
    if (j > 0) t.b2 = p[(g.h + j) * g.cols + kp];
  }
  return t;
}
template <int CLS>
KS_DEVI void gstore(double* __restrict__ p, int j, int kp, const Geom& g, const Tup& t) {
  if (CLS == KS_ORBIT || CLS == KS_RELATIVE) {
    p[j * g.cols + kp] = t.a1;
    p[j * g.cols + g.k + kp] = t.b1;
    if (j > 0) {
      p[(g.h + j) * g.cols + kp] = t.a2;
      p[(g.h + j) * g.cols + g.k + kp] = t.b2;
    }
  } else if (has_a<CLS>(j)) {
    p[j * g.cols + kp] = t.a1;
    if (j > 0) p[(g.h + j) * g.cols + kp] = t.a2;
  } else {
    p[j * g.cols + kp] = t.b1;
    if (j > 0) p[(g.h + j) * g.cols + kp] = t.b2;
  }
}
template <int CLS>
KS_DEVI void gproject(Tup& t, int j) {
  if (!has_a<CLS>(j)) t.a1 = t.a2 = 0.0;
  if (!has_b<CLS>(j)) t.b1 = t.b2 = 0.0;
  if (j == 0) t.a2 = t.b2 = 0.0;
}

// Everything the per-tuple stages need about one orbit (built once per orbit / CTA).
struct OrbitCtx {
  const double* Mo;  // u modes
  const double* Vo;  // v (matvec / rmatvec)
  double* Fo;        // stage-1 state output: F (eqn / costgrad) or J v's first part (matvec)
  double* Oo;        // final output
  double T, Lx, sig, wbase, qbase;
  double vt, vx, vs;  // matvec parameter increments, pre-scaled
  int precond;
};
template <int CLS, int OP>
KS_DEVI OrbitCtx make_ctx(const KsEvalArgs& a, const Geom& g, int orbit, double* tu_scratch) {
  constexpr bool REL = (CLS == KS_RELATIVE);
  const bool fix_t = a.cmask & KS_FIX_T, fix_x = a.cmask & KS_FIX_X, fix_s = (a.cmask & KS_FIX_S) || !REL;
  OrbitCtx c;
  const int v_ld = a.v_ld ? a.v_ld : g.nm, out_ld = a.out_ld ? a.out_ld : g.nm;
  c.Mo = a.modes + (size_t)orbit * g.nm;
  c.Vo = (OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) ? a.vmodes + (size_t)orbit * v_ld : nullptr;
  c.Oo = a.out_state + (size_t)orbit * out_ld;
  c.Fo = (OP == KS_OP_COSTGRAD) ? (a.f_out ? a.f_out + (size_t)orbit * g.nm : tu_scratch) : c.Oo;
  c.T = a.params[orbit * 3 + 0];
  c.Lx = a.params[orbit * 3 + 1];
  const double S = a.params[orbit * 3 + 2];
  c.wbase = -1.0 * (ks::kTwoPi * g.n / c.T);
  c.qbase = ks::kTwoPi * g.m / c.Lx;
  c.sig = REL ? S / c.T : 0.0;
  c.vt = c.vx = c.vs = 0.0;
  c.precond = a.precond;
  if constexpr (OP == KS_OP_MATVEC) {
    double v3[3];
    for (int i = 0; i < 3; ++i) {
      const int slot = ks::free_slot(a.cmask, i);
      v3[i] = a.v_tail ? (slot < 0 ? 0.0 : c.Vo[g.nm + slot]) : a.vparams[orbit * 3 + i];
    }
    c.vt = fix_t ? 0.0 : v3[0] * (-1.0 / c.T);
    c.vx = fix_x ? 0.0 : v3[1] * (1.0 / c.Lx);
    c.vs = fix_s ? 0.0 : v3[2] * (-1.0 / c.T);
  }
  return c;
}
// does stage 1 need the transform of u*u?
template <int OP>
KS_DEVI bool need_nuu(const KsEvalArgs& a) {
  return (OP == KS_OP_EQN || OP == KS_OP_COSTGRAD) || !(a.cmask & KS_FIX_X);
}

struct Sums {
  double ff = 0.0, e = 0.0, x = 0.0, d = 0.0;
};

// P = true-scaled tuple of the forward transform of (u*u); W receives the spectrum of the second inverse transform.
// tuples a stage needs from global memory besides P (loaded by the caller, so that it can batch the loads)
template <int CLS, int OP>
KS_DEVI void stage1_loads(const OrbitCtx& c, const Geom& g, int j, int kp, Tup& Mt, Tup& V) {
  Mt = gload<CLS>(c.Mo, j, kp, g);
  if constexpr (OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) V = gload<CLS>(c.Vo, j, kp, g);
}
template <int CLS, int OP>
KS_DEVI Tup stage2_load(const OrbitCtx& c, const Geom& g, int j, int kp) {
  return (OP == KS_OP_MATVEC || OP == KS_OP_COSTGRAD) ? gload<CLS>(c.Fo, j, kp, g) : gload<CLS>(c.Vo, j, kp, g);
}

// N(u, u) = 1/2 Dx(u u) at (j, kp) from P, the true-scaled tuple of the forward transform of u u (projected on the class)
template <int CLS>
KS_DEVI Tup n_from_p(const OrbitCtx& c, const Geom& g, int j, int kp, const Tup& P) {
  const double qn = 0.5 * (c.qbase * ((kp + 1) * (1.0 / g.m)));
  Tup N{-qn * P.b1, -qn * P.b2, qn * P.a1, qn * P.a2};
  gproject<CLS>(N, j);
  return N;
}
template <int CLS, int OP>
KS_DEVI void stage1_point_n(const OrbitCtx& c, const Geom& g, int j, int kp, const Tup& N, const Tup& Mt, const Tup& V, Tup& W,
                            Sums& s);
template <int CLS, int OP>
KS_DEVI void stage1_point(const OrbitCtx& c, const Geom& g, int j, int kp, const Tup& P, const Tup& Mt, const Tup& V, Tup& W,
                          Sums& s) {
  stage1_point_n<CLS, OP>(c, g, j, kp, n_from_p<CLS>(c, g, j, kp, P), Mt, V, W, s);
}
// the same with N(u, u) given (operator cache of the Krylov drivers, or n_from_p above)
template <int CLS, int OP>
KS_DEVI void stage1_point_n(const OrbitCtx& c, const Geom& g, int j, int kp, const Tup& N, const Tup& Mt, const Tup& V, Tup& W,
                            Sums& s) {
  constexpr bool REL = (CLS == KS_RELATIVE);
  const double w = (j == 0) ? 0.0 : c.wbase * (j * (1.0 / g.n));
  const double q = c.qbase * ((kp + 1) * (1.0 / g.m));
  const double q2 = q * q, q4 = q2 * q2, d = q4 - q2, e = 2.0 * q2 - 4.0 * q4;
  const double sig = c.sig;
  if constexpr (OP == KS_OP_EQN || OP == KS_OP_COSTGRAD) {
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
    gproject<CLS>(F, j);
    const double ff = dot_tup(F, F);
    s.ff += ff;
    gstore<CLS>(c.Fo, j, kp, g, F);
    if constexpr (OP == KS_OP_COSTGRAD) {
      const double mf = dot_tup(Mt, F), xt = w * rot_t(Mt, F);
      s.e += fma(e + d, mf, xt - ff);
      s.x += xt;
      if constexpr (REL) s.d += q * rot_x(Mt, F);
      W = dx_tup(F, q);
    }
  } else if constexpr (OP == KS_OP_RMATVEC) {
    const double dd = REL ? q * rot_x(Mt, V) : 0.0;
    s.e += fma(e, dot_tup(Mt, V), fma(sig, dd, -dot_tup(N, V)));
    s.x += w * rot_t(Mt, V);
    s.d += dd;
    W = dx_tup(V, q);
  } else {
    const double vt = c.vt, vx = c.vx, vs = c.vs;
    const double cdx = REL ? (-sig * vt + sig * vx + vs) * q : 0.0, cdt = vt * w, ce = vx * e;
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
    R.a1 += fma(ce, Mt.a1, fma(-cdt, Mt.a2, fma(-cdx, Mt.b1, -vx * N.a1)));
    R.a2 += fma(ce, Mt.a2, fma(cdt, Mt.a1, fma(-cdx, Mt.b2, -vx * N.a2)));
    R.b1 += fma(ce, Mt.b1, fma(-cdt, Mt.b2, fma(cdx, Mt.a1, -vx * N.b1)));
    R.b2 += fma(ce, Mt.b2, fma(cdt, Mt.b1, fma(cdx, Mt.a2, -vx * N.b2)));
    gproject<CLS>(R, j);
    gstore<CLS>(c.Fo, j, kp, g, R);
    W = V;
  }
}

// totals (t0 = sum ff, t1 = sum e, t2 = sum (x - sig d), t3 = sum d) -> cost and parameter components
template <int CLS, int OP>
KS_DEVI void finish_scalars(const KsEvalArgs& a, const Geom& g, const OrbitCtx& c, int orbit, double t0, double t1, double t2,
                            double t3) {
  constexpr bool REL = (CLS == KS_RELATIVE);
  const bool fix_t = a.cmask & KS_FIX_T, fix_x = a.cmask & KS_FIX_X, fix_s = (a.cmask & KS_FIX_S) || !REL;
  if ((OP == KS_OP_EQN || OP == KS_OP_COSTGRAD) && a.out_cost) a.out_cost[orbit] = 0.5 * t0;
  if ((OP == KS_OP_COSTGRAD || OP == KS_OP_RMATVEC) && (a.out_params || a.out_tail)) {
    double gt = fix_t ? 0.0 : (-1.0 / c.T) * t2, gx = fix_x ? 0.0 : (1.0 / c.Lx) * t1, gs = fix_s ? 0.0 : (-1.0 / c.T) * t3;
    if (OP == KS_OP_COSTGRAD && a.precond) {
      if (!fix_t) gt *= 1.0 / c.T;
      if (!fix_x) gx *= 1.0 / ((c.Lx * c.Lx) * (c.Lx * c.Lx));
    }
    if (a.out_tail) {
      const double g3[3] = {gt, gx, gs};
      for (int i = 0; i < 3; ++i) {
        const int slot = ks::free_slot(a.cmask, i);
        if (slot >= 0) c.Oo[g.nm + slot] = g3[i];
      }
    } else {
      a.out_params[orbit * 3 + 0] = gt;
      a.out_params[orbit * 3 + 1] = gx;
      a.out_params[orbit * 3 + 2] = gs;
    }
  }
}

// P = true-scaled tuple of the forward transform of (second field * u).
// V0 = stage2_load(): F (costgrad), J v's first part (matvec) or v (rmatvec)
template <int CLS, int OP>
KS_DEVI void stage2_point(const OrbitCtx& c, const Geom& g, int j, int kp, const Tup& P, const Tup& V0) {
  constexpr bool REL = (CLS == KS_RELATIVE);
  const double w = (j == 0) ? 0.0 : c.wbase * (j * (1.0 / g.n));
  const double q = c.qbase * ((kp + 1) * (1.0 / g.m));
  const double q2 = q * q, q4 = q2 * q2, d = q4 - q2;
  Tup R;
  if constexpr (OP == KS_OP_MATVEC) {
    const Tup& A0 = V0;
    R = Tup{fma(-q, P.b1, A0.a1), fma(-q, P.b2, A0.a2), fma(q, P.a1, A0.b1), fma(q, P.a2, A0.b2)};
  } else {
    const Tup& V = V0;
    R.a1 = fma(d, V.a1, fma(w, V.a2, -P.a1));
    R.a2 = fma(d, V.a2, fma(-w, V.a1, -P.a2));
    R.b1 = fma(d, V.b1, fma(w, V.b2, -P.b1));
    R.b2 = fma(d, V.b2, fma(-w, V.b1, -P.b2));
    if constexpr (REL) {
      const double sq = c.sig * q;
      R.a1 = fma(-sq, V.b1, R.a1);
      R.a2 = fma(-sq, V.b2, R.a2);
      R.b1 = fma(sq, V.a1, R.b1);
      R.b2 = fma(sq, V.a2, R.b2);
    }
    if (OP == KS_OP_COSTGRAD && c.precond) {
      const double pm = 1.0 / ((w < 0 ? -w : w) + q2 + q4);
      R.a1 *= pm; R.a2 *= pm; R.b1 *= pm; R.b2 *= pm;
    }
  }
  gproject<CLS>(R, j);
  gstore<CLS>(c.Oo, j, kp, g, R);
}

}  // namespace kspw
