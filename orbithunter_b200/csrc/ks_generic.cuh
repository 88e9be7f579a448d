// Generic-size fused spectral kernels: one CTA per orbit, any even n, m >= 4 (power of two or not).
//
// Same mathematics and tuple algebra as ks_warp32.cuh (see the layout notes there), but the 1-D transforms are done
// by a warp cooperatively on a line held in shared memory (radix-2 for powers of two, direct O(N^2) DFT otherwise,
// which covers the reference fixtures' 26- and 34-point grids), and the intermediate arrays (complex spatial modes
// XC[k][n], field FU[n][m], stage-1 state TU[Nm]) live in a per-CTA slot of the caller's workspace, i.e. in L2.
// This is the path for 64x64 ... 256x256 and for the transform API; it is the correctness baseline that the
// size-specialised kernels are checked against and is not yet tuned (multi-pass through L2, strided transposes).
//
// Reference path replaced: orbithunter/ks/orbits.py transforms :2212-2420, :3472-3579, :3869-4005; eqn :526-555;
// matvec :659-709, :2623-2660; rmatvec :711-755, :1984-2045, :3059-3115; costgrad :757-791; precondition :1034-1089.
#pragma once
#include "ks_common.cuh"
#include "ks_pointwise.cuh"

namespace ksg {
using kspw::Geom;
using kspw::Tup;
using kspw::gload;
using kspw::gproject;
using kspw::gstore;

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
enum { KS_OP_TRANSFORM = 4 };

struct GenArgs {
  KsEvalArgs e;
  Geom g;
  int from_basis, to_basis;  // transform op
  const double* tin;
  double* tout;
  size_t slot_doubles;  // workspace doubles per CTA slot
};

KS_HD size_t slot_doubles(int n, int m, int cols) {
  const size_t k = m / 2 - 1;
  const size_t d = 2 * k * n + (size_t)n * m + (size_t)(n - 1) * cols;  // XC (complex) + FU + TU
  return (d + 31) & ~(size_t)31;  // keep every slot 256-byte aligned (XC is accessed as double2)
}
KS_HD size_t smem_bytes(int n, int m) {
  const int nmax = n > m ? n : m;
  return (size_t)(n + m) * 16 + (size_t)kWarps * 2 * nmax * 16 + kWarps * 8 * 8;
}

KS_DEVI double2 cmul(double2 a, double2 w) { return make_double2(fma(a.x, w.x, -a.y * w.y), fma(a.x, w.y, a.y * w.x)); }

// In-place unnormalised DFT of buf[0..N) by one warp.  tw[j] = exp(-2 pi i j / N).  sign=+1 forward, -1 inverse.
KS_DEVI void line_fft(double2* buf, double2* tmp, int N, int logN, const double2* __restrict__ tw, int sign, int lane) {
  __syncwarp();
  if (logN >= 0) {
    for (int i = lane; i < N; i += 32) {
      int r = 0;
      for (int b = 0; b < logN; ++b) r |= ((i >> b) & 1) << (logN - 1 - b);
      if (i < r) {
        const double2 x = buf[i];
        buf[i] = buf[r];
        buf[r] = x;
      }
    }
    __syncwarp();
    for (int len = 2; len <= N; len <<= 1) {
      const int half = len >> 1, step = N / len;
      for (int b = lane; b < N / 2; b += 32) {
        const int grp = b / half, kx = b - grp * half;
        const int i0 = grp * len + kx, i1 = i0 + half;
        double2 w = tw[kx * step];
        if (sign < 0) w.y = -w.y;
        const double2 x0 = buf[i0], x1 = cmul(buf[i1], w);
        buf[i0] = make_double2(x0.x + x1.x, x0.y + x1.y);
        buf[i1] = make_double2(x0.x - x1.x, x0.y - x1.y);
      }
      __syncwarp();
    }
  } else {
    for (int kx = lane; kx < N; kx += 32) {
      double2 acc = make_double2(0.0, 0.0);
      int idx = 0;
      for (int t = 0; t < N; ++t) {
        double2 w = tw[idx];
        if (sign < 0) w.y = -w.y;
        const double2 p = cmul(buf[t], w);
        acc.x += p.x;
        acc.y += p.y;
        idx += kx;
        if (idx >= N) idx -= N;
      }
      tmp[kx] = acc;
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) buf[i] = tmp[i];
    __syncwarp();
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

// ---- passes (each executed by the whole CTA, warp per line) -----------------------------------------------------
// COL inverse from a tuple source: XC[kp][t] <- true spatial modes of the modes-basis array `src`.
template <int CLS, class TupSrc>
KS_DEVI void pass_col_inverse(const Geom& g, TupSrc src, double2* XC, double2* line, double2* tmp, const double2* tw_n) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int kp = warp; kp < g.k; kp += kWarps) {
    for (int j = lane; j <= g.h; j += 32) {
      const Tup t = src(j, kp);
      if (j == 0) {
        line[0] = make_double2(ks::kSqrt2 * t.a1, ks::kSqrt2 * t.b1);
        line[g.n / 2] = make_double2(0.0, 0.0);
      } else {
        line[j] = make_double2(t.a1 - t.b2, t.a2 + t.b1);
        line[g.n - j] = make_double2(t.a1 + t.b2, t.b1 - t.a2);
      }
    }
    line_fft(line, tmp, g.n, g.log_n, tw_n, -1, lane);
    for (int t = lane; t < g.n; t += 32) XC[(size_t)kp * g.n + t] = make_double2(line[t].x * g.ct, line[t].y * g.ct);
    __syncwarp();
  }
}
// COL forward: per wavenumber, forward transform then `sink(j, kp, P)` with the true-scaled tuple P.
// `respec(j, kp, P) -> bool` may return a tuple to write back as inverse-input spectrum (fused second inverse).
template <int CLS, class Sink>
KS_DEVI void pass_col_forward(const Geom& g, double2* XC, double2* line, double2* tmp, const double2* tw_n, Sink sink,
                              bool reinverse) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int kp = warp; kp < g.k; kp += kWarps) {
    for (int t = lane; t < g.n; t += 32) line[t] = XC[(size_t)kp * g.n + t];
    line_fft(line, tmp, g.n, g.log_n, tw_n, 1, lane);
    for (int j = lane; j <= g.h; j += 32) {
      Tup P;
      if (j == 0) {
        const double c0 = ks::kSqrt2 * g.ct;
        P = Tup{line[0].x * c0, 0.0, line[0].y * c0, 0.0};
      } else {
        const double2 y = line[j], z = line[g.n - j];
        P = Tup{(y.x + z.x) * g.ct, (y.y - z.y) * g.ct, (y.y + z.y) * g.ct, (z.x - y.x) * g.ct};
      }
      Tup W{0.0, 0.0, 0.0, 0.0};
      sink(j, kp, P, W);
      if (reinverse) {
        if (j == 0) {
          line[0] = make_double2(ks::kSqrt2 * W.a1, ks::kSqrt2 * W.b1);
          line[g.n / 2] = make_double2(0.0, 0.0);
        } else {
          line[j] = make_double2(W.a1 - W.b2, W.a2 + W.b1);
          line[g.n - j] = make_double2(W.a1 + W.b2, W.b1 - W.a2);
        }
      }
    }
    if (reinverse) {
      line_fft(line, tmp, g.n, g.log_n, tw_n, -1, lane);
      for (int t = lane; t < g.n; t += 32) XC[(size_t)kp * g.n + t] = make_double2(line[t].x * g.ct, line[t].y * g.ct);
    }
    __syncwarp();
  }
}
// ROW inverse: rows (t, t + n/2) of the field from XC; `emit(t, x, ua, ub)` receives true field values and may
// return a product pair to be transformed forward again (fused), written back to XC.
template <class Emit>
KS_DEVI void pass_row(const Geom& g, double2* XC, double2* line, double2* tmp, const double2* tw_m, bool inverse_first,
                      bool forward_after, Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n2 = g.n / 2;
  for (int t = warp; t < n2; t += kWarps) {
    if (inverse_first) {
      for (int kp = lane; kp < g.k; kp += 32) {
        const double2 A = XC[(size_t)kp * g.n + t], B = XC[(size_t)kp * g.n + t + n2];
        line[kp + 1] = make_double2(A.x - B.y, A.y + B.x);
        line[g.m - kp - 1] = make_double2(A.x + B.y, B.x - A.y);
      }
      if (lane == 0) {
        line[0] = make_double2(0.0, 0.0);
        line[g.m / 2] = make_double2(0.0, 0.0);
      }
      line_fft(line, tmp, g.m, g.log_m, tw_m, -1, lane);
    }
    for (int x = lane; x < g.m; x += 32) {
      double2 v = inverse_first ? make_double2(line[x].x * g.cx, line[x].y * g.cx) : make_double2(0.0, 0.0);
      emit(t, x, v);  // may replace v by the pair to transform forward
      line[x] = v;
    }
    if (forward_after) {
      line_fft(line, tmp, g.m, g.log_m, tw_m, 1, lane);
      for (int kp = lane; kp < g.k; kp += 32) {
        const double2 y = line[kp + 1], z = line[g.m - kp - 1];
        XC[(size_t)kp * g.n + t] = make_double2((y.x + z.x) * g.cx, (y.y - z.y) * g.cx);
        XC[(size_t)kp * g.n + t + n2] = make_double2((y.y + z.y) * g.cx, (z.x - y.x) * g.cx);
      }
    }
    __syncwarp();
  }
}

template <int CLS, int OP>
__global__ void __launch_bounds__(kThreads) ks_generic_kernel(const GenArgs ga) {
  const KsEvalArgs& a = ga.e;
  const Geom g = ga.g;
  KS_DYN_SMEM(double2, smem);
  double2* tw_n = smem;
  double2* tw_m = smem + g.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double2* line = smem + g.n + g.m + (size_t)warp * 2 * g.nmax;
  double2* tmp = line + g.nmax;
  double* red = reinterpret_cast<double*>(smem + g.n + g.m + (size_t)kWarps * 2 * g.nmax);  // [kWarps][8]
  for (int i = threadIdx.x; i < g.n; i += kThreads) {
    double s, c;
    sincospi(2.0 * i / g.n, &s, &c);
    tw_n[i] = make_double2(c, -s);
  }
  for (int i = threadIdx.x; i < g.m; i += kThreads) {
    double s, c;
    sincospi(2.0 * i / g.m, &s, &c);
    tw_m[i] = make_double2(c, -s);
  }
  __syncthreads();
  double* slot = a.scratch + (size_t)blockIdx.x * ga.slot_doubles;
  double2* XC = reinterpret_cast<double2*>(slot);
  double* FU = slot + 2 * (size_t)g.k * g.n;  // even offset: stays 16-byte aligned
  double* TU = FU + (size_t)g.n * g.m;
  const int n2 = g.n / 2;

  for (int orbit = blockIdx.x; orbit < a.batch; orbit += gridDim.x) {
    if constexpr (OP == KS_OP_TRANSFORM) {
      const int fb = ga.from_basis, tb = ga.to_basis;
      const size_t in_size = fb == 0 ? (size_t)g.n * g.m : fb == 1 ? (size_t)g.n * 2 * g.k : (size_t)g.nm;
      const size_t out_size = tb == 0 ? (size_t)g.n * g.m : tb == 1 ? (size_t)g.n * 2 * g.k : (size_t)g.nm;
      const double* in = ga.tin + (size_t)orbit * in_size;
      double* out = ga.tout + (size_t)orbit * out_size;
      // stage 1: reach the complex spatial-mode array XC
      if (fb == 2) {
        pass_col_inverse<CLS>(g, [&](int j, int kp) { return gload<CLS>(in, j, kp, g); }, XC, line, tmp, tw_n);
      } else if (fb == 1) {
        for (int i = threadIdx.x; i < g.n * g.k; i += kThreads) {
          const int t = i / g.k, kp = i - t * g.k;
          XC[(size_t)kp * g.n + t] = make_double2(in[(size_t)t * 2 * g.k + kp], in[(size_t)t * 2 * g.k + g.k + kp]);
        }
      } else {
        pass_row(g, XC, line, tmp, tw_m, false, true, [&](int t, int x, double2& v) {
          v = make_double2(in[(size_t)t * g.m + x], in[(size_t)(t + n2) * g.m + x]);
        });
      }
      __syncthreads();
      // stage 2: leave it
      if (tb == 0) {
        pass_row(g, XC, line, tmp, tw_m, true, false, [&](int t, int x, double2& v) {
          out[(size_t)t * g.m + x] = v.x;
          out[(size_t)(t + n2) * g.m + x] = v.y;
        });
      } else if (tb == 1) {
        for (int i = threadIdx.x; i < g.n * g.k; i += kThreads) {
          const int t = i / g.k, kp = i - t * g.k;
          const double2 v = XC[(size_t)kp * g.n + t];
          out[(size_t)t * 2 * g.k + kp] = v.x;
          out[(size_t)t * 2 * g.k + g.k + kp] = v.y;
        }
      } else {
        pass_col_forward<CLS>(g, XC, line, tmp, tw_n, [&](int j, int kp, Tup P, Tup&) {
          gproject<CLS>(P, j);
          gstore<CLS>(out, j, kp, g, P);
        }, false);
      }
      __syncthreads();
      continue;
    } else {
      const kspw::OrbitCtx c = kspw::make_ctx<CLS, OP>(a, g, orbit, TU);
      const bool need_nuu = kspw::need_nuu<OP>(a);
      // A: u modes -> XC
      pass_col_inverse<CLS>(g, [&](int j, int kp) { return gload<CLS>(c.Mo, j, kp, g); }, XC, line, tmp, tw_n);
      __syncthreads();
      // B: XC -> field u (kept in FU); u^2 -> XC
      pass_row(g, XC, line, tmp, tw_m, true, need_nuu, [&](int t, int x, double2& v) {
        FU[(size_t)t * g.m + x] = v.x;
        FU[(size_t)(t + n2) * g.m + x] = v.y;
        v = make_double2(v.x * v.x, v.y * v.y);
      });
      __syncthreads();
      // C: tuple stage 1 (+ fused second inverse)
      kspw::Sums sm;
      if (!need_nuu) {
        for (int i = threadIdx.x; i < g.k * g.n; i += kThreads) XC[i] = make_double2(0.0, 0.0);
        __syncthreads();
      }
      pass_col_forward<CLS>(g, XC, line, tmp, tw_n, [&](int j, int kp, Tup P, Tup& W) {
        Tup Mt, V{0.0, 0.0, 0.0, 0.0};
        kspw::stage1_loads<CLS, OP>(c, g, j, kp, Mt, V);
        kspw::stage1_point<CLS, OP>(c, g, j, kp, P, Mt, V, W, sm);
      }, OP != KS_OP_EQN);
      // deterministic CTA reduction of the scalar sums
      {
        const double r0 = warp_sum(sm.ff), r1 = warp_sum(sm.e), r2 = warp_sum(fma(-c.sig, sm.d, sm.x)), r3 = warp_sum(sm.d);
        if (lane == 0) {
          red[warp * 8 + 0] = r0;
          red[warp * 8 + 1] = r1;
          red[warp * 8 + 2] = r2;
          red[warp * 8 + 3] = r3;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
          for (int wq = 0; wq < kWarps; ++wq) {
            t0 += red[wq * 8 + 0];
            t1 += red[wq * 8 + 1];
            t2 += red[wq * 8 + 2];
            t3 += red[wq * 8 + 3];
          }
          kspw::finish_scalars<CLS, OP>(a, g, c, orbit, t0, t1, t2, t3);
        }
      }
      __syncthreads();
      if constexpr (OP == KS_OP_EQN) continue;
      // D: second field times u -> XC
      pass_row(g, XC, line, tmp, tw_m, true, true, [&](int t, int x, double2& v) {
        v = make_double2(v.x * FU[(size_t)t * g.m + x], v.y * FU[(size_t)(t + n2) * g.m + x]);
      });
      __syncthreads();
      // E: tuple stage 2
      pass_col_forward<CLS>(g, XC, line, tmp, tw_n, [&](int j, int kp, Tup P, Tup&) {
        kspw::stage2_point<CLS, OP>(c, g, j, kp, P, kspw::stage2_load<CLS, OP>(c, g, j, kp));
      }, false);
      __syncthreads();
    }
  }
}

// OrbitKS.precondition (ks/orbits.py:1034-1089) as a standalone elementwise op: out = in / (|w_j| + q^2 + q^4),
// t <- t * T^-pexp_t, x <- x * L^-pexp_x for unconstrained parameters, s untouched.
struct PrecArgs {
  const double* in;
  const double* gp_in;
  const double* params;
  double* out;
  double* gp_out;
  int batch, n, m, cols;
  unsigned cmask;
  double pexp_t, pexp_x;
};
__global__ void ks_precondition_kernel(const PrecArgs a) {
  const int orbit = blockIdx.y;
  const int nm = (a.n - 1) * a.cols, h = a.n / 2 - 1, k = a.m / 2 - 1;
  const double T = a.params[orbit * 3 + 0], Lx = a.params[orbit * 3 + 1];
  const double wbase = -1.0 * (ks::kTwoPi * a.n / T), qbase = ks::kTwoPi * a.m / Lx;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nm; i += gridDim.x * blockDim.x) {
    const int r = i / a.cols, c = i - r * a.cols;
    const int j = r <= h ? r : r - h, kp = c < k ? c : c - k;
    const double w = wbase * (j * (1.0 / a.n)), q = qbase * ((kp + 1) * (1.0 / a.m));
    const double q2 = q * q;
    a.out[(size_t)orbit * nm + i] = a.in[(size_t)orbit * nm + i] * (1.0 / ((w < 0 ? -w : w) + q2 + q2 * q2));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.gp_out) {
    double gt = a.gp_in[orbit * 3 + 0], gx = a.gp_in[orbit * 3 + 1];
    if (!(a.cmask & KS_FIX_T)) gt *= pow(T, -a.pexp_t);
    if (!(a.cmask & KS_FIX_X)) gx *= pow(Lx, -a.pexp_x);
    a.gp_out[orbit * 3 + 0] = gt;
    a.gp_out[orbit * 3 + 1] = gx;
    a.gp_out[orbit * 3 + 2] = a.gp_in[orbit * 3 + 2];
  }
}

// OrbitKS.dt (ks/orbits.py:379-440) and OrbitKS.dx (:442-524) as standalone elementwise ops on real-packed arrays.
//   axis 0 (time): `in` has rows [0 | Re j=1..h | Im j=1..h] (modes basis); out = swap^order( [0, c1 w^p, c2 w^p] * in )
//   axis 1 (space): `in` has columns [Re k=1..K | Im k=1..K] when half == 0 (spatial_modes, or modes of the full
//                   classes); when half == 1 it has K columns (SR/Anti modes) and only even orders are defined,
//                   using the first K table entries (:505-508).
// (c1, c2) = so2_coefficients(order) (:5213-5236); odd orders swap the real and imaginary blocks (swap_modes :5164).
struct DerivArgs {
  const double* in;
  const double* params;
  double* out;
  int batch, rows, cols, n, m, axis, order, half;
};
__global__ void ks_deriv_kernel(const DerivArgs a) {
  const int orbit = blockIdx.y;
  const int h = a.n / 2 - 1, k = a.m / 2 - 1;
  const size_t sz = (size_t)a.rows * a.cols;
  const double T = a.params[orbit * 3 + 0], Lx = a.params[orbit * 3 + 1];
  const double wbase = -1.0 * (ks::kTwoPi * a.n / T), qbase = ks::kTwoPi * a.m / Lx;
  const int o4 = a.order & 3;
  const double c1 = (o4 == 0 || o4 == 1) ? 1.0 : -1.0, c2 = (o4 == 0 || o4 == 3) ? 1.0 : -1.0;
  const bool odd = a.order & 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int)sz; i += gridDim.x * blockDim.x) {
    const int r = i / a.cols, c = i - r * a.cols;
    double val;
    int dst;
    if (a.axis == 0) {
      if (r == 0) {
        val = 0.0;
        dst = i;
      } else {
        const bool imag = r > h;
        const int j = imag ? r - h : r;
        const double f = pow(wbase * (j * (1.0 / a.n)), (double)a.order);
        val = (imag ? c2 : c1) * f * a.in[orbit * sz + i];
        const int rr = odd ? (imag ? r - h : r + h) : r;
        dst = rr * a.cols + c;
      }
    } else {
      const bool imag = !a.half && c >= k;
      const int kp = imag ? c - k : c;
      const double f = pow(qbase * ((kp + 1) * (1.0 / a.m)), (double)a.order);
      val = (imag ? c2 : c1) * f * a.in[orbit * sz + i];
      const int cc = (odd && !a.half) ? (imag ? c - k : c + k) : c;
      dst = r * a.cols + cc;
    }
    a.out[orbit * sz + dst] = val;
  }
}

}  // namespace ksg
