// Random initial conditions on the device: OrbitKS._populate_state (orbithunter/ks/orbits.py:1738-1849, every spatial /
// temporal modulation profile and xshift coupling the reference offers) for a whole batch of integer seeds.
//
// The reference draws randn(n-1, cols) from NumPy's legacy global generator after np.random.seed(seed): MT19937 seeded by
// Knuth's recurrence (numpy/random/src/mt19937/mt19937.c mt19937_seed), 53-bit doubles from two 32-bit draws
// (genrand_res53) and the polar Box-Muller method of legacy-distributions.c (legacy_gauss: every attempt consumes exactly
// four 32-bit words, an accepted attempt yields f*x2 first and f*x1 second).  Third-party code, restated from its
// published algorithm; the 32-bit stream is reproduced bit for bit, the normals to the last ulp or two (device log()).
// One warp per seed: the twist of the 624-word state is done in three data-parallel phases, the 156 attempts of one
// state block are evaluated by the lanes in parallel and compacted in stream order with ballots.
#pragma once
#include <cstdint>

#include "../../include/ks_b200.h"
#include "ks_common.cuh"

namespace kspop {
constexpr int kWarps = 4;
constexpr int kMtN = 624, kMtM = 397;
constexpr int kWarpSmemWords = 2 * kMtN;  // two state buffers
constexpr int kSmemBytes = kWarps * kWarpSmemWords * 4;

struct PopArgs {
  const uint32_t* seeds;  // [B]
  const double* params;   // [B][3]
  double* out;            // [B][rows*cols]
  int batch, n, m, cols;
  // modulation profiles (ks/orbits.py:1797-1842); ids as in include/ks_b200.h; scale / variance overrides, NaN = the
  // reference's default derived from (T, L)
  int smod, tmod, xshift;
  double tscale, xscale, xvar, tvar;
};

KS_DEVI uint32_t twist(uint32_t cur, uint32_t nxt, uint32_t far) {
  const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
  return far ^ (y >> 1) ^ ((nxt & 1u) ? 0x9908b0dfu : 0u);
}
KS_DEVI uint32_t temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
KS_DEVI double res53(uint32_t w0, uint32_t w1) {
  const double a = (double)(w0 >> 5), b = (double)(w1 >> 6);
  return (a * 67108864.0 + b) / 9007199254740992.0;  // exact: a * 2^26 + b < 2^53
}
#ifdef KS_HOST_EMUL
KS_DEVI unsigned ballot(bool p, uint32_t* scratch, int lane) {
  scratch[lane] = p ? 1u : 0u;
  __syncwarp();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= scratch[i] << i;
  __syncwarp();
  return r;
}
KS_DEVI int popc(unsigned x) { return __builtin_popcount(x); }
KS_DEVI double add_rn(double a, double b) { return a + b; }
#else
KS_DEVI unsigned ballot(bool p, uint32_t*, int) { return __ballot_sync(0xffffffffu, p); }
KS_DEVI int popc(unsigned x) { return __popc(x); }
KS_DEVI double add_rn(double a, double b) { return __dadd_rn(a, b); }
#endif
KS_DEVI bool ks_isnan(double v) { return v != v; }
KS_DEVI double warp_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

__global__ void __launch_bounds__(kWarps * 32) ks_populate_kernel(const PopArgs a) {
  KS_DYN_SMEM(uint32_t, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarps + warp;
  if (b >= a.batch) return;  // whole warp
  uint32_t* s0 = smem + warp * kWarpSmemWords;
  uint32_t* s1 = s0 + kMtN;
  const int rows = a.n - 1, cols = a.cols, nm = rows * cols, h = a.n / 2 - 1, k = a.m / 2 - 1;
  double* out = a.out + (size_t)b * nm;
  // mt19937_seed
  if (lane == 0) {
    uint32_t s = a.seeds[b];
    for (int pos = 0; pos < kMtN; ++pos) {
      s0[pos] = s;
      s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)pos + 1u;
    }
  }
  __syncwarp();
  uint32_t* cur = s0;
  uint32_t* nxt = s1;
  int produced = 0;  // normals written so far (same value in every lane)
  double ss = 0.0;   // sum of squares of this lane's normals
  while (produced < nm) {
    // regenerate 624 words: nxt <- twist(cur)
    for (int i = lane; i < kMtN - kMtM; i += 32) nxt[i] = twist(cur[i], cur[i + 1], cur[i + kMtM]);
    __syncwarp();
    for (int i = kMtN - kMtM + lane; i < 2 * (kMtN - kMtM); i += 32) nxt[i] = twist(cur[i], cur[i + 1], nxt[i - (kMtN - kMtM)]);
    __syncwarp();
    for (int i = 2 * (kMtN - kMtM) + lane; i < kMtN - 1; i += 32) nxt[i] = twist(cur[i], cur[i + 1], nxt[i - (kMtN - kMtM)]);
    __syncwarp();
    if (lane == 0) nxt[kMtN - 1] = twist(cur[kMtN - 1], nxt[0], nxt[kMtM - 1]);
    __syncwarp();
    // 156 attempts of four words each, in stream order
    for (int base = 0; base < kMtN / 4 && produced < nm; base += 32) {
      const int t = base + lane;
      bool ok = false;
      double g0 = 0.0, g1 = 0.0;
      if (t < kMtN / 4) {
        const double d1 = res53(temper(nxt[4 * t]), temper(nxt[4 * t + 1]));
        const double d2 = res53(temper(nxt[4 * t + 2]), temper(nxt[4 * t + 3]));
        const double x1 = 2.0 * d1 - 1.0, x2 = 2.0 * d2 - 1.0;
        const double r2 = add_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
        ok = !(r2 >= 1.0 || r2 == 0.0);
        if (ok) {
          const double f = sqrt(-2.0 * log(r2) / r2);
          g0 = __dmul_rn(f, x2);  // returned first
          g1 = __dmul_rn(f, x1);  // cached, returned by the next call
        }
      }
      const unsigned mask = ballot(ok, cur, lane);  // CPU emulation only: `cur` is dead until the buffers swap, used as scratch
      const int pos = produced + 2 * popc(mask & ((1u << lane) - 1u));
      if (ok) {
        if (pos < nm) {
          out[pos] = g0;
          ss = fma(g0, g0, ss);
        }
        if (pos + 1 < nm) {
          out[pos + 1] = g1;
          ss = fma(g1, g1, ss);
        }
      }
      produced += 2 * popc(mask);
    }
    __syncwarp();
    uint32_t* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  __syncwarp();
  const double norm0 = sqrt(warp_sum(ss));
  // modulation (ks/orbits.py:1797-1845) and renormalisation to the pre-modulation norm
  const double T = a.params[b * 3 + 0], L = a.params[b * 3 + 1];
  const double tscale = ks_isnan(a.tscale) ? rint(T / 20.0) : a.tscale;
  const double xscale = ks_isnan(a.xscale) ? rint(L / (ks::kTwoPi * ks::kSqrt2)) : a.xscale;
  const double xvar = ks_isnan(a.xvar) ? fmax(sqrt(xscale), 1.0) : a.xvar, tvar = ks_isnan(a.tvar) ? fmax(sqrt(tscale), 1.0) : a.tvar;
  const double sx = sqrt(xvar), st = sqrt(tvar);
  const double fn = ks::kTwoPi * a.n / ks::kTwoPi, fm = ks::kTwoPi * a.m / ks::kTwoPi, rn = 1.0 / a.n, rm = 1.0 / a.m;
  double s2 = 0.0;
  for (int i = lane; i < nm; i += 32) {
    const int r = i / cols, c = i - r * cols;
    // the reference truncates its floating-point frequency tables to integers (:1788-1791: spatial_frequencies(2 pi, m) =
    // (2 pi m / 2 pi) * (kappa * (1 / m)), temporal likewise, then .astype(int)); for m, n that are not powers of two a few
    // entries land on kappa - 1.  Same expression, same rounding, here.
    const int ji = r <= h ? r : r - h, ki = (c < k ? c : c - k) + 1;
    const double j = floor(__dmul_rn(fn, __dmul_rn((double)ji, rn))), kap = floor(__dmul_rn(fm, __dmul_rn((double)ki, rm)));
    // centre of the spatial profile, coupled to the time index (:1797-1805)
    const double ctr = xscale + (a.xshift == KS_XSHIFT_SQRT ? sqrt(j) : a.xshift == KS_XSHIFT_LOG ? log(1.0 + j)
                                 : a.xshift == KS_XSHIFT_LIN ? j : 0.0);
    double ps = 1.0;
    switch (a.smod) {
      case KS_SMOD_GAUSSIAN: ps = exp(-((kap - ctr) * (kap - ctr)) / (2.0 * xvar)); break;
      case KS_SMOD_LAPLACE: ps = exp(-1.0 * fabs(kap - ctr) / sx); break;
      case KS_SMOD_LAPLACE_SQRT: ps = exp(-1.0 * sqrt(fabs(kap - ctr)) / sx); break;
      case KS_SMOD_PLATEAU_LINEAR: {
        const double q = ks::kTwoPi * kap / L, q2 = q * q;
        ps = kap <= ctr ? 1.0 : 1.0 / (q2 - q2 * q2);
        break;
      }
      case KS_SMOD_EXPONENTIAL_LINEAR: {
        const double q = ks::kTwoPi * kap / L, q2 = q * q;
        ps = exp(q2 - q2 * q2);
        break;
      }
      case KS_SMOD_FLAT_TOP: {
        const double z = fabs(kap - ctr) / xvar, z2 = z * z;
        ps = exp(-(z2 * z2 * z));
        break;
      }
      default: break;
    }
    double pt = 1.0;
    switch (a.tmod) {
      case KS_TMOD_GAUSSIAN: pt = exp(-((j - tscale) * (j - tscale)) / (2.0 * tvar)); break;
      case KS_TMOD_LAPLACE: pt = exp(-1.0 * fabs(j - tscale) / st); break;
      case KS_TMOD_FLAT_TOP: {
        const double z = fabs(j - tscale) / tvar, z2 = z * z;
        pt = exp(-(z2 * z2 * z));
        break;
      }
      case KS_TMOD_LAPLACE_PLATEAU: pt = j < tscale ? 1.0 : exp(-1.0 * fabs(j - tscale) / st); break;
      case KS_TMOD_TRUNCATE: pt = j > tscale ? 0.0 : 1.0; break;
      default: break;
    }
    double v = ps * out[i];
    v = pt * v;
    out[i] = v;
    s2 = fma(v, v, s2);
  }
  const double scale = norm0 / sqrt(warp_sum(s2));
  for (int i = lane; i < nm; i += 32) out[i] *= scale;
}

}  // namespace kspop
