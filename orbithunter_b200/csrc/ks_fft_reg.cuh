// In-register complex FFT of compile-time length N (power of two, N <= 64) for one thread.
//
// Replaces the per-line work of scipy.fft.rfft/irfft at the reference call sites
// orbithunter/ks/orbits.py:2245, :2307, :2352, :2409 (and the SR/Anti variants :3505, :3572, :3914, :3999).
// Everything is fully unrolled so the N complex values live in registers (no dynamic indexing); twiddles are
// compile-time constants.  Radix-2 decimation in time; non-trivial butterflies use the 6-FMA "ratio twiddle"
// form (factor out the larger of |cos|,|sin|), so a length-32 transform is 388 fp64 instructions.
// fp64 pipe throughput (64 lanes/clk/SM on B200) is the binding resource of the 32x32 kernels, so the
// instruction count here is what matters.
#pragma once
#include <utility>

#include "ks_platform.cuh"

namespace ksfft {

template <int B, int E, class F>
KS_DEVI void ks_static_for_fft(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    ks_static_for_fft<B + 1, E>(f);
  }
}

// cos(2*pi*k/64), k = 0..16 (one quadrant); everything else by symmetry.
KS_HD constexpr double quarter_cos64(int k) {
  constexpr double C[17] = {1.0,
                            0.995184726672196886244837,
                            0.9807852804032304491261822,
                            0.9569403357322088649357979,
                            0.9238795325112867561281832,
                            0.8819212643483550297127569,
                            0.8314696123025452370787884,
                            0.7730104533627369608109066,
                            0.7071067811865475244008444,
                            0.6343932841636454982151716,
                            0.5555702330196022247428308,
                            0.4713967368259976485563876,
                            0.38268343236508977172846,
                            0.2902846772544623676361924,
                            0.1950903220161282678482849,
                            0.09801714032956060199419556,
                            0.0};
  return C[k];
}
// cos(2*pi*e/64) for any integer e
KS_HD constexpr double cos64(int e) {
  e = ((e % 64) + 64) % 64;
  return e <= 16 ? quarter_cos64(e) : e <= 32 ? -quarter_cos64(32 - e) : e <= 48 ? -quarter_cos64(e - 32) : quarter_cos64(64 - e);
}
KS_HD constexpr double sin64(int e) { return cos64(e - 16); }
KS_HD constexpr double cabs_(double x) { return x < 0 ? -x : x; }
KS_HD constexpr int brev(int x, int bits) {
  int r = 0;
  for (int b = 0; b < bits; ++b) r |= ((x >> b) & 1) << (bits - 1 - b);
  return r;
}
KS_HD constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n / 2); }

// One radix-2 DIT butterfly:  (a, b) <- (a + W b, a - W b),  W = exp(-i * SIGN * 2*pi*E64/64).
template <int E64, int SIGN>
KS_DEVI void bfly(double& ar, double& ai, double& br, double& bi) {
  constexpr int E = ((E64 % 64) + 64) % 64;
  if constexpr (E == 0) {
    const double tr = br, ti = bi;
    br = ar - tr; bi = ai - ti; ar = ar + tr; ai = ai + ti;
  } else if constexpr (E == 16) {
    // W = -i*SIGN :  W b = SIGN * (bi, -br)
    const double tr = (SIGN > 0) ? bi : -bi, ti = (SIGN > 0) ? -br : br;
    br = ar - tr; bi = ai - ti; ar = ar + tr; ai = ai + ti;
  } else {
    constexpr double wr = cos64(E), wi = -SIGN * sin64(E);
    if constexpr (cabs_(wr) >= cabs_(wi)) {
      constexpr double t = wi / wr;
      const double p = fma(-t, bi, br), q = fma(t, br, bi);
      br = fma(-wr, p, ar); bi = fma(-wr, q, ai);
      ar = fma(wr, p, ar);  ai = fma(wr, q, ai);
    } else {
      constexpr double c = wr / wi;
      const double p = fma(c, br, -bi), q = fma(c, bi, br);
      br = fma(-wi, p, ar); bi = fma(-wi, q, ai);
      ar = fma(wi, p, ar);  ai = fma(wi, q, ai);
    }
  }
}

template <int N, int SIGN, int LEN, int K>
KS_DEVI void stage_k(double (&r)[N], double (&i)[N]) {
#pragma unroll
  for (int g = 0; g < N; g += LEN) bfly<K * (64 / LEN), SIGN>(r[g + K], i[g + K], r[g + K + LEN / 2], i[g + K + LEN / 2]);
}
template <int N, int SIGN, int LEN, int... K>
KS_DEVI void stage(double (&r)[N], double (&i)[N], std::integer_sequence<int, K...>) {
  (stage_k<N, SIGN, LEN, K>(r, i), ...);
}
template <int N, int SIGN, int LEN>
KS_DEVI void stages(double (&r)[N], double (&i)[N]) {
  if constexpr (LEN <= N) {
    stage<N, SIGN, LEN>(r, i, std::make_integer_sequence<int, LEN / 2>{});
    stages<N, SIGN, LEN * 2>(r, i);
  }
}

// Unnormalised DFT, natural order in and out:  X[k] = sum_n x[n] exp(-i*SIGN*2*pi*k*n/N).  SIGN=+1 forward, -1 inverse.
template <int N, int SIGN>
KS_DEVI void fft(double (&xr)[N], double (&xi)[N]) {
  static_assert(N >= 2 && N <= 64 && (N & (N - 1)) == 0, "power of two up to 64");
  double r[N], i[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    r[n] = xr[brev(n, ilog2(N))];
    i[n] = xi[brev(n, ilog2(N))];
  }
  stages<N, SIGN, 2>(r, i);
#pragma unroll
  for (int n = 0; n < N; ++n) {
    xr[n] = r[n];
    xi[n] = i[n];
  }
}

// Real-input forward DFT by decimation in time on the real sequence itself (no complex packing / post-processing):
//   X[k] = E[k] + W_N^k O[k],  X[N/2 - k] = conj(E[k] - W_N^k O[k]),  E / O = real DFTs of the even / odd samples.
// Reads x[O + S*n], n = 0..N-1 and writes the non-redundant half X[0..N/2] (yi[0] = yi[N/2] = 0).  Unnormalised,
// forward sign.  6 FMAs per non-trivial butterfly: 164 fp64 instructions for N = 32 (the packed complex-16 route
// costs 234).
template <int N, int S, int O, int M>
KS_DEVI void rfft_dit(const double (&x)[M], double (&yr)[N / 2 + 1], double (&yi)[N / 2 + 1]) {
  static_assert(N >= 2 && (N & (N - 1)) == 0, "power of two");
  if constexpr (N == 2) {
    yr[0] = x[O] + x[O + S];
    yr[1] = x[O] - x[O + S];
    yi[0] = 0.0;
    yi[1] = 0.0;
  } else {
    double er[N / 4 + 1], ei[N / 4 + 1], qr[N / 4 + 1], qi[N / 4 + 1];
    rfft_dit<N / 2, 2 * S, O, M>(x, er, ei);
    rfft_dit<N / 2, 2 * S, O + S, M>(x, qr, qi);
    yr[0] = er[0] + qr[0];
    yr[N / 2] = er[0] - qr[0];
    yi[0] = 0.0;
    yi[N / 2] = 0.0;
    yr[N / 4] = er[N / 4];
    yi[N / 4] = -qr[N / 4];
    ks_static_for_fft<1, N / 4>([&](auto Kc) {
      constexpr int K = decltype(Kc)::value;
      constexpr double wr = cos64(K * (64 / N)), wi = -sin64(K * (64 / N));  // W_N^K
      if constexpr (cabs_(wr) >= cabs_(wi)) {
        constexpr double t = wi / wr;
        const double p = fma(-t, qi[K], qr[K]), q = fma(t, qr[K], qi[K]);
        yr[K] = fma(wr, p, er[K]);
        yi[K] = fma(wr, q, ei[K]);
        yr[N / 2 - K] = fma(-wr, p, er[K]);
        yi[N / 2 - K] = fma(wr, q, -ei[K]);
      } else {
        constexpr double c = wr / wi;
        const double p = fma(c, qr[K], -qi[K]), q = fma(c, qi[K], qr[K]);
        yr[K] = fma(wi, p, er[K]);
        yi[K] = fma(wi, q, ei[K]);
        yr[N / 2 - K] = fma(-wi, p, er[K]);
        yi[N / 2 - K] = fma(wi, q, -ei[K]);
      }
    });
  }
}

}  // namespace ksfft
