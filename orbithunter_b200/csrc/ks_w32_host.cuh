// Launch code of the 32x32 register-resident kernels for ONE symmetry class (KS_W32_CLS), included by ks_w32_c{0..3}.cu so
// the four classes compile in parallel.  Everything here has internal linkage except ks_w32_run_c<CLS>.
#pragma once
#include <cstdio>

#include "ks_host.cuh"
#include "ks_warp32.cuh"
#include "ks_warp32x1.cuh"
#include "ks_warp32x3.cuh"
#include "ks_warp32pp.cuh"

#ifndef KS_W32_CLS
#error "define KS_W32_CLS (0..3) before including ks_w32_host.cuh"
#endif

namespace {
using namespace ksh;
constexpr int kCls = KS_W32_CLS;
int warp32_grid(int batch, int warps_per_cta) {
  const int npairs = (batch + 1) / 2;
  const int ctas = (npairs + warps_per_cta - 1) / warps_per_cta;
  const int sms = sm_count();
  return ctas < sms ? ctas : sms;
}
template <int CLS, int OP, bool PC>
int launch_warp32_as_pc(const KsEvalArgs& a, cudaStream_t s) {
  using Cfg = ksw32::Cfg<false>;
  auto kern = ksw32::ks_warp32_kernel<CLS, OP, false, true, PC, false>;
  static DynSmemOptIn optin;
  if (!optin.ensure(kern, Cfg::smem_bytes)) return fail(KS_ERR_LAUNCH, "cudaFuncSetAttribute(warp32 async) failed");
  if (KS_LAUNCH_PDL(kern, (unsigned)warp32_grid(a.batch, Cfg::warps_per_cta), (unsigned)(Cfg::warps_per_cta * 32), (size_t)Cfg::smem_bytes,
                    s, a) != cudaSuccess)
    return fail(KS_ERR_LAUNCH, "warp32 async kernel launch failed");
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "warp32 async kernel launch failed");
}
template <int CLS, int OP>
int launch_warp32_as(const KsEvalArgs& a, cudaStream_t s) {
  if constexpr (OP == KS_OP_COSTGRAD) {
    if (a.precond) return launch_warp32_as_pc<CLS, OP, true>(a, s);
  }
  return launch_warp32_as_pc<CLS, OP, false>(a, s);
}
template <int CLS, int OP, bool PC>
int launch_warp32x1_pc(const KsEvalArgs& a, cudaStream_t s) {
  auto kern = ksx1::ks_warp32x1_kernel<CLS, OP, PC>;
  constexpr int smem = ksx1::smem_bytes(OP == KS_OP_ADJSTEP || OP == KS_OP_ADJMULTI, OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC);
  static DynSmemOptIn optin;
  if (!optin.ensure(kern, smem)) return fail(KS_ERR_LAUNCH, "cudaFuncSetAttribute(warp32x1) failed");
  const int ctas = (a.batch + ksx1::kWarpsPerCta - 1) / ksx1::kWarpsPerCta;
  const int sms = sm_count() * ksx1::kCtasPerSm;
  const cudaError_t ce = KS_LAUNCH_PDL(kern, (unsigned)(ctas < sms ? ctas : sms), (unsigned)(ksx1::kWarpsPerCta * 32), (size_t)smem, s, a);
  if (ce != cudaSuccess) {
    static char msg[160];
    std::snprintf(msg, sizeof(msg), "warp32x1 kernel launch failed: %s", cudaGetErrorString(ce));
    return fail(KS_ERR_LAUNCH, msg);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "warp32x1 kernel launch failed");
}
// variant 5: one orbit per warp, 12 warps per SM (ks_warp32x3.cuh)
template <int CLS, int OP, bool PC>
int launch_warp32x3_pc(const KsEvalArgs& a, cudaStream_t s) {
  auto kern = ksx3::ks_warp32x3_kernel<CLS, OP, PC>;
  static DynSmemOptIn optin;
  if (!optin.ensure(kern, ksx3::kSmemBytes)) return fail(KS_ERR_LAUNCH, "cudaFuncSetAttribute(warp32x3) failed");
  const int ctas = (a.batch + ksx3::kWarpsPerCta - 1) / ksx3::kWarpsPerCta;
  const int sms = sm_count();
  const cudaError_t ce = KS_LAUNCH_PDL(kern, (unsigned)(ctas < sms ? ctas : sms), (unsigned)(ksx3::kWarpsPerCta * 32),
                                       (size_t)ksx3::kSmemBytes, s, a);
  if (ce != cudaSuccess) {
    static char msg[160];
    std::snprintf(msg, sizeof(msg), "warp32x3 kernel launch failed: %s", cudaGetErrorString(ce));
    return fail(KS_ERR_LAUNCH, msg);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "warp32x3 kernel launch failed");
}
// variant 6: variant 4's work split, the two warp groups of a CTA take turns on the fp64 pipe (ks_warp32pp.cuh)
template <int CLS, int OP, bool PC>
int launch_warp32pp_pc(const KsEvalArgs& a, cudaStream_t s) {
  auto kern = kspp::ks_warp32pp_kernel<CLS, OP, PC>;
  static DynSmemOptIn optin;
  if (!optin.ensure(kern, kspp::kSmemBytes)) return fail(KS_ERR_LAUNCH, "cudaFuncSetAttribute(warp32pp) failed");
  const int ctas = (a.batch + kspp::kWarpsPerCta - 1) / kspp::kWarpsPerCta;
  const int sms = sm_count();
  const cudaError_t ce = KS_LAUNCH_PDL(kern, (unsigned)(ctas < sms ? ctas : sms), (unsigned)(kspp::kWarpsPerCta * 32),
                                       (size_t)kspp::kSmemBytes, s, a);
  if (ce != cudaSuccess) {
    static char msg[160];
    std::snprintf(msg, sizeof(msg), "warp32pp kernel launch failed: %s", cudaGetErrorString(ce));
    return fail(KS_ERR_LAUNCH, msg);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "warp32pp kernel launch failed");
}
template <int CLS>
int launch_adjstep(const KsEvalArgs& a, cudaStream_t s) {
  return a.precond ? launch_warp32x1_pc<CLS, KS_OP_ADJSTEP, true>(a, s) : launch_warp32x1_pc<CLS, KS_OP_ADJSTEP, false>(a, s);
}
template <int CLS>
int launch_adjmulti(const KsEvalArgs& a, cudaStream_t s) {
  return a.precond ? launch_warp32x1_pc<CLS, KS_OP_ADJMULTI, true>(a, s) : launch_warp32x1_pc<CLS, KS_OP_ADJMULTI, false>(a, s);
}
template <int CLS, int OP>
int launch_warp32(int variant, const KsEvalArgs& a, cudaStream_t s) {
  if constexpr (OP == KS_OP_EQN || OP == KS_OP_COSTGRAD) {
    if (variant == 6) {
      if constexpr (OP == KS_OP_COSTGRAD) {
        if (a.precond) return launch_warp32pp_pc<CLS, OP, true>(a, s);
      }
      return launch_warp32pp_pc<CLS, OP, false>(a, s);
    }
    if (variant == 5) {
      if constexpr (OP == KS_OP_COSTGRAD) {
        if (a.precond) return launch_warp32x3_pc<CLS, OP, true>(a, s);
      }
      return launch_warp32x3_pc<CLS, OP, false>(a, s);
    }
    if (variant == 4 || variant == 7) {
      if constexpr (OP == KS_OP_COSTGRAD) {
        if (a.precond) return launch_warp32x1_pc<CLS, OP, true>(a, s);
      }
      return launch_warp32x1_pc<CLS, OP, false>(a, s);
    }
  }
  if constexpr (OP == KS_OP_MATVEC || OP == KS_OP_RMATVEC) {
    // variant 4 (default): matvec / rmatvec on the one-orbit-per-warp kernel too (variant 7 keeps them on the two-orbits-per-warp
    // kernel); the column modes (u_div > 0) exist only there
    if (variant == 4 || a.u_div > 0) return launch_warp32x1_pc<CLS, OP, false>(a, s);
  }
  return launch_warp32_as<CLS, OP>(a, s);
}

int w32_run(int op, int variant, const KsEvalArgs& a, cudaStream_t s) {
  switch (op) {
    case KS_OP_EQN: return launch_warp32<kCls, KS_OP_EQN>(variant, a, s);
    case KS_OP_COSTGRAD: return launch_warp32<kCls, KS_OP_COSTGRAD>(variant, a, s);
    case KS_OP_MATVEC: return launch_warp32<kCls, KS_OP_MATVEC>(variant, a, s);
    case KS_OP_RMATVEC: return launch_warp32<kCls, KS_OP_RMATVEC>(variant, a, s);
    case KS_OP_ADJSTEP: return launch_adjstep<kCls>(a, s);
    case KS_OP_ADJMULTI: return launch_adjmulti<kCls>(a, s);
    default: return fail(KS_ERR_ARG, "warp32 path: unknown op");
  }
}
}  // namespace

#define KS_W32_CAT_(a, b) a##b
#define KS_W32_CAT(a, b) KS_W32_CAT_(a, b)
int KS_W32_CAT(ks_w32_run_c, KS_W32_CLS)(int op, int variant, const KsEvalArgs& a, cudaStream_t s) { return w32_run(op, variant, a, s); }
