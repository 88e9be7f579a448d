// extern "C" entry point of the device seed generator (see ks_populate.cuh).
#include "ks_populate.cuh"

#include "ks_host.cuh"

#include <cmath>

extern "C" int ks_populate_state_ex(int cls, int n, int m, int batch, const uint32_t* seeds, const double* params,
                                    const KsPopulateOptions* opt, double* modes_out, void* stream) {
  using namespace ksh;
  if (opt && (opt->spatial_modulation < 0 || opt->spatial_modulation > KS_SMOD_FLAT_TOP || opt->temporal_modulation < 0 ||
              opt->temporal_modulation > KS_TMOD_TRUNCATE || opt->xshift < 0 || opt->xshift > KS_XSHIFT_LIN))
    return fail(KS_ERR_ARG, "populate: unknown modulation id");
  if (cls < 0 || cls > 3 || n < 4 || m < 4 || (n % 2) || (m % 2) || n > 4096 || m > 4096)
    return fail(KS_ERR_SHAPE, "unsupported class / discretization (need even n, m >= 4)");
  if (batch < 0) return fail(KS_ERR_ARG, "negative batch");
  if (batch == 0) return 0;
  if (!seeds || !params || !modes_out) return fail(KS_ERR_ARG, "null array argument");
  kspop::PopArgs a;
  a.seeds = seeds;
  a.params = params;
  a.out = modes_out;
  a.batch = batch;
  a.n = n;
  a.m = m;
  a.cols = mode_cols(cls, m);
  a.smod = opt ? opt->spatial_modulation : KS_SMOD_LAPLACE;
  a.tmod = opt ? opt->temporal_modulation : KS_TMOD_GAUSSIAN;
  a.xshift = opt ? opt->xshift : KS_XSHIFT_SQRT;
  a.tscale = opt ? opt->tscale : NAN;
  a.xscale = opt ? opt->xscale : NAN;
  a.xvar = opt ? opt->xvar : NAN;
  a.tvar = opt ? opt->tvar : NAN;
  KS_LAUNCH(kspop::ks_populate_kernel, (batch + kspop::kWarps - 1) / kspop::kWarps, kspop::kWarps * 32, kspop::kSmemBytes,
            static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? 0 : fail(KS_ERR_LAUNCH, "populate kernel launch failed");
}

extern "C" int ks_populate_state(int cls, int n, int m, int batch, const uint32_t* seeds, const double* params, double* modes_out,
                                 void* stream) {
  return ks_populate_state_ex(cls, n, m, batch, seeds, params, nullptr, modes_out, stream);
}
