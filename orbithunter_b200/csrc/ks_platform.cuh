// Build-mode switch.  Product builds are nvcc for sm_100a.  -DKS_HOST_EMUL (tests/emul only) swaps in the
// CPU-thread shim so the identical kernel sources can be logic-checked in a container without a GPU.
#pragma once
#ifdef KS_HOST_EMUL
#include "cuda_emul.h"
using std::fma;
#define KS_HD inline
#define KS_DEVI inline __attribute__((always_inline))
#else
#include <cuda_runtime.h>
#define KS_HD __host__ __device__ __forceinline__
#define KS_DEVI __device__ __forceinline__
#define KS_DYN_SMEM(type, name)                                  \
  extern __shared__ __align__(16) unsigned char ks_dyn_smem_[];  \
  type* name = reinterpret_cast<type*>(ks_dyn_smem_)
#define KS_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif
