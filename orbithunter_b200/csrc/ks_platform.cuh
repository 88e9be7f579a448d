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
// thread-block clusters (sm_90+): rank / size / barrier / distributed-shared-memory mapping, and the launch with a
// cluster dimension.  tests/emul/cuda_emul.h provides the same names on CPU threads.
#include <cooperative_groups.h>
__device__ __forceinline__ unsigned ks_cluster_rank() { return cooperative_groups::this_cluster().block_rank(); }
__device__ __forceinline__ unsigned ks_cluster_size() { return cooperative_groups::this_cluster().num_blocks(); }
__device__ __forceinline__ void ks_cluster_sync() { cooperative_groups::this_cluster().sync(); }
template <class T>
__device__ __forceinline__ T* ks_cluster_map(T* p, unsigned rank) {
  return cooperative_groups::this_cluster().map_shared_rank(p, rank);
}
template <class... KArgs, class... Args>
inline cudaError_t ks_launch_cluster(void (*kernel)(KArgs...), unsigned grid, unsigned block, unsigned cluster, size_t smem,
                                     cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define KS_LAUNCH_CLUSTER(kernel, grid, block, cluster, smem, stream, ...) \
  ks_launch_cluster(kernel, grid, block, cluster, smem, stream, __VA_ARGS__)
// programmatic dependent launch (sm_90+): the launch processing and CTA scheduling of a kernel overlap the tail of the
// previous kernel in the stream; the kernel itself waits (ks_pdl_wait) before it touches global memory, so stream order
// is preserved.
__device__ __forceinline__ void ks_nanosleep(unsigned ns) { asm volatile("nanosleep.u32 %0;" ::"r"(ns)); }
__device__ __forceinline__ void ks_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void ks_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <class... KArgs, class... Args>
inline cudaError_t ks_launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                                 Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define KS_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) ks_launch_pdl(kernel, grid, block, smem, stream, __VA_ARGS__)
#endif
