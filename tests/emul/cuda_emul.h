// Minimal "CUDA on CPU threads" shim -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// The build container has nvcc but no GPU.  To catch indexing / scaling / synchronisation-logic bugs before
// spending GPU minutes, tests/emul/build_emul.py compiles the *same* kernel sources (orbithunter_b200/csrc/*.cu*)
// with g++ and -DKS_HOST_EMUL; this header supplies just enough of the CUDA programming model for them:
// one OS thread per CUDA thread, blocks executed one after another, __syncthreads/__syncwarp as barriers,
// warp shuffles through a per-warp exchange buffer, dynamic shared memory as a per-block heap buffer.
// It proves nothing about performance and is not a fallback: orbithunter_b200 only ever loads the nvcc-built library.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __shared__ static  /* blocks run one after another, so one static instance emulates per-block shared memory */
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__ __restrict

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emul"; }
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  std::memcpy(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
  std::memset(p, v, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsyncD2D(void* d, const void* s, size_t n, cudaStream_t) {
  std::memcpy(d, s, n);
  return cudaSuccess;
}

namespace ks_emul {
struct BlockCtx {
  unsigned nthreads;
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint64_t> shfl;  // one 64-bit slot per thread
  unsigned char* smem;
};
struct ThreadCtx {
  dim3 tid, bid, bdim, gdim;
  BlockCtx* blk;
  // thread-block cluster (launch_cluster): the blocks of one cluster run concurrently
  unsigned cluster_rank = 0, cluster_size = 1;
  BlockCtx* const* cluster_blocks = nullptr;
  std::barrier<>* cluster_bar = nullptr;
};
inline thread_local ThreadCtx tctx;
inline int device_sm_count = 4;  // tiny "GPU" so persistent loops iterate
}  // namespace ks_emul

#define threadIdx (ks_emul::tctx.tid)
#define blockIdx (ks_emul::tctx.bid)
#define blockDim (ks_emul::tctx.bdim)
#define gridDim (ks_emul::tctx.gdim)

static inline void __syncthreads() { ks_emul::tctx.blk->block_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  ks_emul::tctx.blk->warp_bar[threadIdx.x >> 5]->arrive_and_wait();
}
static inline double __shfl_xor_sync(unsigned, double v, int lane_mask) {
  auto* b = ks_emul::tctx.blk;
  unsigned t = threadIdx.x;
  std::memcpy(&b->shfl[t], &v, 8);
  b->warp_bar[t >> 5]->arrive_and_wait();
  double r;
  std::memcpy(&r, &b->shfl[t ^ (unsigned)lane_mask], 8);
  b->warp_bar[t >> 5]->arrive_and_wait();
  return r;
}
static inline double __shfl_sync(unsigned, double v, int src_lane) {
  auto* b = ks_emul::tctx.blk;
  unsigned t = threadIdx.x;
  std::memcpy(&b->shfl[t], &v, 8);
  b->warp_bar[t >> 5]->arrive_and_wait();
  double r;
  std::memcpy(&r, &b->shfl[(t & ~31u) | (unsigned)src_lane], 8);
  b->warp_bar[t >> 5]->arrive_and_wait();
  return r;
}
static inline int __shfl_sync(unsigned m, int v, int src_lane) { return (int)__shfl_sync(m, (double)v, src_lane); }
using std::pow;
using std::fmax;
using std::fabs;
using std::sqrt;
static inline double __ldg(const double* p) { return *p; }
static inline void ks_nanosleep(unsigned) {}
static inline void sincospi(double x, double* s, double* c) {
  const long double a = 3.14159265358979323846264338327950288L * (long double)x;
  *s = (double)sinl(a);
  *c = (double)cosl(a);
}
static inline double fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline int atomicAdd(int* p, int v) {
  return reinterpret_cast<std::atomic<int>*>(p)->fetch_add(v);
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) {
  return reinterpret_cast<std::atomic<unsigned>*>(p)->fetch_add(v);
}

// dynamic shared memory: kernels declare `extern __shared__ double2 smem[]` through KS_DYN_SMEM(type, name)
#define KS_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(ks_emul::tctx.blk->smem)

namespace ks_emul {
template <class Kernel, class... Args>
void launch(Kernel kernel, dim3 grid, dim3 block, size_t smem_bytes, Args... args) {
  unsigned nthreads = block.x * block.y * block.z;
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      BlockCtx blk;
      blk.nthreads = nthreads;
      blk.block_bar = std::make_unique<std::barrier<>>(nthreads);
      unsigned nwarps = (nthreads + 31) / 32;
      for (unsigned w = 0; w < nwarps; ++w) {
        unsigned cnt = std::min(32u, nthreads - 32 * w);
        blk.warp_bar.emplace_back(std::make_unique<std::barrier<>>(cnt));
      }
      blk.shfl.assign(nwarps * 32, 0);
      std::vector<unsigned char> smem(smem_bytes + 64);
      blk.smem = smem.data() + ((64 - (reinterpret_cast<uintptr_t>(smem.data()) & 63)) & 63);
      std::vector<std::thread> ts;
      ts.reserve(nthreads);
      for (unsigned t = 0; t < nthreads; ++t) {
        ts.emplace_back([&, t]() {
          tctx.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
          tctx.bid = dim3(bx, by, 0);
          tctx.bdim = block;
          tctx.gdim = grid;
          tctx.blk = &blk;
          kernel(args...);
        });
      }
      for (auto& th : ts) th.join();
    }
}
}  // namespace ks_emul

namespace ks_emul {
// grid.x must be a multiple of `cluster`; the `cluster` blocks of one cluster run concurrently (one OS thread per CUDA
// thread), clusters one after another.  Kernels launched this way must keep ALL shared memory in the dynamic buffer
// (KS_DYN_SMEM): `__shared__` statics are a single instance per process here.
template <class Kernel, class... Args>
void launch_cluster(Kernel kernel, unsigned grid, unsigned block, unsigned cluster, size_t smem_bytes, Args... args) {
  const unsigned nwarps = (block + 31) / 32;
  for (unsigned c0 = 0; c0 < grid; c0 += cluster) {
    std::vector<BlockCtx> blks(cluster);
    std::vector<BlockCtx*> ptrs(cluster);
    std::vector<std::vector<unsigned char>> mem(cluster);
    for (unsigned r = 0; r < cluster; ++r) {
      BlockCtx& blk = blks[r];
      blk.nthreads = block;
      blk.block_bar = std::make_unique<std::barrier<>>(block);
      for (unsigned w = 0; w < nwarps; ++w) blk.warp_bar.emplace_back(std::make_unique<std::barrier<>>(std::min(32u, block - 32 * w)));
      blk.shfl.assign(nwarps * 32, 0);
      mem[r].resize(smem_bytes + 64);
      blk.smem = mem[r].data() + ((64 - (reinterpret_cast<uintptr_t>(mem[r].data()) & 63)) & 63);
      ptrs[r] = &blk;
    }
    std::barrier<> cbar(cluster * block);
    std::vector<std::thread> ts;
    ts.reserve(cluster * block);
    for (unsigned r = 0; r < cluster; ++r)
      for (unsigned t = 0; t < block; ++t)
        ts.emplace_back([&, r, t]() {
          tctx.tid = dim3(t);
          tctx.bid = dim3(c0 + r);
          tctx.bdim = dim3(block);
          tctx.gdim = dim3(grid);
          tctx.blk = ptrs[r];
          tctx.cluster_rank = r;
          tctx.cluster_size = cluster;
          tctx.cluster_blocks = ptrs.data();
          tctx.cluster_bar = &cbar;
          kernel(args...);
          cbar.arrive_and_drop();  // a block that returned no longer takes part in cluster barriers
        });
    for (auto& th : ts) th.join();
  }
}
}  // namespace ks_emul
static inline unsigned ks_cluster_rank() { return ks_emul::tctx.cluster_rank; }
static inline unsigned ks_cluster_size() { return ks_emul::tctx.cluster_size; }
static inline void ks_cluster_sync() {
  if (ks_emul::tctx.cluster_bar) ks_emul::tctx.cluster_bar->arrive_and_wait();
}
template <class T>
static inline T* ks_cluster_map(T* p, unsigned rank) {
  auto& c = ks_emul::tctx;
  if (!c.cluster_blocks) return p;
  const size_t off = reinterpret_cast<unsigned char*>(p) - c.blk->smem;
  return reinterpret_cast<T*>(c.cluster_blocks[rank]->smem + off);
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

#define KS_LAUNCH(kernel, grid, block, smem, stream, ...) ks_emul::launch(kernel, dim3(grid), dim3(block), smem, __VA_ARGS__)
static inline void ks_pdl_wait() {}
static inline void ks_pdl_launch_dependents() {}
#define KS_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) \
  (ks_emul::launch(kernel, dim3(grid), dim3(block), smem, __VA_ARGS__), cudaSuccess)
#define KS_LAUNCH_CLUSTER(kernel, grid, block, cluster, smem, stream, ...) \
  (ks_emul::launch_cluster(kernel, grid, block, cluster, smem, __VA_ARGS__), cudaSuccess)
