// Tuning microbenchmark: how close to the fp64 pipe peak does the register-resident FFT instruction stream of the 32x32
// kernels get by itself (no memory traffic), as a function of resident warps per SM?  One JSON line per configuration.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I orbithunter_b200/csrc -I include -o tools/bin/fft_stream tools/fft_stream.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "ks_warp32x1.cuh"

// KIND 0: irfft32 -> rfft32 round trips in registers (the row-phase pattern without the product)
// KIND 1: the same plus a shared-memory exchange of the 30 spectrum values per lane between the two transforms
template <int KIND, int MAXT, int UNROLL>
__global__ void __launch_bounds__(MAXT, 1) fft_stream_kernel(double* out, int iters, double seed) {
  extern __shared__ double sm[];
  double hr[16], hi[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    hr[k] = seed * (k + 1) + threadIdx.x * 1e-3;
    hi[k] = seed * (k + 2) - threadIdx.x * 1e-3;
  }
  hr[0] = hi[0] = 0.0;
  double2* row = reinterpret_cast<double2*>(sm) + (threadIdx.x >> 5) * 480 + (threadIdx.x & 31) * 15;
#pragma unroll UNROLL
  for (int it = 0; it < iters; ++it) {
    double x[32];
    ksx1::irfft32(0.0, hr, hi, x);
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] *= 0.03125;
    double c0;
    ksx1::rfft32_fwd(x, c0, hr, hi);
    if (KIND == 1) {
#pragma unroll
      for (int k = 1; k < 16; ++k) row[k - 1] = make_double2(hr[k], hi[k]);
      __syncwarp();
      const double2* r2 = reinterpret_cast<const double2*>(sm) + (threadIdx.x >> 5) * 480 + ((threadIdx.x + 1) & 31) * 15;
#pragma unroll
      for (int k = 1; k < 16; ++k) {
        const double2 v = r2[k - 1];
        hr[k] = v.x;
        hi[k] = v.y;
      }
      __syncwarp();
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += hr[k] + hi[k];
  if (s == 12345.678) out[0] = s;
}

template <int KIND, int MAXT, int UNROLL>
void run(int sms, int warps, double ghz, int fp64_per_iter) {
  double* out;
  cudaMalloc(&out, 8);
  auto k = fft_stream_kernel<KIND, MAXT, UNROLL>;
  const int smem = warps * 480 * 16;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 2016;
  k<<<sms, warps * 32, smem>>>(out, 10, 1.0);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<<<sms, warps * 32, smem>>>(out, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double lane_ops = (double)sms * warps * 32 * iters * fp64_per_iter;
  printf("{\"kind\": %d, \"unroll\": %d, \"warps_per_sm\": %d, \"max_threads\": %d, \"fp64_lanes_per_clk_per_sm\": %.2f, \"ms\": %.3f, \"err\": \"%s\"}\n", KIND, UNROLL, warps,
         MAXT, lane_ops / (best * 1e-3) / sms / (ghz * 1e9), best, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double ghz = clk_khz * 1e-6;
  const int sms = p.multiProcessorCount;
  const int per_iter = 233 + 164 + 32;  // irfft32 + rfft32 + scaling (SASS counts, profiles/)
  run<0, 256, 1>(sms, 4, ghz, per_iter);
  run<0, 256, 1>(sms, 8, ghz, per_iter);
  run<0, 384, 1>(sms, 12, ghz, per_iter);
  run<1, 256, 1>(sms, 4, ghz, per_iter);
  run<1, 256, 1>(sms, 8, ghz, per_iter);
  run<1, 384, 1>(sms, 12, ghz, per_iter);
  // the same loop unrolled: 4 x and 7 x the code (7 x 430 instructions = 48 KB, the size of one fully unrolled evaluation)
  run<1, 256, 4>(sms, 8, ghz, per_iter);
  run<1, 384, 4>(sms, 12, ghz, per_iter);
  run<1, 256, 7>(sms, 4, ghz, per_iter);
  run<1, 256, 7>(sms, 8, ghz, per_iter);
  run<1, 384, 7>(sms, 12, ghz, per_iter);
  run<0, 256, 7>(sms, 8, ghz, per_iter);
  run<0, 384, 7>(sms, 12, ghz, per_iter);
  return 0;
}
