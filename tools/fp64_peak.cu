// Measured fp64 peak of the device: dependent-chain-free DFMA / DADD streams from registers, one JSON line per
// configuration.  SURVEY.md section 7 asks for a measured (not computed) fp64 ceiling for the 32x32 kernels; this is it.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OPKIND>  // OPKIND 0 = DFMA, 1 = DADD, 2 = DMUL
__global__ void __launch_bounds__(1024) stream_kernel(double* out, int iters, double seed) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = seed + i + threadIdx.x * 1e-9;
  const double a = 1.0000000001, b = 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (OPKIND == 0) acc[i] = fma(acc[i], a, b);
        else if (OPKIND == 1) acc[i] = acc[i] + b;
        else acc[i] = acc[i] * a;
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;  // keep the chains alive
}

template <int ILP, int OPKIND>
void run(const char* op, int sms, int warps_per_sm, double clock_ghz) {
  const int threads = warps_per_sm * 32, iters = 4096;
  double* out;
  cudaMalloc(&out, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  stream_kernel<ILP, OPKIND><<<sms, threads>>>(out, 64, 1.0);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    stream_kernel<ILP, OPKIND><<<sms, threads>>>(out, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double lane_ops = (double)sms * threads * iters * 16.0 * ILP;
  const double rate = lane_ops / (best * 1e-3);
  printf("{\"op\": \"%s\", \"warps_per_sm\": %d, \"ilp\": %d, \"lane_ops_per_s\": %.4e, \"lanes_per_clk_per_sm\": %.2f, \"ms\": %.3f}\n", op,
         warps_per_sm, ILP, rate, rate / sms / (clock_ghz * 1e9), best);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double ghz = clk_khz * 1e-6;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f}\n", p.name, p.multiProcessorCount, ghz);
  const int sms = p.multiProcessorCount;
  for (int w : {4, 8, 12, 16, 32}) {
    run<1, 0>("DFMA", sms, w, ghz);
    run<2, 0>("DFMA", sms, w, ghz);
    run<4, 0>("DFMA", sms, w, ghz);
    run<8, 0>("DFMA", sms, w, ghz);
  }
  for (int w : {8, 16, 32}) {
    run<8, 1>("DADD", sms, w, ghz);
    run<8, 2>("DMUL", sms, w, ghz);
  }
  return 0;
}
