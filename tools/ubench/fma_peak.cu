// Micro-benchmark: vector-FMA peak of the device (the FP32 / FP64 ceilings of the roofline in bench.py).
// Hot loop (fits the instruction caches) of independent FMA chains, 32 warps per SM, event-timed.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_peak fma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void __launch_bounds__(1024) k(T* o, int reps) {
  T a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = o[threadIdx.x + i];
  const T b = o[blockIdx.x], c = o[blockIdx.x + 1];
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) a[i] = a[i] * b + c;
    }
  }
  T s = 0;
  for (int i = 0; i < ILP; ++i) s += a[i];
  o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename T, int ILP> void run(const char* name, T* o, int sms) {
  const int reps = 4096;
  k<T, ILP><<<sms, 1024>>>(o, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(e0);
    k<T, ILP><<<sms, 1024>>>(o, reps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  double flop = 2.0 * double(sms) * 1024 * reps * 16 * ILP;
  printf("%-28s %8.3f ms  %8.2f TFLOP/s  err=%d\n", name, best, flop / best / 1e9, (int)cudaGetLastError());
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%s, %d SMs, max SM clock %d MHz: nominal FP32 FMA peak = SMs x 128 lanes x 2 x clock = %.1f TFLOP/s\n", p.name,
         p.multiProcessorCount, clk / 1000, p.multiProcessorCount * 128.0 * 2 * clk * 1e3 / 1e12);
  void* o; cudaMalloc(&o, 64 << 20); cudaMemset(o, 0, 64 << 20);
  run<float, 8>("FFMA (3 registers), ILP 8", (float*)o, p.multiProcessorCount);
  run<float, 16>("FFMA (3 registers), ILP 16", (float*)o, p.multiProcessorCount);
  run<double, 8>("DFMA, ILP 8", (double*)o, p.multiProcessorCount);
  run<double, 16>("DFMA, ILP 16", (double*)o, p.multiProcessorCount);
  return 0;
}
