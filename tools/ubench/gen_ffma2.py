#!/usr/bin/env python3
"""Micro-benchmark 3: throughput of packed FFMA2 (fma.rn.f32x2) with a broadcast 32-bit immediate,
small code (pipe-limited) vs long straight-line code (fetch-limited), at 4..16 warps per SM."""
import random
random.seed(3)
SIZES = [256, 12288]
out = ['#include <cstdio>', '#include <cstdlib>', '#include <cuda_runtime.h>', '''
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  ra = *reinterpret_cast<unsigned long long*>(&a);
  rb = *reinterpret_cast<unsigned long long*>(&b);
  rc = *reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}''']
for n in SIZES:
    out.append(f'__global__ void __launch_bounds__(512) body_{n}(float2* o, int reps, long long* cyc) {{')
    out.append('  float2 a[32];')
    out.append('#pragma unroll')
    out.append('  for (int i = 0; i < 32; ++i) a[i] = o[threadIdx.x + i];')
    out.append('  const float2 b = o[blockIdx.x];')
    out.append('  long long t0 = clock64();')
    out.append('#pragma unroll 1')
    out.append('  for (int r = 0; r < reps; ++r) {')
    for i in range(n):
        c = random.uniform(0.5, 1.5)
        out.append(f'    a[{i % 32}] = ffma2(a[{i % 32}], make_float2({c:.7f}f, {c:.7f}f), b);')
    out.append('  }')
    out.append('  long long t1 = clock64();')
    out.append('  float2 s = make_float2(0.f, 0.f); for (int i = 0; i < 32; ++i) { s.x += a[i].x; s.y += a[i].y; }')
    out.append('  o[blockIdx.x * blockDim.x + threadIdx.x] = s;')
    out.append('  if (threadIdx.x % 32 == 0) atomicMax((unsigned long long*)&cyc[blockIdx.x], (unsigned long long)(t1 - t0));')
    out.append('}')
out.append('''
template <typename K> void run(K k, const char* name, int n, int warps, int ctasPerSm, float2* o, long long* cyc) {
  int grid = 148 * ctasPerSm;
  long long total = 1ll << 21;
  int reps = (int)(total / n);
  k<<<grid, warps * 32>>>(o, 2, cyc);
  cudaMemset(cyc, 0, grid * sizeof(long long));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<<<grid, warps * 32>>>(o, reps, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148 * 8]; cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  double inst = (double)n * reps * warps * ctasPerSm;
  printf("%-12s code %6.0f KB warps/CTA %2d CTAs/SM %d : %8.3f ms  FFMA2 IPC/SMSP %.3f  => %.1f TFLOP/s  cudaErr=%d\\n", name, n * 16 / 1024.0, warps,
         ctasPerSm, ms, inst / 4.0 / (double)mx, inst * 148 * 32 * 4 / (ms * 1e-3) / 1e12, (int)cudaGetLastError());
}
int main() {
  float2* o; long long* cyc;
  cudaMalloc(&o, 148 * 8 * 1024 * sizeof(float2) + 4096); cudaMemset(o, 0, 148 * 8 * 1024 * sizeof(float2) + 4096);
  cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
''')
for n in SIZES:
    for warps, ctas in [(4, 1), (8, 1), (16, 1), (4, 2), (4, 4)]:
        out.append(f'  run(body_{n}, "ffma2_{n}", {n}, {warps}, {ctas}, o, cyc);')
out.append('  return 0;\n}')
open('ffma2.cu', 'w').write('\n'.join(out))
