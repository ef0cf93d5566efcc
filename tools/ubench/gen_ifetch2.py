#!/usr/bin/env python3
"""Micro-benchmark 2: does a block barrier every K instructions (warps forced into a narrow code
window) let the warps of an SM share instruction-cache lines of long straight-line code?"""
import random
random.seed(2)
N = 12288
KS = [0, 128, 256, 512, 1024, 2048]
out = ['#include <cstdio>', '#include <cstdlib>', '#include <cuda_runtime.h>']
for K in KS:
    out.append(f'__global__ void __launch_bounds__(1024) body_k{K}(float* o, int reps, long long* cyc) {{')
    out.append('  float a[32];')
    out.append('#pragma unroll')
    out.append('  for (int i = 0; i < 32; ++i) a[i] = o[threadIdx.x + i];')
    out.append('  const float b = o[blockIdx.x];')
    out.append('  long long t0 = clock64();')
    out.append('#pragma unroll 1')
    out.append('  for (int r = 0; r < reps; ++r) {')
    for i in range(N):
        c = random.uniform(0.5, 1.5)
        out.append(f'    a[{i % 32}] = fmaf(a[{i % 32}], {c:.7f}f, b);')
        if K and (i + 1) % K == 0:
            out.append('    __syncthreads();')
    out.append('  }')
    out.append('  long long t1 = clock64();')
    out.append('  float s = 0; for (int i = 0; i < 32; ++i) s += a[i];')
    out.append('  o[blockIdx.x * blockDim.x + threadIdx.x] = s;')
    out.append('  if (threadIdx.x % 32 == 0) atomicMax((unsigned long long*)&cyc[blockIdx.x], (unsigned long long)(t1 - t0));')
    out.append('}')
out.append('''
template <typename K> void run(K k, const char* name, int n, int warps, int ctasPerSm, float* o, long long* cyc) {
  int grid = 148 * ctasPerSm;
  int reps = 128;
  cudaMemset(cyc, 0, grid * sizeof(long long));
  k<<<grid, warps * 32>>>(o, 2, cyc);   // warm
  cudaMemset(cyc, 0, grid * sizeof(long long));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<<<grid, warps * 32>>>(o, reps, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148 * 8]; cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  double inst = (double)n * reps * warps * ctasPerSm;  // warp-instructions per SM
  printf("%-12s warps/CTA %2d CTAs/SM %d : %8.3f ms  IPC/SMSP %.3f cudaErr=%d\\n", name, warps,
         ctasPerSm, ms, inst / 4.0 / (double)mx, (int)cudaGetLastError());
}
int main() {
  float* o; long long* cyc;
  cudaMalloc(&o, 148 * 8 * 1024 * sizeof(float) + 4096); cudaMemset(o, 0, 148 * 8 * 1024 * sizeof(float) + 4096);
  cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
''')
for K in KS:
    for warps, ctas in [(8, 1), (16, 1), (32, 1), (8, 2), (16, 2)]:
        out.append(f'  run(body_k{K}, "bar_every_{K}", {N}, {warps}, {ctas}, o, cyc);')
out.append('  return 0;\n}')
open('ifetch2.cu', 'w').write('\n'.join(out))
