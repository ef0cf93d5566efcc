#!/usr/bin/env python3
"""Micro-benchmark 3: instruction-fetch throughput of straight-line FFMA-immediate code as a function of
(a) loop-body size (where do the L0 / L1.5 / L2 instruction tiers end on B200),
(b) resident warps per SM (4..32),
(c) the number of DISTINCT instruction streams executed concurrently by the warps of an SM
    (M = 1, 2, 4: warp w runs body w % M) -- decides whether splitting a column between threads
    (more warps, shorter but different streams) can beat the per-warp streaming limit,
(d) packed FFMA2 with a broadcast immediate.
Emits ifetch3.cu; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ifetch3 ifetch3.cu"""
import random
random.seed(5)
SIZES = [1024, 1536, 2048, 3072, 4096, 6144, 8192, 12288]
MULTI = [(2, 6144), (4, 6144), (4, 3072), (2, 2048), (4, 1024)]
out = ['#include <cstdio>', '#include <cstdlib>', '#include <cuda_runtime.h>', '''
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  ra = *reinterpret_cast<unsigned long long*>(&a);
  rb = *reinterpret_cast<unsigned long long*>(&b);
  rc = *reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}''']


def body(n, indent='    '):
    lines = []
    for i in range(n):
        c = random.uniform(0.5, 1.5)
        lines.append(f'{indent}a[{i % 24}] = fmaf(a[{i % 24}], {c:.7f}f, b);')
    return lines


def kernel_head(name, typ='float'):
    return [f'__global__ void __launch_bounds__(1024, 1) {name}({typ}* o, int reps, long long* cyc) {{',
            f'  {typ} a[24];', '#pragma unroll', '  for (int i = 0; i < 24; ++i) a[i] = o[threadIdx.x + i];',
            f'  const {typ} b = o[blockIdx.x];', '  long long t0 = clock64();', '#pragma unroll 1',
            '  for (int r = 0; r < reps; ++r) {']


def kernel_tail(typ='float'):
    red = ('  float s = 0; for (int i = 0; i < 24; ++i) s += a[i];' if typ == 'float' else
           '  float2 s = make_float2(0.f, 0.f); for (int i = 0; i < 24; ++i) { s.x += a[i].x; s.y += a[i].y; }')
    return ['  }', '  long long t1 = clock64();', red, '  o[blockIdx.x * blockDim.x + threadIdx.x] = s;',
            '  if (threadIdx.x % 32 == 0) atomicMax((unsigned long long*)&cyc[blockIdx.x], (unsigned long long)(t1 - t0));',
            '}']


for n in SIZES:
    out += kernel_head(f'one_{n}') + body(n) + kernel_tail()
for m, n in MULTI:
    out += kernel_head(f'multi_{m}_{n}')
    out.append(f'    switch ((threadIdx.x >> 5) % {m}) {{')
    for j in range(m):
        out.append(f'    case {j}: {{')
        out += body(n, '      ')
        out.append('    } break;')
    out.append('    }')
    out += kernel_tail()
for n in (2048, 12288):
    out += kernel_head(f'pk_{n}', 'float2')
    for i in range(n):
        c = random.uniform(0.5, 1.5)
        out.append(f'    a[{i % 24}] = ffma2(a[{i % 24}], make_float2({c:.7f}f, {c:.7f}f), b);')
    out += kernel_tail('float2')
out.append('''
template <typename K, typename T> void run(K k, const char* name, int n, int warps, T* o, long long* cyc) {
  int grid = 148;
  long long total = 1ll << 20;  // instructions per warp
  int reps = (int)(total / n);
  k<<<grid, warps * 32>>>(o, 2, cyc);   // warm
  cudaMemset(cyc, 0, grid * sizeof(long long));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<<<grid, warps * 32>>>(o, reps, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  double inst = (double)n * reps * warps;  // warp-instructions per SM
  printf("%-16s body %6.0f KB warps/SM %2d : %8.3f ms  IPC/SMSP %.3f cudaErr=%d\\n", name, n * 16 / 1024.0, warps,
         ms, inst / 4.0 / (double)mx, (int)cudaGetLastError());
}
int main() {
  float* o; long long* cyc;
  cudaMalloc(&o, 148 * 8 * 1024 * sizeof(float) + 4096); cudaMemset(o, 0, 148 * 8 * 1024 * sizeof(float) + 4096);
  cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
''')
for n in SIZES:
    for warps in (4, 8, 16, 32):
        out.append(f'  run(one_{n}, "one_{n}", {n}, {warps}, o, cyc);')
for m, n in MULTI:
    for warps in (8, 16, 32):
        out.append(f'  run(multi_{m}_{n}, "multi_{m}_{n}", {n}, {warps}, o, cyc);')
for n in (2048, 12288):
    for warps in (4, 8, 16, 32):
        out.append(f'  run(pk_{n}, "pk_{n}", {n}, {warps}, (float2*)o, cyc);')
out.append('  return 0;\n}')
open('ifetch3.cu', 'w').write('\n'.join(out))
