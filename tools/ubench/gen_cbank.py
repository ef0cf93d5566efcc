#!/usr/bin/env python3
"""Micro-benchmark: double-precision straight-line FMA code with per-instruction constants.
DFMA has no 64-bit immediate form, so `fma(x, 1.2345, b)` costs DFMA + 2 moves (UMOV/IMAD.MOV) per constant.
Alternative measured here: constants in a __constant__ table with static indices, which ptxas folds into the
DFMA as a constant-bank operand (`DFMA R, R, c[0x3][off], R`): one instruction per FMA, 8 bytes of constant-cache
traffic per instruction.  Emits cbank.cu; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cbank cbank.cu"""
import random
random.seed(7)
SIZES = [2048, 6144]
out = ['#include <cstdio>', '#include <cstdlib>', '#include <cuda_runtime.h>']
consts = [random.uniform(0.5, 1.5) for _ in range(max(SIZES))]
out.append('__constant__ double ctab[%d] = {%s};' % (len(consts), ', '.join(repr(c) for c in consts)))


def head(name):
    return [f'__global__ void __launch_bounds__(512, 1) {name}(double* o, int reps, long long* cyc) {{',
            '  double a[24];', '#pragma unroll', '  for (int i = 0; i < 24; ++i) a[i] = o[threadIdx.x + i];',
            '  const double b = o[blockIdx.x];', '  long long t0 = clock64();', '#pragma unroll 1',
            '  for (int r = 0; r < reps; ++r) {']


def tail():
    return ['  }', '  long long t1 = clock64();', '  double s = 0; for (int i = 0; i < 24; ++i) s += a[i];',
            '  o[blockIdx.x * blockDim.x + threadIdx.x] = s;',
            '  if (threadIdx.x % 32 == 0) atomicMax((unsigned long long*)&cyc[blockIdx.x], (unsigned long long)(t1 - t0));', '}']


for n in SIZES:
    out += head(f'imm_{n}') + [f'    a[{i % 24}] = fma(a[{i % 24}], {consts[i]!r}, b);' for i in range(n)] + tail()
    out += head(f'cb_{n}') + [f'    a[{i % 24}] = fma(a[{i % 24}], ctab[{i}], b);' for i in range(n)] + tail()
out.append('''
template <typename K> void run(K k, const char* name, int n, int warps, double* o, long long* cyc) {
  int grid = 148;
  long long total = 1 << 20;
  int reps = (int)(total / n);
  k<<<grid, warps * 32>>>(o, 2, cyc);
  cudaMemset(cyc, 0, grid * sizeof(long long));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<<<grid, warps * 32>>>(o, reps, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  double fma = (double)n * reps * warps;  // warp-FMAs per SM
  printf("%-10s %5d DFMA/iter warps/SM %2d : %8.3f ms  DFMA/clk/SMSP %.3f (peak 0.5)  TFLOP/s %.1f cudaErr=%d\\n", name, n, warps,
         ms, fma / 4.0 / (double)mx, fma * 64.0 * 148 / (ms * 1e-3) / 1e12, (int)cudaGetLastError());
}
int main() {
  double* o; long long* cyc;
  cudaMalloc(&o, 148 * 8 * 1024 * sizeof(double) + 4096); cudaMemset(o, 0, 148 * 8 * 1024 * sizeof(double) + 4096);
  cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
''')
for n in SIZES:
    for kind in ('imm', 'cb'):
        for warps in (4, 8, 16):
            out.append(f'  run({kind}_{n}, "{kind}_{n}", {n}, {warps}, o, cyc);')
out.append('  return 0;\n}')
open('cbank.cu', 'w').write('\n'.join(out))
