#!/usr/bin/env python3
"""Micro-benchmark (round 2): ways to feed per-instruction double-precision constants to straight-line DFMA code.
  imm   : literal constants (DFMA + 2 moves per FMA)                                   -- what the f64 kernels do
  imm2  : two accumulator sets per thread, one materialised constant feeds two DFMAs   -- "two elements per thread"
  sm    : constants streamed from a shared-memory table in order of use, one broadcast LDS.128 per two DFMAs
  sm2   : both: LDS.128 per four DFMAs
Emits cbank2.cu; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cbank2 cbank2.cu"""
import random
random.seed(7)
SIZES = [4096]
NA = 24
out = ['#include <cstdio>', '#include <cstdlib>', '#include <cuda_runtime.h>']
consts = [random.uniform(0.5, 1.5) for _ in range(max(SIZES))]
out.append('__device__ double gtab[%d] = {%s};' % (len(consts), ', '.join(repr(c) for c in consts)))


def head(name, two, sm):
    h = [f'__global__ void __launch_bounds__(512, 1) {name}(double* o, int reps, long long* cyc) {{',
         f'  double a[{NA}];', '#pragma unroll', f'  for (int i = 0; i < {NA}; ++i) a[i] = o[threadIdx.x + i];']
    if two:
        h += [f'  double c[{NA}];', '#pragma unroll', f'  for (int i = 0; i < {NA}; ++i) c[i] = o[threadIdx.x + 64 + i];']
    if sm:
        h += [f'  __shared__ double2 stab[{max(SIZES) // 2}];',
              f'  for (int i = threadIdx.x; i < {max(SIZES) // 2}; i += blockDim.x) stab[i] = reinterpret_cast<const double2*>(gtab)[i];',
              '  __syncthreads();', '  const unsigned sbase = (unsigned)__cvta_generic_to_shared(stab);']
    h += ['  const double b = o[blockIdx.x];', '  long long t0 = clock64();', '#pragma unroll 1',
          '  for (int r = 0; r < reps; ++r) {']
    return h


def tail(two):
    t = ['  }', '  long long t1 = clock64();', f'  double s = 0; for (int i = 0; i < {NA}; ++i) s += a[i];']
    if two:
        t += [f'  for (int i = 0; i < {NA}; ++i) s += c[i];']
    t += ['  o[blockIdx.x * blockDim.x + threadIdx.x] = s;',
          '  if (threadIdx.x % 32 == 0) atomicMax((unsigned long long*)&cyc[blockIdx.x], (unsigned long long)(t1 - t0));', '}']
    return t


def body(n, two, sm):
    # n = DFMAs per iteration per thread in every variant
    lines = []
    nconst = n // 2 if two else n
    if sm:
        for j in range(0, nconst, 2):
            lines.append(f'    {{ double2 k; asm volatile("ld.shared.v2.f64 {{%0,%1}}, [%2+{j * 8}];" : "=d"(k.x), "=d"(k.y) : "r"(sbase));')
            for t, comp in ((0, 'x'), (1, 'y')):
                i = (j + t) % NA
                lines.append(f'      a[{i}] = fma(a[{i}], k.{comp}, b);')
                if two:
                    lines.append(f'      c[{i}] = fma(c[{i}], k.{comp}, b);')
            lines.append('    }')
    else:
        for j in range(nconst):
            i = j % NA
            lines.append(f'    a[{i}] = fma(a[{i}], {consts[j]!r}, b);')
            if two:
                lines.append(f'    c[{i}] = fma(c[{i}], {consts[j]!r}, b);')
    return lines


KINDS = [('imm', False, False), ('imm2', True, False), ('sm', False, True), ('sm2', True, True)]
for n in SIZES:
    for kind, two, sm in KINDS:
        out += head(f'{kind}_{n}', two, sm) + body(n, two, sm) + tail(two)
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
    for kind, _, _ in KINDS:
        for warps in (4, 8, 16):
            out.append(f'  run({kind}_{n}, "{kind}_{n}", {n}, {warps}, o, cyc);')
out.append('  return 0;\n}')
open('cbank2.cu', 'w').write('\n'.join(out))
