// Micro-benchmark: throughput of SMALL 1-D bulk copies (cp.async.bulk, SASS UBLKCP) per SM as a function of the copy
// size, for loads (global -> shared, mbarrier completion), stores (shared -> global, bulk group) and reduce-add stores
// (cp.reduce.async.bulk .add.f32).  The element-update kernels move every element tensor with its own bulk copy
// (864 B .. 2 KB): is the TMA unit's operation rate, not bytes, what bounds them?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// mode 0: loads, 1: stores, 2: reduce-add stores, 3: loads from RANDOM block addresses (every copy its own random
// block of `bytes`, like the neighbor gather of the proxy's random topology).  Every CTA issues `ops` copies of `bytes` per round from `lanes` threads.
template <int MODE>
__global__ void __launch_bounds__(128) k(char* g, size_t gbytes, int bytes, int ops, int rounds) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  unsigned char* buf = sm + 128;
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(bar, ops); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  uint32_t phase = 0;
  size_t span = (size_t)bytes * ops;
  size_t nslots = gbytes / span;
  for (int r = 0; r < rounds; ++r) {
    size_t slot = ((size_t)blockIdx.x * 7919 + (size_t)r * 104729) % nslots;
    if (tid < ops) {
      char* gp = g + slot * span + (size_t)tid * bytes;
      unsigned char* sp = buf + (size_t)tid * bytes;
      if (MODE == 3) {
        unsigned long long h = (unsigned long long)blockIdx.x * 0x9E3779B97F4A7C15ull + (unsigned long long)r * 0xC2B2AE3D27D4EB4Full + (unsigned long long)tid * 0x165667B19E3779F9ull;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        gp = g + (h % (gbytes / bytes)) * (size_t)bytes;
      }
      if (MODE == 0 || MODE == 3) {
        mbar_expect(bar, bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sp)), "l"(gp), "r"(bytes), "r"(smem_u32(bar)) : "memory");
      } else if (MODE == 1) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gp), "r"(smem_u32(sp)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      } else {
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gp), "r"(smem_u32(sp)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
    if (MODE == 0 || MODE == 3) { mbar_wait(bar, phase); phase ^= 1; }
    __syncthreads();
  }
  if (MODE != 0 && tid < ops) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int MODE> void run(const char* name, char* g, size_t gbytes, int bytes, int ops, int ctasPerSm) {
  int rounds = 400;
  size_t smem = 128 + (size_t)bytes * ops;
  if (smem * ctasPerSm > 220 * 1024) return;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024 / ctasPerSm));
  size_t req = 220 * 1024 / ctasPerSm - 2048;   // pins the number of resident CTAs per SM
  if (req < smem) req = smem;
  k<MODE><<<148 * ctasPerSm, 128, req>>>(g, gbytes, bytes, ops, 5);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * ctasPerSm, 128, req>>>(g, gbytes, bytes, ops, rounds);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double nops = (double)148 * ctasPerSm * ops * rounds;
  printf("%-7s %6d B x %3d ops/round, %d CTA/SM: %8.3f ms  %8.1f GB/s  %7.1f ns/op/SM (%6.0f cycles @1.9GHz)  err=%d\n", name, bytes, ops,
         ctasPerSm, ms, nops * bytes / ms / 1e6, ms * 1e6 / (nops / 148), ms * 1e6 / (nops / 148) * 1.9, (int)cudaGetLastError());
}
int main() {
  size_t gbytes = (size_t)4 << 30;
  char* g; cudaMalloc(&g, gbytes); cudaMemset(g, 0, gbytes);
  int sizes[] = {256, 864, 2016, 4096, 8192, 24576};
  for (int c : {1, 2, 4}) {
    for (int s : sizes) {
      for (int ops : {1, 4, 16, 28, 64, 112}) {
        if (ops > 128) continue;
        run<0>("load", g, gbytes, s, ops, c);
      }
    }
  }
  for (int c : {2, 4}) for (int s : {256, 864, 2016, 4096}) for (int ops : {28, 64, 112}) run<3>("gather", g, gbytes, s, ops, c);
  for (int s : sizes) for (int ops : {1, 4, 28, 112}) run<1>("store", g, gbytes, s, ops, 2);
  for (int s : sizes) for (int ops : {1, 4, 28, 112}) run<2>("redadd", g, gbytes, s, ops, 2);
  return 0;
}
