// Bring-up test of the hand-written tcgen05 layer (seissol_b200/csrc/umma.cuh): one 128 x 64 x 56 GEMM in TF32 and in
// the 3-term split (hi/lo) that the dynamic-rupture interpolation uses, checked against a double-precision host result.
// Tries both readings of the descriptor's leading/stride byte offsets and prints which one is right.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../seissol_b200/csrc -o umma_test umma_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "umma.cuh"
using namespace ssb;

constexpr int M = 128, N = 64, K = 56;
// canonical K-major no-swizzle placement: core matrices contiguous along K
// (core matrices 144 B apart: 16 B of padding, the layout kernels_dr_tc.cuh uses)
__host__ __device__ inline int canon(int r, int k, int kTotal) { return (r / 8) * (kTotal / 4) * 36 + (k / 4) * 36 + (r % 8) * 4 + (k % 4); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(umma::smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(umma::smem_addr(bar)), "r"(parity) : "memory");
}

// variant 0: LBO = K-direction core-matrix stride (128 B), SBO = 8-row-group stride; variant 1: swapped
__device__ __forceinline__ int tid0() { return threadIdx.x; }
__global__ void __launch_bounds__(128) k_test(const float* A, const float* B, float* D, int variant, int split) {
  extern __shared__ __align__(128) unsigned char raw[];
  constexpr int AF = (M / 8) * (K / 4) * 36, BF = (N / 8) * (K / 4) * 36;   // floats per padded operand
  float* sAh = reinterpret_cast<float*>(raw);
  float* sAl = sAh + AF;
  float* sBh = sAl + AF;
  float* sBl = sBh + BF;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBl + BF);
  for (int i = tid0(); i < 2 * AF + 2 * BF; i += 128) {
    sAh[i] = 0.f;
  }
  __syncthreads();
  uint32_t* tmemBase = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < M * K; i += 128) {
    const int r = i / K, k = i % K;
    const float x = A[i], h = umma::tf32_hi(x);
    sAh[canon(r, k, K)] = split ? h : x;
    sAl[canon(r, k, K)] = umma::tf32_lo(x, h);
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    const float x = B[i], h = umma::tf32_hi(x);
    sBh[canon(r, k, K)] = split ? h : x;
    sBl[canon(r, k, K)] = umma::tf32_lo(x, h);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    umma::tmem_alloc(tmemBase, 64);
  }
  // generic-proxy writes to shared memory -> visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmemBase;
  if (tid == 0) {
    const uint32_t kStride = 144, rowStride = (K / 4) * 144;
    const uint32_t lbo = variant == 0 ? kStride : rowStride, sbo = variant == 0 ? rowStride : kStride;
    const uint32_t idesc = umma::instr_desc_tf32(M, N);
    bool acc = false;
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t off = ks * 2 * kStride;
      const uint64_t ah = umma::smem_desc(umma::smem_addr(sAh) + off, lbo, sbo), al = umma::smem_desc(umma::smem_addr(sAl) + off, lbo, sbo);
      const uint64_t bh = umma::smem_desc(umma::smem_addr(sBh) + off, lbo, sbo), bl = umma::smem_desc(umma::smem_addr(sBl) + off, lbo, sbo);
      umma::mma_tf32(tmem, ah, bh, idesc, acc);
      acc = true;
      if (split) {
        umma::mma_tf32(tmem, ah, bl, idesc, true);
        umma::mma_tf32(tmem, al, bh, idesc, true);
      }
    }
    umma::commit(bar);
  }
  mbar_wait(bar, 0);
  umma::fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    umma::tmem_ld8(tmem + ((warp * 32u) << 16) + c0, v);
    umma::tmem_ld_wait();
    for (int j = 0; j < 8; ++j) {
      D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    umma::tmem_dealloc(tmem, 64);
  }
}

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1);
  for (auto& x : A) x = (rand() / (float)RAND_MAX) * 2 - 1;
  for (auto& x : B) x = (rand() / (float)RAND_MAX) * 2 - 1;
  std::vector<double> ref(M * N, 0.0);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; ref[m * N + n] = s; }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (2 * (M / 8) * (K / 4) * 36 + 2 * (N / 8) * (K / 4) * 36) * 4 + 64;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int variant = 0; variant < 2; ++variant) {
    for (int split = 0; split < 2; ++split) {
      cudaMemset(dD, 0xff, D.size() * 4);
      k_test<<<1, 128, smem>>>(dA, dB, dD, variant, split);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double num = 0, den = 0, mx = 0;
      for (int i = 0; i < M * N; ++i) { double d = D[i] - ref[i]; num += d * d; den += ref[i] * ref[i]; mx = fmax(mx, fabs(d)); }
      printf("variant %d (%s) %s: rel L2 error %.3e  max abs %.3e  cuda=%s\n", variant,
             variant == 0 ? "LBO = K stride, SBO = row-group stride" : "LBO = row-group stride, SBO = K stride",
             split ? "3 x TF32 split" : "plain TF32    ", sqrt(num / den), mx, cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
