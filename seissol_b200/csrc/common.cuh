// Shared device helpers: TMA bulk copies + mbarriers (sm_100a inline PTX), 128-bit
// shared-memory vector access, and the per-element array descriptor used at the
// C-ABI boundary.  Hand-written for sm_100a; there is no fallback path.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ssb {

// ---------------------------------------------------------------------------------------
// Boundary descriptor: one logical array with one entry per element.  Either a device
// table of device pointers (the reference's batched `real**` interface,
// src/Initializer/BatchRecorders/DataTypes/Table.h:25-61) or base + index * stride.
struct ElemArray {
  const void* const* ptrs;
  const char* base;
  long long strideBytes;
  __device__ __forceinline__ const void* at(long long i) const {
    return ptrs != nullptr ? ptrs[i] : static_cast<const void*>(base + i * strideBytes);
  }
  __host__ __device__ bool present() const { return ptrs != nullptr || base != nullptr; }
};

// ---------------------------------------------------------------------------------------
// mbarrier + 1-D TMA (cp.async.bulk) wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load(void* smemDst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smemDst)),
               "l"(gsrc),
               "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void tma_store(void* gdst, const void* smemSrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smemSrc)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the committed bulk stores have finished READING shared memory
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA store source)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------
// 4 consecutive reals, 16-byte aligned: one 128-bit access (float) or two (double)
__device__ __forceinline__ void ld4(float (&v)[4], const float* p) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(double (&v)[4], const double* p) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void st4(double* p, double a, double b, double c, double d) {
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
  *reinterpret_cast<double2*>(p + 2) = make_double2(c, d);
}

// pad a per-element shared-memory stride (in reals) so that stride*sizeof/16 is odd:
// 128-bit accesses of 8 consecutive lanes (one element each) then hit 8 distinct
// 16-byte bank groups -> conflict-free.
__host__ __device__ constexpr int padOdd16(int reals, int realSize) {
  const int per = 16 / realSize;
  const int units = (reals + per - 1) / per;
  return (units % 2 == 1 ? units : units + 1) * per;
}
__host__ __device__ constexpr int pad4i(int n) { return (n + 3) / 4 * 4; }

// CSC layout of the 24-entry elastic star matrix (codegen/matrices/star.xml; values
// stored column by column).  For column q: rows and offset.
__device__ constexpr int kStarRow[9][3] = {{6, 7, 8}, {6, 7, 8}, {6, 7, 8}, {6, 7, 6}, {7, 8, 7},
                                           {6, 8, 6}, {0, 3, 5}, {1, 3, 4}, {2, 4, 5}};
__device__ constexpr int kStarCnt[9] = {3, 3, 3, 2, 2, 2, 3, 3, 3};
__device__ constexpr int kStarOff[9] = {0, 3, 6, 9, 11, 13, 15, 18, 21};

} // namespace ssb
