// Minimal hand-written tcgen05 (5th-generation tensor core) layer for sm_100a: TMEM allocation, shared-memory matrix
// descriptors for the canonical no-swizzle K-major layout, the kind::tf32 MMA, commit -> mbarrier, TMEM -> registers.
// Bit layouts follow the PTX ISA "tcgen05" chapter as mirrored in CUTLASS (cute/arch/mma_sm100_desc.hpp: SmemDescriptor,
// InstrDescriptor; cute/atom/mma_traits_sm100.hpp: canonical layouts) -- no CUTLASS code is used.
//
// Canonical K-major, no-swizzle operand (R rows = M or N, K columns, 32-bit elements, T = 4 elements per 16 bytes):
//   element (r, k), r = r0 + 8 r1, k = k0 + 4 k1  lives at  r0*16 B + k0*4 B + r1*SBO + k1*LBO
// i.e. 8x4 "core matrices" of 128 contiguous bytes; LBO = byte stride between core matrices along K, SBO = byte
// stride between 8-row groups.  One kind::tf32 instruction consumes K = 8 (two core matrices along K).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ssb {
namespace umma {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// 64-bit shared-memory matrix descriptor (SWIZZLE_NONE, version 1 = Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lboBytes, uint32_t sboBytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3fffu);              // start address, bits [0,14)
  d |= static_cast<uint64_t>((lboBytes >> 4) & 0x3fffu) << 16;    // leading-dimension byte offset, bits [16,30)
  d |= static_cast<uint64_t>((sboBytes >> 4) & 0x3fffu) << 32;    // stride-dimension byte offset, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                            // descriptor version (Blackwell)
  return d;                                                        // base offset 0, lbo mode 0, layout type 0 (no swizzle)
}

// 32-bit instruction descriptor: D (f32) += A (tf32, K-major) . B (tf32, K-major), M x N tile
__host__ __device__ constexpr uint32_t instr_desc_tf32(int m, int n) {
  return (1u << 4)                                 // c_format = F32
         | (2u << 7) | (2u << 10)                  // a_format = b_format = TF32
         | (0u << 15) | (0u << 16)                 // a_major = b_major = K
         | (static_cast<uint32_t>(n >> 3) << 17)   // n_dim
         | (static_cast<uint32_t>(m >> 4) << 24);  // m_dim
}

// one warp allocates `cols` TMEM columns (power of two >= 32); the base address lands in shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t* smemDst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(smemDst)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  const uint32_t zero = 0u;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmemD),
      "l"(descA), "l"(descB), "r"(idesc), "r"(acc), "r"(zero)
      : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier when they are done (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// TMEM -> registers: the calling warp w (w % 4 selects lanes 32 (w % 4) ..) reads 8 consecutive columns of its 32 lanes
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// split of an fp32 value into two tf32 values: hi = the 10 explicit mantissa bits that the tensor core reads,
// lo = the remainder (exact in fp32), itself cut to tf32.  a b ~ ah bh + ah bl + al bh with relative error ~ 2^-21.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ float tf32_lo(float x, float hi) { return __uint_as_float(__float_as_uint(x - hi) & 0xffffe000u); }

}  // namespace umma
}  // namespace ssb
