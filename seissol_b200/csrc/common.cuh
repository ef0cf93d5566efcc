// Shared device helpers: TMA bulk copies + mbarriers (sm_100a inline PTX), 128-bit
// shared-memory vector access, and the per-element array descriptor used at the
// C-ABI boundary.  Hand-written for sm_100a; there is no fallback path.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// Lock-step markers of the generated operator code (codegen/gen_ops.py emits SSB_LS(i) every 64 FMAs in the operators
// that every thread of a CTA executes uniformly).  With SSB_LOCKSTEP = P > 0 every (P/64)-th marker is a CTA barrier:
// the warps of a CTA cannot drift apart in a straight-line instruction stream that is larger than the instruction
// caches, so they share every fetched line (tools/ubench/gen_ifetch4.py).
#ifndef SSB_LOCKSTEP
#define SSB_LOCKSTEP 0
#endif
#if SSB_LOCKSTEP > 0
#define SSB_LS(i) do { if ((i) % (SSB_LOCKSTEP / 64) == 0) __syncthreads(); } while (0)
#else
#define SSB_LS(i) do { } while (0)
#endif

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


// ---------------------------------------------------------------------------------------
// Packed lane type: a thread processes V elements at once.  float -> float2 driven by the
// Blackwell packed-FP32 instructions (fma.rn.f32x2 -> SASS FFMA2; with a broadcast
// compile-time constant ptxas emits the immediate form `FFMA2 R, R, imm, R`, i.e. two FMAs
// per 16-byte instruction: half the instruction-fetch bytes per flop).
// double -> plain double (V = 1).
template <typename R> struct Pk;
#ifndef SSB_PACK_F32
#define SSB_PACK_F32 1  // 2: two elements per thread on packed FFMA2; 1: one element per thread, scalar FFMA
#endif
#if SSB_PACK_F32 == 2
template <> struct Pk<float> { using T = float2; static constexpr int V = 2; };
#else
template <> struct Pk<float> { using T = float; static constexpr int V = 1; };
#endif
#ifndef SSB_PACK_F64
#define SSB_PACK_F64 1  // 2: two elements per thread in double precision (one materialised constant feeds two DFMAs)
#endif
#if SSB_PACK_F64 == 2
template <> struct Pk<double> { using T = double2; static constexpr int V = 2; };
#else
template <> struct Pk<double> { using T = double; static constexpr int V = 1; };
#endif

__device__ __forceinline__ unsigned long long pk_bits(float2 a) { return *reinterpret_cast<unsigned long long*>(&a); }
__device__ __forceinline__ float2 pk_unbits(unsigned long long a) { return *reinterpret_cast<float2*>(&a); }
__device__ __forceinline__ float2 pk_fma(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk_bits(a)), "l"(pk_bits(b)), "l"(pk_bits(c)));
  return pk_unbits(d);
}
__device__ __forceinline__ float2 pk_mul(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_bits(a)), "l"(pk_bits(b)));
  return pk_unbits(d);
}
__device__ __forceinline__ float2 pk_add(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk_bits(a)), "l"(pk_bits(b)));
  return pk_unbits(d);
}
__device__ __forceinline__ float2 pk_fmac(float c, float2 x, float2 acc) { return pk_fma(make_float2(c, c), x, acc); }
__device__ __forceinline__ float2 pk_mulc(float c, float2 x) { return pk_mul(make_float2(c, c), x); }
__device__ __forceinline__ float pk_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float pk_mul(float a, float b) { return a * b; }
__device__ __forceinline__ float pk_add(float a, float b) { return a + b; }
__device__ __forceinline__ float pk_fmac(float c, float x, float acc) { return fmaf(c, x, acc); }
__device__ __forceinline__ float pk_mulc(float c, float x) { return c * x; }
__device__ __forceinline__ float pk_get(float a, int) { return a; }
__device__ __forceinline__ float pk_make(const float (&a)[1]) { return a[0]; }
__device__ __forceinline__ float pk_sel01(float a, float) { return a; }
__device__ __forceinline__ void ldv4(float (&v)[4], const float* p) { ld4(v, p); }
__device__ __forceinline__ void stv4(float* p, float a, float b, float c, float d) { st4(p, a, b, c, d); }
__device__ __forceinline__ void st4lane(float* p, int, float a, float b, float c, float d) { st4(p, a, b, c, d); }
__device__ __forceinline__ double pk_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ double pk_mul(double a, double b) { return a * b; }
__device__ __forceinline__ double pk_add(double a, double b) { return a + b; }
__device__ __forceinline__ double pk_fmac(double c, double x, double acc) { return fma(c, x, acc); }
__device__ __forceinline__ double pk_mulc(double c, double x) { return c * x; }
// packed double: two elements per thread.  DFMA has no immediate operand: a literal costs two 32-bit moves, which two
// elements now share (tools/ubench/gen_cbank2.py: 0.11 -> 0.34 DFMA/clk/SMSP at 8 warps per SM)
__device__ __forceinline__ double2 pk_fma(double2 a, double2 b, double2 c) {
  return make_double2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y));
}
__device__ __forceinline__ double2 pk_mul(double2 a, double2 b) { return make_double2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ double2 pk_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 pk_fmac(double c, double2 x, double2 acc) {
  return make_double2(fma(c, x.x, acc.x), fma(c, x.y, acc.y));
}
__device__ __forceinline__ double2 pk_mulc(double c, double2 x) { return make_double2(c * x.x, c * x.y); }
template <typename R> __device__ __forceinline__ typename Pk<R>::T pk_zero();
#if SSB_PACK_F32 == 2
template <> __device__ __forceinline__ float2 pk_zero<float>() { return make_float2(0.f, 0.f); }
#else
template <> __device__ __forceinline__ float pk_zero<float>() { return 0.f; }
#endif
#if SSB_PACK_F64 == 2
template <> __device__ __forceinline__ double2 pk_zero<double>() { return make_double2(0.0, 0.0); }
#else
template <> __device__ __forceinline__ double pk_zero<double>() { return 0.0; }
#endif
// lane v of a packed value / packed value from per-lane scalars
__device__ __forceinline__ float pk_get(float2 a, int v) { return v == 0 ? a.x : a.y; }
__device__ __forceinline__ double pk_get(double a, int) { return a; }
__device__ __forceinline__ float2 pk_make(const float (&a)[2]) { return make_float2(a[0], a[1]); }
__device__ __forceinline__ double pk_make(const double (&a)[1]) { return a[0]; }
__device__ __forceinline__ double pk_get(double2 a, int v) { return v == 0 ? a.x : a.y; }
__device__ __forceinline__ double2 pk_make(const double (&a)[2]) { return make_double2(a[0], a[1]); }
__device__ __forceinline__ double2 pk_sel01(double2 a, double2 b) { return make_double2(a.x, b.y); }
// keep lane 0 of a, lane 1 of b
__device__ __forceinline__ float2 pk_sel01(float2 a, float2 b) { return make_float2(a.x, b.y); }
__device__ __forceinline__ double pk_sel01(double a, double) { return a; }

// keep lane v of a packed value only if keep[v], else exactly zero (a select, not a multiply: the dropped lane may
// hold NaN/Inf computed from stale shared memory)
__device__ __forceinline__ float2 pk_keep(float2 a, const bool (&keep)[2]) {
  return make_float2(keep[0] ? a.x : 0.f, keep[1] ? a.y : 0.f);
}
__device__ __forceinline__ float pk_keep(float a, const bool (&keep)[1]) { return keep[0] ? a : 0.f; }
__device__ __forceinline__ double pk_keep(double a, const bool (&keep)[1]) { return keep[0] ? a : 0.0; }
__device__ __forceinline__ double2 pk_keep(double2 a, const bool (&keep)[2]) {
  return make_double2(keep[0] ? a.x : 0.0, keep[1] ? a.y : 0.0);
}

// 4 consecutive packed entries from a lane-interleaved column ([k][V] reals)
__device__ __forceinline__ void ldv4(float2 (&v)[4], const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
  v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
}
__device__ __forceinline__ void ldv4(double (&v)[4], const double* p) { ld4(v, p); }
__device__ __forceinline__ void ldv4(double2 (&v)[4], const double* p) {
  v[0] = *reinterpret_cast<const double2*>(p); v[1] = *reinterpret_cast<const double2*>(p + 2);
  v[2] = *reinterpret_cast<const double2*>(p + 4); v[3] = *reinterpret_cast<const double2*>(p + 6);
}
__device__ __forceinline__ void stv4(float* p, float2 a, float2 b, float2 c, float2 d) {
  *reinterpret_cast<float4*>(p) = make_float4(a.x, a.y, b.x, b.y);
  *reinterpret_cast<float4*>(p + 4) = make_float4(c.x, c.y, d.x, d.y);
}
__device__ __forceinline__ void stv4(double* p, double a, double b, double c, double d) { st4(p, a, b, c, d); }
__device__ __forceinline__ void stv4(double* p, double2 a, double2 b, double2 c, double2 d) {
  *reinterpret_cast<double2*>(p) = a; *reinterpret_cast<double2*>(p + 2) = b;
  *reinterpret_cast<double2*>(p + 4) = c; *reinterpret_cast<double2*>(p + 6) = d;
}

// 2 consecutive packed entries (one 128-bit access for float2 / double, 64-bit for scalar float)
__device__ __forceinline__ void ldv2(float2 (&v)[2], const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
}
__device__ __forceinline__ void ldv2(float (&v)[2], const float* p) {
  const float2 a = *reinterpret_cast<const float2*>(p);
  v[0] = a.x; v[1] = a.y;
}
__device__ __forceinline__ void ldv2(double (&v)[2], const double* p) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  v[0] = a.x; v[1] = a.y;
}
__device__ __forceinline__ void ldv2(double2 (&v)[2], const double* p) {
  v[0] = *reinterpret_cast<const double2*>(p); v[1] = *reinterpret_cast<const double2*>(p + 2);
}
__device__ __forceinline__ void stv2(double* p, double2 a, double2 b) {
  *reinterpret_cast<double2*>(p) = a; *reinterpret_cast<double2*>(p + 2) = b;
}
__device__ __forceinline__ void stv2(float* p, float2 a, float2 b) {
  *reinterpret_cast<float4*>(p) = make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void stv2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void stv2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }
// lane v of 2 packed values -> 2 consecutive reals
__device__ __forceinline__ void st2lane(float* p, int v, float2 a, float2 b) {
  *reinterpret_cast<float2*>(p) = v == 0 ? make_float2(a.x, b.x) : make_float2(a.y, b.y);
}
__device__ __forceinline__ void st2lane(float* p, int, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void st2lane(double* p, int, double a, double b) {
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
}
__device__ __forceinline__ void st2lane(double* p, int v, double2 a, double2 b) {
  *reinterpret_cast<double2*>(p) = v == 0 ? make_double2(a.x, b.x) : make_double2(a.y, b.y);
}

// Matrix values of the generated operator code (generation 4): literals, or -- double precision with SSB_KTAB = 1 -- pairs
// streamed from a per-function constant table in global memory (L1-resident, every lane reads the same 16 bytes):
// one LDG.128 per two DFMAs instead of two 32-bit moves per DFMA (codegen/gen_ops4.py::emit_acc).
#ifndef SSB_KTAB
#define SSB_KTAB 0
#endif
template <typename R> struct KPair {};
#if SSB_KTAB == 1
template <> struct KPair<double> { double x, y; };
#endif
template <typename R> __device__ __forceinline__ KPair<R> kload(const double* tab, int j) {
  KPair<R> k;
#if SSB_KTAB == 1
  if constexpr (sizeof(R) == 8) {
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(k.x), "=d"(k.y) : "l"(tab + 2 * j));
  }
#endif
  return k;
}
template <typename R, int I> __device__ __forceinline__ R ksel(const KPair<R>& k, R literal) {
#if SSB_KTAB == 1
  if constexpr (sizeof(R) == 8) {
    return I == 0 ? k.x : k.y;
  }
#endif
  return literal;
}

// Operand sources of the generated sparse mat-vec code: a lane-interleaved column, or V
// separate per-element columns (TMA landing zones) packed on the fly.
template <typename R> struct SrcPk {
  const R* p;
  __device__ __forceinline__ void load4(typename Pk<R>::T (&v)[4], int k0) const { ldv4(v, p + k0 * Pk<R>::V); }
  __device__ __forceinline__ void load2(typename Pk<R>::T (&v)[2], int k0) const { ldv2(v, p + k0 * Pk<R>::V); }
};
template <typename R> struct SrcSplit;
template <> struct SrcSplit<float> {
  const float* p0; const float* p1;
  __device__ __forceinline__ void load4(float2 (&v)[4], int k0) const {
    const float4 a = *reinterpret_cast<const float4*>(p0 + k0);
    const float4 b = *reinterpret_cast<const float4*>(p1 + k0);
    v[0] = make_float2(a.x, b.x); v[1] = make_float2(a.y, b.y);
    v[2] = make_float2(a.z, b.z); v[3] = make_float2(a.w, b.w);
  }
  __device__ __forceinline__ void load4(float (&v)[4], int k0) const { ld4(v, p0 + k0); }
  __device__ __forceinline__ void load2(float2 (&v)[2], int k0) const {
    const float2 a = *reinterpret_cast<const float2*>(p0 + k0);
    const float2 b = *reinterpret_cast<const float2*>(p1 + k0);
    v[0] = make_float2(a.x, b.x); v[1] = make_float2(a.y, b.y);
  }
  __device__ __forceinline__ void load2(float (&v)[2], int k0) const { ldv2(v, p0 + k0); }
};
template <> struct SrcSplit<double> {
  const double* p0; const double* p1;
  __device__ __forceinline__ void load4(double (&v)[4], int k0) const { ld4(v, p0 + k0); }
  __device__ __forceinline__ void load2(double (&v)[2], int k0) const { ldv2(v, p0 + k0); }
  __device__ __forceinline__ void load4(double2 (&v)[4], int k0) const {
    double a[4], b[4];
    ld4(a, p0 + k0); ld4(b, p1 + k0);
    v[0] = make_double2(a[0], b[0]); v[1] = make_double2(a[1], b[1]);
    v[2] = make_double2(a[2], b[2]); v[3] = make_double2(a[3], b[3]);
  }
  __device__ __forceinline__ void load2(double2 (&v)[2], int k0) const {
    const double2 a = *reinterpret_cast<const double2*>(p0 + k0);
    const double2 b = *reinterpret_cast<const double2*>(p1 + k0);
    v[0] = make_double2(a.x, b.x); v[1] = make_double2(a.y, b.y);
  }
};
// lane v of 4 packed values -> one 128-bit (float) / 256-bit (double) store
__device__ __forceinline__ void st4lane(float* p, int v, float2 a, float2 b, float2 c, float2 d) {
  st4(p, pk_get(a, v), pk_get(b, v), pk_get(c, v), pk_get(d, v));
}
__device__ __forceinline__ void st4lane(double* p, int, double a, double b, double c, double d) { st4(p, a, b, c, d); }
__device__ __forceinline__ void st4lane(double* p, int v, double2 a, double2 b, double2 c, double2 d) {
  st4(p, pk_get(a, v), pk_get(b, v), pk_get(c, v), pk_get(d, v));
}


// scalar type of a packed lane type, zero of a packed lane type, scalar broadcast
template <typename PT> struct PkScalar;
template <> struct PkScalar<float2> { using T = float; static constexpr int V = 2; };
template <> struct PkScalar<float> { using T = float; static constexpr int V = 1; };
template <> struct PkScalar<double> { using T = double; static constexpr int V = 1; };
template <> struct PkScalar<double2> { using T = double; static constexpr int V = 2; };
template <typename PT> __device__ __forceinline__ PT pk_zero_of();
template <> __device__ __forceinline__ float2 pk_zero_of<float2>() { return make_float2(0.f, 0.f); }
template <> __device__ __forceinline__ float pk_zero_of<float>() { return 0.f; }
template <> __device__ __forceinline__ double pk_zero_of<double>() { return 0.0; }
template <> __device__ __forceinline__ double2 pk_zero_of<double2>() { return make_double2(0.0, 0.0); }
template <typename PT> __device__ __forceinline__ PT pk_bcast(typename PkScalar<PT>::T x);
template <> __device__ __forceinline__ float2 pk_bcast<float2>(float x) { return make_float2(x, x); }
template <> __device__ __forceinline__ float pk_bcast<float>(float x) { return x; }
template <> __device__ __forceinline__ double pk_bcast<double>(double x) { return x; }
template <> __device__ __forceinline__ double2 pk_bcast<double2>(double x) { return make_double2(x, x); }
// Dynamic-rupture kernels always pack two TIME POINTS per thread in single precision.
template <typename R> struct PkDr;
template <> struct PkDr<float> { using T = float2; static constexpr int V = 2; };
template <> struct PkDr<double> { using T = double; static constexpr int V = 1; };
// operand source: a register-resident column
template <typename PT, int N> struct SrcReg {
  const PT (&x)[N];
  __device__ __forceinline__ void load4(PT (&v)[4], int k0) const {
    v[0] = x[k0]; v[1] = x[k0 + 1]; v[2] = x[k0 + 2]; v[3] = x[k0 + 3];
  }
};

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
