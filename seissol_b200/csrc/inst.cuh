// Instantiation body shared by inst_o{4,5,6}_f{32,64}.cu.  The including file defines
//   SSB_ORDER (4|5|6), SSB_REAL (float|double), SSB_TAG (o6_f32), SSB_E_LOCAL, SSB_E_NEIGH
// and includes gen/ops_o<SSB_ORDER>.cuh before this header.
#include "api_internal.h"
#include "kernels_elastic.cuh"
#include "kernels_faceproj.cuh"

#define SSB_CAT2(a, b) a##b
#define SSB_CAT(a, b) SSB_CAT2(a, b)
#define SSB_GEN SSB_CAT(gen_o, SSB_ORDER)

#include <cstdlib>

namespace ssb {
namespace {

// grid size: one CTA per batch, or (SSB_PERSISTENT=k) k CTAs per SM walking the batches in lock step
inline unsigned gridFor(long long batches) {
  static int persistent = -1;
  static int sms = 0;
  if (persistent < 0) {
    const char* e = std::getenv("SSB_PERSISTENT");
    persistent = e != nullptr ? std::atoi(e) : 0;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (persistent > 0 && batches > static_cast<long long>(persistent) * sms) {
    return static_cast<unsigned>(persistent * sms);
  }
  return static_cast<unsigned>(batches);
}

// development aid: SSB_SMEM_PAD=<bytes> requests extra dynamic shared memory per CTA, i.e. lowers the number of
// resident CTAs per SM without recompiling (occupancy experiments)
inline size_t smemPad() {
  static long pad = -1;
  if (pad < 0) {
    const char* e = std::getenv("SSB_SMEM_PAD");
    pad = e != nullptr ? std::atol(e) : 0;
  }
  return static_cast<size_t>(pad);
}

// Traits binding the generated operator code of one order to the kernel templates.
struct Gen {
  static constexpr int kOrder = SSB_GEN::kOrder;
  static constexpr int kB = SSB_GEN::kB;
  static constexpr int kB2 = SSB_GEN::kB2;
  static constexpr int kBP = SSB_GEN::kBP;
  static constexpr int kB2P = SSB_GEN::kB2P;
  __host__ __device__ static constexpr int derRows(int d) {
    return (kOrder - d) * (kOrder - d + 1) * (kOrder - d + 2) / 6;
  }
  template <typename R, typename Ctx>
  __device__ __forceinline__ static void ader_run(Ctx& c) { SSB_GEN::ader_run<R>(c); }
  template <typename R>
  __device__ __forceinline__ static void volume(typename Pk<R>::T (&acc)[kBP], const R* a0, const R* a1, const R* a2,
                                                const typename Pk<R>::T (&s)[3][3]) {
    SSB_GEN::volume<R>(acc, a0, a1, a2, s);
  }
  template <typename R, typename Src>
  __device__ __forceinline__ static void fMrT(int f, typename Pk<R>::T (&out)[kB2P], const Src& x) {
    SSB_GEN::fMrT<R>(f, out, x);
  }
  template <typename R, typename Src>
  __device__ __forceinline__ static void rT(int f, typename Pk<R>::T (&out)[kB2P], const Src& x) {
    SSB_GEN::rT<R>(f, out, x);
  }
  template <typename R, typename JS>
  __device__ __forceinline__ static void rDivM(int f, typename Pk<R>::T (&out)[kBP], const JS& js) {
    SSB_GEN::rDivM<R>(f, out, js);
  }
  template <typename R>
  __device__ __forceinline__ static void fP(int h, typename Pk<R>::T (&out)[kB2P], const typename Pk<R>::T (&x)[kB2P]) {
    SSB_GEN::fP<R>(h, out, x);
  }
  template <typename R>
  __device__ __forceinline__ static void fPD(int h, R (&out)[kB2P], const R (&x)[kB2P]) {
    static_assert(Pk<R>::V == 1, "the face-projection path is scalar");
    SSB_GEN::fPD<R>(h, out, x);
  }
  template <typename R>
  __device__ __forceinline__ static void rDivMr(int f, R (&out)[kBP], const R (&x)[kB2P]) {
    SSB_GEN::rDivMr<R>(f, out, x);
  }
};

using Real = SSB_REAL;
constexpr int EL = SSB_E_LOCAL;  // groups of V elements per CTA
constexpr int EN = SSB_E_NEIGH;
constexpr int VL = Pk<Real>::V;

template <bool ADER, bool LOCAL>
int launchLocal(const LocalParams& p, cudaStream_t stream) {
  using P = LocalPlan<Real, Gen, EL>;
  auto kern = k_local<Real, Gen, EL, ADER, LOCAL>;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(P::bytes + smemPad()));
    if (err != cudaSuccess) {
      return int(err);
    }
    configured = true;
  }
  if (p.n <= 0) {
    return 0;
  }
  const unsigned grid = gridFor((p.n + EL * VL - 1) / (EL * VL));
  kern<<<grid, blockThreads(EL), P::bytes + smemPad(), stream>>>(p);
  return int(cudaGetLastError());
}

int localEntry(const LocalParams& p, bool ader, bool local, cudaStream_t stream) {
  if (ader && local) {
    return launchLocal<true, true>(p, stream);
  }
  if (ader) {
    return launchLocal<true, false>(p, stream);
  }
  return launchLocal<false, true>(p, stream);
}

// elements per CTA of the face-projection neighbor kernel: two resident CTAs per SM (about 100 KB each)
constexpr int neighFpElems() {
  const int per = padOdd16(4 * kNQ * Gen::kB2P + 4 * 81, sizeof(Real)) * int(sizeof(Real));
  const int e = 101000 / per;
  return e > 28 ? 28 : e;
}
#ifdef SSB_E_NEIGHFP
constexpr int ENF = SSB_E_NEIGHFP;
#else
constexpr int ENF = neighFpElems();
#endif

int neighFpEntry(const NeighParams& p, cudaStream_t stream) {
#if SSB_PACK_F32 == 1
  using P = NeighFpPlan<Real, Gen, ENF>;
  auto kern = k_neigh_fp<Real, Gen, ENF>;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(P::bytes + smemPad()));
    if (err != cudaSuccess) {
      return int(err);
    }
    configured = true;
  }
  if (p.n <= 0) {
    return 0;
  }
  const unsigned grid = gridFor((p.n + ENF - 1) / ENF);
  kern<<<grid, P::threads, P::bytes + smemPad(), stream>>>(p);
  return int(cudaGetLastError());
#else
  return -2;
#endif
}

int neighEntry(const NeighParams& p, cudaStream_t stream) {
  if (p.faceProj.present() || p.useFaceProj) {
    return neighFpEntry(p, stream);
  }
  using P = NeighPlan<Real, Gen, EN>;
  auto kern = k_neigh<Real, Gen, EN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(P::bytes + smemPad()));
    if (err != cudaSuccess) {
      return int(err);
    }
    configured = true;
  }
  if (p.n <= 0) {
    return 0;
  }
  const unsigned grid = gridFor((p.n + EN * VL - 1) / (EN * VL));
  kern<<<grid, blockThreads(EN), P::bytes + smemPad(), stream>>>(p);
  return int(cudaGetLastError());
}

int taylorEntry(const TaylorParams& p, cudaStream_t stream) {
  if (p.n <= 0) {
    return 0;
  }
  const size_t bytes = (size_t(kTaylorE) * (p.derSize + kNQ * p.ldg) * sizeof(Real) + 15) / 16 * 16 + 16;
  static size_t configured = 0;
  if (bytes > configured) {
    cudaError_t err = cudaFuncSetAttribute(k_taylor<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    if (err != cudaSuccess) {
      return int(err);
    }
    configured = bytes;
  }
  k_taylor<Real><<<unsigned((p.n + kTaylorE - 1) / kTaylorE), 256, bytes, stream>>>(p);
  return int(cudaGetLastError());
}

int accumulateEntry(const AccumulateParams& p, cudaStream_t stream) {
  if (p.n <= 0) {
    return 0;
  }
  long long blocks = (p.n * p.vec4PerElem + 255) / 256;
  if (blocks > 148LL * 16) {
    blocks = 148LL * 16;
  }
  k_lts_accumulate<Real><<<unsigned(blocks), 256, 0, stream>>>(p);
  return int(cudaGetLastError());
}

const KernelSet kSet = {Gen::kOrder, int(sizeof(Real)), Gen::kB, Gen::kBP, EL * VL, EN * VL, localEntry, neighEntry, taylorEntry, accumulateEntry};

}  // namespace

const KernelSet* SSB_CAT(kernelset_, SSB_TAG)() { return &kSet; }

}  // namespace ssb
