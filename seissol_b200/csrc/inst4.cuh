// Instantiation body of kernel generation 4 (half-column threads, kernels_elastic4.cuh), shared by
// inst4_o{4,5,6}_f{32,64}.cu.  The including file defines SSB_ORDER, SSB_REAL, SSB_TAG, SSB_NP (groups per
// CTA) and includes gen/ops4_o<SSB_ORDER>.cuh before this header.  Single precision is compiled with
// SSB_PACK_F32 = 2 (two elements per thread on packed FFMA2).
#include "api_internal.h"
#include "kernels_elastic4.cuh"

#define SSB_CAT2(a, b) a##b
#define SSB_CAT(a, b) SSB_CAT2(a, b)
#define SSB_GEN SSB_CAT(gen4_o, SSB_ORDER)

#include <cstdlib>

namespace ssb {
namespace {

struct Gen {
  static constexpr int kOrder = SSB_GEN::kOrder;
  static constexpr int kB = SSB_GEN::kB;
  static constexpr int kB2 = SSB_GEN::kB2;
  static constexpr int kBP = SSB_GEN::kBP;
  static constexpr int kB2P = SSB_GEN::kB2P;
  __host__ __device__ static constexpr int derRows(int d) {
    return (kOrder - d) * (kOrder - d + 1) * (kOrder - d + 2) / 6;
  }
  template <typename R, int H, typename Ctx>
  __device__ __forceinline__ static void ader_run(Ctx& c) { SSB_GEN::ader_run<R, H>(c); }
  template <typename R, int H>
  __device__ __forceinline__ static void volume(typename Pk<R>::T (&acc)[kBP / 2], const R* a0, const R* a1,
                                                const R* a2, const typename Pk<R>::T (&s)[3][3]) {
    SSB_GEN::volume<R, H>(acc, a0, a1, a2, s);
  }
  template <typename R, int H, typename Src>
  __device__ __forceinline__ static void fMrT(int f, typename Pk<R>::T (&out)[kB2P / 2], const Src& x) {
    SSB_GEN::fMrT<R, H>(f, out, x);
  }
  template <typename R, int H, typename Src>
  __device__ __forceinline__ static void rT(int f, typename Pk<R>::T (&out)[kB2P / 2], const Src& x) {
    SSB_GEN::rT<R, H>(f, out, x);
  }
  template <typename R, int H, typename Src>
  __device__ __forceinline__ static void fP(int h, typename Pk<R>::T (&out)[kB2P / 2], const Src& x) {
    SSB_GEN::fP<R, H>(h, out, x);
  }
  template <typename R, int H, typename JS>
  __device__ __forceinline__ static void rDivM(int f, typename Pk<R>::T (&out)[kBP / 2], const JS& js) {
    SSB_GEN::rDivM<R, H>(f, out, js);
  }
};

using Real = SSB_REAL;
constexpr int NP = SSB_NP;  // groups of V elements per CTA
constexpr int VL = Pk<Real>::V;

inline unsigned gridFor4(long long batches) {
  static int persistent = -1;
  if (persistent < 0) {
    const char* e = std::getenv("SSB_PERSISTENT");
    persistent = e != nullptr ? std::atoi(e) : 0;
  }
  const int sms = persistent > 0 ? smCountOfCurrentDevice() : 0;
  if (persistent > 0 && batches > static_cast<long long>(persistent) * sms) {
    return static_cast<unsigned>(persistent * sms);
  }
  return static_cast<unsigned>(batches);
}

template <bool ADER, bool LOCAL>
int launchLocal(const LocalParams& p, cudaStream_t stream) {
  using P = LocalPlan<Real, Gen, NP>;
  auto kern = k_local4<Real, Gen, NP, ADER, LOCAL>;
  static SmemConfig smemConfig;
  if (int rc = ensureDynamicSmem(kern, P::bytes, smemConfig)) {
    return rc;
  }
  if (p.n <= 0) {
    return 0;
  }
  const unsigned grid = gridFor4((p.n + NP * VL - 1) / (NP * VL));
  kern<<<grid, 2 * halfThreads(NP), P::bytes, stream>>>(p);
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

int neighEntry(const NeighParams& p, cudaStream_t stream) {
  using P = NeighPlan<Real, Gen, NP>;
  auto kern = k_neigh4<Real, Gen, NP>;
  static SmemConfig smemConfig;
  if (int rc = ensureDynamicSmem(kern, P::bytes, smemConfig)) {
    return rc;
  }
  if (p.n <= 0) {
    return 0;
  }
  const unsigned grid = gridFor4((p.n + NP * VL - 1) / (NP * VL));
  kern<<<grid, 2 * halfThreads(NP), P::bytes, stream>>>(p);
  return int(cudaGetLastError());
}

int taylorEntry(const TaylorParams& p, cudaStream_t stream) {
  if (p.n <= 0) {
    return 0;
  }
  const size_t bytes = (size_t(kTaylorE) * (p.derSize + kNQ * p.ldg) * sizeof(Real) + 15) / 16 * 16 + 16;
  // k_taylor<Real> is shared by every (order, generation) unit of this precision: one configuration object
  if (int rc = ensureDynamicSmem(k_taylor<Real>, bytes, sharedSmemConfig<Real>())) {
    return rc;
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

const KernelSet kSet = {Gen::kOrder, int(sizeof(Real)), Gen::kB, Gen::kBP, NP * VL, NP * VL, localEntry, neighEntry, taylorEntry, accumulateEntry};

}  // namespace

const KernelSet* SSB_CAT(kernelset4_, SSB_TAG)() { return &kSet; }

}  // namespace ssb
