// Instantiation body shared by inst_o{4,5,6}_f{32,64}.cu.  The including file defines
//   SSB_ORDER (4|5|6), SSB_REAL (float|double), SSB_TAG (o6_f32), SSB_E_LOCAL, SSB_E_NEIGH
// and includes gen/ops_o<SSB_ORDER>.cuh before this header.
#include "api_internal.h"

#define SSB_CAT2(a, b) a##b
#define SSB_CAT(a, b) SSB_CAT2(a, b)
#define SSB_GEN SSB_CAT(gen_o, SSB_ORDER)

namespace ssb {
namespace {

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
  __device__ __forceinline__ static void volume(R (&acc)[kBP], const R* a0, const R* a1, const R* a2,
                                                const R (&s)[3][3]) {
    SSB_GEN::volume<R>(acc, a0, a1, a2, s);
  }
  template <typename R>
  __device__ __forceinline__ static void fMrT(int f, R (&out)[kB2P], const R* x) { SSB_GEN::fMrT<R>(f, out, x); }
  template <typename R>
  __device__ __forceinline__ static void rT(int f, R (&out)[kB2P], const R* x) { SSB_GEN::rT<R>(f, out, x); }
  template <typename R>
  __device__ __forceinline__ static void rDivM(int f, R (&out)[kBP], const R (&x)[kB2P]) {
    SSB_GEN::rDivM<R>(f, out, x);
  }
  template <typename R>
  __device__ __forceinline__ static void fP(int h, R (&out)[kB2P], const R (&x)[kB2P]) {
    SSB_GEN::fP<R>(h, out, x);
  }
};

using Real = SSB_REAL;
constexpr int EL = SSB_E_LOCAL;
constexpr int EN = SSB_E_NEIGH;

template <bool ADER, bool LOCAL>
int launchLocal(const LocalParams& p, cudaStream_t stream) {
  using P = LocalPlan<Real, Gen, EL>;
  auto kern = k_local<Real, Gen, EL, ADER, LOCAL>;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(P::bytes));
    if (err != cudaSuccess) {
      return int(err);
    }
    configured = true;
  }
  if (p.n <= 0) {
    return 0;
  }
  const unsigned grid = unsigned((p.n + EL - 1) / EL);
  kern<<<grid, blockThreads(EL), P::bytes, stream>>>(p);
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
  using P = NeighPlan<Real, Gen, EN>;
  auto kern = k_neigh<Real, Gen, EN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(P::bytes));
    if (err != cudaSuccess) {
      return int(err);
    }
    configured = true;
  }
  if (p.n <= 0) {
    return 0;
  }
  const unsigned grid = unsigned((p.n + EN - 1) / EN);
  kern<<<grid, blockThreads(EN), P::bytes, stream>>>(p);
  return int(cudaGetLastError());
}

int taylorEntry(const TaylorParams& p, cudaStream_t stream) {
  if (p.n <= 0) {
    return 0;
  }
  const long long total = p.n * kNQ * ((p.rows[0] + 3) / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) {
    blocks = 148LL * 32;
  }
  k_taylor<Real><<<unsigned(blocks), 256, 0, stream>>>(p);
  return int(cudaGetLastError());
}

const KernelSet kSet = {Gen::kOrder, int(sizeof(Real)), Gen::kB, Gen::kBP, EL, EN, localEntry, neighEntry, taylorEntry};

}  // namespace

const KernelSet* SSB_CAT(kernelset_, SSB_TAG)() { return &kSet; }

}  // namespace ssb
