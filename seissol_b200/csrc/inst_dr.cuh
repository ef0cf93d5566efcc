// Instantiation body of the dynamic-rupture kernels, shared by inst_dr_o{4,5,6}_f{32,64}.cu.
// The including file defines SSB_ORDER, SSB_REAL, SSB_TAG and includes gen/dr_o<SSB_ORDER>.cuh.
#include "api_internal.h"
#include "kernels_dr.cuh"

#define SSB_CAT2(a, b) a##b
#define SSB_CAT(a, b) SSB_CAT2(a, b)
#define SSB_GEN_DR SSB_CAT(gen_dr_o, SSB_ORDER)

namespace ssb {
namespace {

struct GenO {
  static constexpr int kOrder = SSB_ORDER;
  static constexpr int kB = SSB_GEN_DR::kB;
  static constexpr int kBP = SSB_GEN_DR::kBP;
  __host__ __device__ static constexpr int derRows(int d) {
    return (kOrder - d) * (kOrder - d + 1) * (kOrder - d + 2) / 6;
  }
};
struct GenDr {
  static constexpr int kP = SSB_GEN_DR::kP;
  static constexpr int kPP = SSB_GEN_DR::kPP;
  template <typename PT, typename Src>
  __device__ __forceinline__ static void V3mTo2n(int i, PT (&out)[kPP], const Src& x) {
    SSB_GEN_DR::V3mTo2n<PT>(i, out, x);
  }
  template <typename R>
  __device__ __forceinline__ static const R* nodalFluxTable() { return SSB_GEN_DR::nodalFluxTable<R>(); }
};

using Real = SSB_REAL;
constexpr int NF = sizeof(SSB_REAL) == 8 ? 4 : 8;  // faces per CTA (k_dr_interp): <= 256 threads, 255 registers
constexpr int NDW = sizeof(SSB_REAL) == 8 ? 4 : 8;  // warps = elements per CTA (k_neigh_dr); static shared memory <= 48 KB

template <typename K>
int setSmem(K kern, size_t bytes, SmemConfig& configured) {
  return ensureDynamicSmem(kern, bytes, configured);
}

int drInterpEntry(DrInterpParams& p, cudaStream_t stream) {
  if (p.n <= 0) {
    return 0;
  }
  auto kern = k_dr_interp<Real, GenO, GenDr, NF>;
  p.rawStride = padOdd16(p.derSize, sizeof(Real));
  const size_t bytes = ((size_t(NF) * (2 * p.rawStride + 84) * sizeof(Real) + 15) / 16) * 16 + 16;
  static SmemConfig configured;
  if (int rc = setSmem(kern, bytes, configured)) {
    return rc;
  }
  const unsigned grid = unsigned((p.n + NF - 1) / NF);
  kern<<<grid, DrInterpCfg<Real, GenO, GenDr, NF>::threads, bytes, stream>>>(p);
  return int(cudaGetLastError());
}

int frictionEntry(const FrictionParams& p, int law, cudaStream_t stream) {
  if (p.n <= 0) {
    return 0;
  }
  constexpr int P = GenDr::kP;
  constexpr int threads = (P + 31) / 32 * 32;
  if (law == 0) {
    k_friction<Real, SSB_ORDER, P, Law::LinearSlipWeakening><<<unsigned(p.n), threads, 0, stream>>>(p);
  } else {
    k_friction<Real, SSB_ORDER, P, Law::RateAndStateFVW><<<unsigned(p.n), threads, 0, stream>>>(p);
  }
  return int(cudaGetLastError());
}

int neighDrEntry(const NeighDrParams& p, cudaStream_t stream) {
  if (p.n <= 0) {
    return 0;
  }
  const unsigned grid = unsigned((p.n + NDW - 1) / NDW);
  k_neigh_dr<Real, GenO, GenDr, NDW><<<grid, 32 * NDW, 0, stream>>>(p);
  return int(cudaGetLastError());
}

const DrKernelSet kDrSet = {SSB_ORDER, int(sizeof(Real)), GenDr::kP, GenDr::kPP, drInterpEntry, frictionEntry, neighDrEntry};

}  // namespace

const DrKernelSet* SSB_CAT(drkernelset_, SSB_TAG)() { return &kDrSet; }

}  // namespace ssb
