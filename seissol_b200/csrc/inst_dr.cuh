// Instantiation body of the dynamic-rupture kernels, shared by inst_dr_o{4,5,6}_f{32,64}.cu.
// The including file defines SSB_ORDER, SSB_REAL, SSB_TAG and includes gen/dr_o<SSB_ORDER>.cuh.
#include "api_internal.h"
#include "kernels_dr.cuh"
#include "kernels_dr_tc.cuh"
#include <cstdlib>
#include <type_traits>

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
  // tensor-core path: V3mTo2n as canonical tf32 hi / lo operands (gen_dr.py)
  static constexpr int kTcK = SSB_GEN_DR::kTcK;
  static constexpr int kTcN = SSB_GEN_DR::kTcN;
  static constexpr int kTcOperand = SSB_GEN_DR::kTcOperand;
  __device__ __forceinline__ static const float* v3CanonHi(int v) { return SSB_GEN_DR::v3CanonHi(v); }
  __device__ __forceinline__ static const float* v3CanonLo(int v) { return SSB_GEN_DR::v3CanonLo(v); }
};

using Real = SSB_REAL;
constexpr int NF = sizeof(SSB_REAL) == 8 ? 4 : 8;  // faces per CTA (k_dr_interp): <= 256 threads, 255 registers
constexpr int NDW = sizeof(SSB_REAL) == 8 ? 4 : 8;  // warps = elements per CTA (k_neigh_dr); static shared memory <= 48 KB

template <typename K>
int setSmem(K kern, size_t bytes, SmemConfig& configured) {
  return ensureDynamicSmem(kern, bytes, configured);
}

// single precision, order >= 5: the space-time interpolation runs on the tensor cores (k_dr_interp_tc: tcgen05 kind::tf32,
// 3-term hi/lo split, A operand in tensor memory): 2.8 vs 5.0 ms at 200 k faces of order 6 (profiles/r2_dr_tc.txt); at
// order 4 the CUDA-core kernel is as fast (1.11 vs 1.17 ms) and stays.  SSB_DR_TC=0 / 1 forces one or the other.
inline bool useTensorCores() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("SSB_DR_TC");
    v = e == nullptr ? (SSB_ORDER >= 5 ? 1 : 0) : (e[0] == '1' ? 1 : 0);
  }
  return v == 1 && std::is_same<Real, float>::value;
}

int drInterpTcEntry(DrInterpParams& p, cudaStream_t stream) {
  using C = DrTcCfg<GenO, GenDr>;
  auto kern = k_dr_interp_tc<GenO, GenDr>;
  p.rawStride = padOdd16(p.derSize, sizeof(float));
  const size_t bytes = C::bytes(p.rawStride);
  if (bytes > 227 * 1024) {
    return -1;   // (padded reference layout of the derivative families at order 6): the caller takes the CUDA-core kernel
  }
  static SmemConfig configured;
  if (int rc = setSmem(kern, bytes, configured)) {
    return rc;
  }
  const long long tiles = (p.n + C::FPT - 1) / C::FPT;
  const long long ctas = (tiles + C::NW - 1) / C::NW;
  const long long sms = smCountOfCurrentDevice();
  const unsigned grid = unsigned(ctas < sms ? ctas : sms);
  kern<<<grid, C::THREADS, bytes, stream>>>(p);
  return int(cudaGetLastError());
}

int drInterpEntry(DrInterpParams& p, cudaStream_t stream) {
  if (p.n <= 0) {
    return 0;
  }
  if (useTensorCores()) {
    const int rc = drInterpTcEntry(p, stream);
    if (rc != -1) {
      return rc;
    }
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
