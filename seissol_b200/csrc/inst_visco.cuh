// Instantiation body of the viscoelastic2 kernels (kernels_visco.cuh), shared by inst_visco_o5_f{32,64}.cu.
// The including file defines SSB_ORDER, SSB_REAL, SSB_TAG, SSB_VE (elements per CTA) and includes
// gen/ops_o<SSB_ORDER>.cuh before this header.
#include "api_internal.h"
#include "kernels_visco.cuh"

#define SSB_CAT2(a, b) a##b
#define SSB_CAT(a, b) SSB_CAT2(a, b)
#define SSB_GEN SSB_CAT(gen_o, SSB_ORDER)

namespace ssb {
namespace {

struct Gen {
  static constexpr int kOrder = SSB_GEN::kOrder;
  static constexpr int kB = SSB_GEN::kB;
  static constexpr int kBP = SSB_GEN::kBP;
  static constexpr int kB2P = SSB_GEN::kB2P;
  static constexpr int kR1P = ((kOrder - 1) * kOrder * (kOrder + 1) / 6 + 3) / 4 * 4;  // rows of dQext(d), padded
  __host__ __device__ static constexpr int derRows(int d) {
    return (kOrder - d) * (kOrder - d + 1) * (kOrder - d + 2) / 6;
  }
  template <typename R>
  __device__ __forceinline__ static void ader_step_1(R (&out)[kR1P], const R* a0, const R* a1,
                                                     const R* a2, const R (&s)[3][3]) {
    SSB_GEN::ader_step_1<R>(out, a0, a1, a2, s);
  }
  template <typename R>
  __device__ __forceinline__ static void volume(R (&acc)[kBP], const R* a0, const R* a1, const R* a2, const R (&s)[3][3]) {
    SSB_GEN::volume<R>(acc, a0, a1, a2, s);
  }
  template <typename R, typename Src>
  __device__ __forceinline__ static void fMrT(int f, R (&out)[kB2P], const Src& x) { SSB_GEN::fMrT<R>(f, out, x); }
  template <typename R, typename Src>
  __device__ __forceinline__ static void rT(int f, R (&out)[kB2P], const Src& x) { SSB_GEN::rT<R>(f, out, x); }
  template <typename R>
  __device__ __forceinline__ static void fP(int h, R (&out)[kB2P], const R (&x)[kB2P]) { SSB_GEN::fP<R>(h, out, x); }
  template <typename R, typename JS>
  __device__ __forceinline__ static void rDivM(int f, R (&out)[kBP], const JS& js) { SSB_GEN::rDivM<R>(f, out, js); }
};

using Real = SSB_REAL;
constexpr int VE = SSB_VE;
using Plan = ViscoPlan<Real, Gen, VE>;

template <typename K>
int configure(K kern, size_t bytes, SmemConfig& done) {
  return ensureDynamicSmem(kern, bytes, done);
}

int aderEntry(const ViscoLocalParams& p, cudaStream_t stream) {
  static SmemConfig done;
  auto kern = k_visco_ader<Real, Gen, VE>;
  const size_t bytes = Plan::bytes(Plan::aderReals);
  if (int rc = configure(kern, bytes, done)) {
    return rc;
  }
  if (p.n <= 0) {
    return 0;
  }
  kern<<<unsigned((p.n + VE - 1) / VE), Plan::threads, bytes, stream>>>(p);
  return int(cudaGetLastError());
}

int localEntry(const ViscoLocalParams& p, cudaStream_t stream) {
  static SmemConfig done;
  auto kern = k_visco_local<Real, Gen, VE>;
  const size_t bytes = Plan::bytes(Plan::localReals);
  if (int rc = configure(kern, bytes, done)) {
    return rc;
  }
  if (p.n <= 0) {
    return 0;
  }
  kern<<<unsigned((p.n + VE - 1) / VE), Plan::threads, bytes, stream>>>(p);
  return int(cudaGetLastError());
}

struct GenDr {
  static constexpr int kP = SSB_CAT(gen_dr_o, SSB_ORDER)::kP;
  template <typename R>
  __device__ __forceinline__ static const R* nodalFluxTable() { return SSB_CAT(gen_dr_o, SSB_ORDER)::nodalFluxTable<R>(); }
};

int neighDrEntry(const ViscoNeighDrParams& p, cudaStream_t stream) {
  if (p.n <= 0) {
    return 0;
  }
  constexpr int W = 4;   // warps = elements per CTA
  k_visco_neigh_dr<Real, Gen, GenDr, W><<<unsigned((p.n + W - 1) / W), 32 * W, 0, stream>>>(p);
  return int(cudaGetLastError());
}

int neighEntry(const ViscoNeighParams& p, cudaStream_t stream) {
  static SmemConfig done;
  auto kern = k_visco_neigh<Real, Gen, VE>;
  const size_t bytes = Plan::bytes(Plan::neighReals);
  if (int rc = configure(kern, bytes, done)) {
    return rc;
  }
  if (p.n <= 0) {
    return 0;
  }
  kern<<<unsigned((p.n + VE - 1) / VE), Plan::threads, bytes, stream>>>(p);
  return int(cudaGetLastError());
}

const ViscoKernelSet kSet = {Gen::kOrder, int(sizeof(Real)), Gen::kB, Gen::kBP, VE, aderEntry, localEntry, neighEntry,
                             neighDrEntry};

}  // namespace

const ViscoKernelSet* SSB_CAT(viscokernelset_, SSB_TAG)() { return &kSet; }

}  // namespace ssb
