// Instantiation body shared by inst_o{4,5,6}_f{32,64}.cu.  The including file defines
//   SSB_ORDER (4|5|6), SSB_REAL (float|double), SSB_TAG (o6_f32), SSB_E_LOCAL, SSB_E_NEIGH
// and includes gen/ops_o<SSB_ORDER>.cuh before this header.
#include "api_internal.h"
#include "kernels_elastic.cuh"
#include "kernels_faceproj.cuh"
#include "kernels_elastic5.cuh"
#ifdef SSB_HAVE_TC
#include "kernels_tc.cuh"
#endif
#include "kernels_nodal.cuh"

#define SSB_CAT2(a, b) a##b
#define SSB_CAT(a, b) SSB_CAT2(a, b)
#define SSB_GEN SSB_CAT(gen_o, SSB_ORDER)

#include <cstdlib>

namespace ssb {
namespace {

// grid size: one CTA per batch, or (SSB_PERSISTENT=k) k CTAs per SM walking the batches in lock step
inline unsigned gridFor(long long batches) {
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

inline int smCount() { return smCountOfCurrentDevice(); }

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

// Kernel generation of the split predictor / local kernels: generation 5 = persistent software-pipelined CTAs
// (kernels_elastic5.cuh), generation 3 = one batch per CTA.  SSB_KGEN5 is a bit mask: 1 = local on generation 5,
// 2 = predictor on generation 5.  Default (measured, profiles/r2_gen5_quick_bench.txt): single precision runs the local
// kernel on generation 5 (3.65 vs 4.13 ms with face projections) and the predictor on generation 3 (1.83 vs 2.03 ms:
// with all 16 warps of an SM issuing at once the instruction fetch of the straight-line stream thrashes more than the
// pipelining gains); double precision stays on generations 3/4 (generation 5 spills there).
inline int gen5Mask() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("SSB_KGEN5");
    v = e != nullptr ? std::atoi(e) : (sizeof(Real) == 4 ? 1 : 0);
  }
  return VL == 1 ? v : 0;
}

#if SSB_PACK_F32 == 1
template <bool ADER>
int launchLocal5(const LocalParams& p, cudaStream_t stream) {
  using P = LocalPlan<Real, Gen, EL>;
  auto kern = ADER ? k_ader5<Real, Gen, EL> : k_local5<Real, Gen, EL>;
  static SmemConfig smemConfig;
  if (int rc = ensureDynamicSmem(kern, P::bytes + smemPad(), smemConfig)) {
    return rc;
  }
  if (p.n <= 0) {
    return 0;
  }
  if (p.n >= (1ll << 31)) {
    return -1;   // SSB_EINVAL: generation 5 keeps element ids in 32 bits (one GPU holds ~2e7 elements)
  }
  const long long batches = (p.n + EL - 1) / EL;
  const long long resident = static_cast<long long>(SSB_MINB_LOCAL5) * smCount();
  const unsigned grid = static_cast<unsigned>(batches < resident ? batches : resident);
  kern<<<grid, blockThreads(EL), P::bytes + smemPad(), stream>>>(p);
  return int(cudaGetLastError());
}
#endif

template <bool ADER, bool LOCAL>
int launchLocal(const LocalParams& p, cudaStream_t stream) {
  using P = LocalPlan<Real, Gen, EL>;
  auto kern = k_local<Real, Gen, EL, ADER, LOCAL>;
  static SmemConfig smemConfig;
  if (int rc = ensureDynamicSmem(kern, P::bytes + smemPad(), smemConfig)) {
    return rc;
  }
  if (p.n <= 0) {
    return 0;
  }
  const unsigned grid = gridFor((p.n + EL * VL - 1) / (EL * VL));
  kern<<<grid, blockThreads(EL), P::bytes + smemPad(), stream>>>(p);
  return int(cudaGetLastError());
}

#ifdef SSB_HAVE_TC
// Tensor-core predictor (kernels_tc.cuh), single precision: SSB_ADER_TC=1 (the CUDA-core kernel is the default).
struct GenTc {
  __device__ __forceinline__ static const float* aderOps() { return SSB_CAT(gen_tc_o, SSB_ORDER)::aderOps(); }
};
inline bool useAderTc() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("SSB_ADER_TC");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
int launchAderTc(const LocalParams& p, cudaStream_t stream) {
  using C = tc::AderTc<Gen>;
  auto kern = tc::k_ader_tc<Gen, GenTc>;
  static SmemConfig smemConfig;
  if (int rc = ensureDynamicSmem(kern, C::bytes, smemConfig)) {
    return rc;
  }
  if (p.n <= 0) {
    return 0;
  }
  const long long tiles = (p.n + tc::kTE - 1) / tc::kTE;
  const long long ctas = (tiles + tc::kNW - 1) / tc::kNW;
  const long long sms = smCount();
  kern<<<unsigned(ctas < sms ? ctas : sms), tc::kNW * tc::kWT, C::bytes, stream>>>(p);
  return int(cudaGetLastError());
}
#endif

int localEntry(const LocalParams& p, bool ader, bool local, cudaStream_t stream) {
  if (ader && local) {
    return launchLocal<true, true>(p, stream);
  }
#ifdef SSB_HAVE_TC
  if (ader && useAderTc()) {
    return launchAderTc(p, stream);
  }
#endif
#if SSB_PACK_F32 == 1
  if (ader && (gen5Mask() & 2)) {
    return launchLocal5<true>(p, stream);
  }
  if (!ader && (gen5Mask() & 1)) {
    return launchLocal5<false>(p, stream);
  }
#endif
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

// elements per CTA of the pipelined face-projection neighbor kernel (k_neigh_fp2): two resident CTAs per SM
constexpr int neighFp2Elems() {
  const int b2p = Gen::kB2P, prj = kNQ * b2p, ten = kNQ * Gen::kBP;
  const int tail = ten > 2 * prj ? pad4i(ten - 2 * prj) : 0;
  const int per = padOdd16(4 * prj + tail, sizeof(Real)) * int(sizeof(Real));
  const int e = 107000 / per;
  return e > 28 ? 28 : e;
}
#ifdef SSB_E_NEIGHFP2
constexpr int ENF2 = SSB_E_NEIGHFP2;
#else
constexpr int ENF2 = neighFp2Elems();
#endif

// 1 = one batch per CTA (k_neigh_fp, default), 2 = persistent pipelined CTAs (k_neigh_fp2; SSB_NEIGHFP=2)
inline int neighFpGeneration() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("SSB_NEIGHFP");
    v = (e != nullptr && e[0] == '2') ? 2 : 1;   // measured equal (2.63 vs 2.68 ms at 1 M elements): the simpler kernel is the default
  }
  return v;
}

#if SSB_PACK_F32 == 1
int neighFp2Entry(const NeighParams& p, cudaStream_t stream) {
  using P = NeighFp2Plan<Real, Gen, ENF2>;
  auto kern = k_neigh_fp2<Real, Gen, ENF2>;
  static SmemConfig smemConfig;
  if (int rc = ensureDynamicSmem(kern, P::bytes + smemPad(), smemConfig)) {
    return rc;
  }
  if (p.n <= 0) {
    return 0;
  }
  const long long batches = (p.n + ENF2 - 1) / ENF2;
  const long long resident = static_cast<long long>(SSB_MINB_NEIGHFP2) * smCount();
  const unsigned grid = static_cast<unsigned>(batches < resident ? batches : resident);
  kern<<<grid, P::threads, P::bytes + smemPad(), stream>>>(p);
  return int(cudaGetLastError());
}
#endif

int neighFpEntry(const NeighParams& p, cudaStream_t stream) {
#if SSB_PACK_F32 == 1
  if (neighFpGeneration() == 2) {
    return neighFp2Entry(p, stream);
  }
  using P = NeighFpPlan<Real, Gen, ENF>;
  auto kern = k_neigh_fp<Real, Gen, ENF>;
  static SmemConfig smemConfig;
  if (int rc = ensureDynamicSmem(kern, P::bytes + smemPad(), smemConfig)) {
    return rc;
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
  static SmemConfig smemConfig;
  if (int rc = ensureDynamicSmem(kern, P::bytes + smemPad(), smemConfig)) {
    return rc;
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
  // k_taylor<Real> is shared by every (order, generation) unit of this precision: one configuration object
  if (int rc = ensureDynamicSmem(k_taylor<Real>, bytes, sharedSmemConfig<Real>())) {
    return rc;
  }
  k_taylor<Real><<<unsigned((p.n + kTaylorE - 1) / kTaylorE), 256, bytes, stream>>>(p);
  return int(cudaGetLastError());
}

int taylorFpEntry(const TaylorParams& p, cudaStream_t stream) {
#if SSB_PACK_F32 == 1
  if (p.n <= 0) {
    return 0;
  }
  const size_t bytes = (size_t(kTaylorFpE) * (p.derSize + kNQ * Gen::kBP + kNQ * Gen::kB2P) * sizeof(Real) + 15) / 16 * 16 + 16;
  static SmemConfig smemConfig;
  if (int rc = ensureDynamicSmem(k_taylor_fp<Real, Gen>, bytes, smemConfig)) {
    return rc;
  }
  k_taylor_fp<Real, Gen><<<unsigned((p.n + kTaylorFpE - 1) / kTaylorFpE), 256, bytes, stream>>>(p);
  return int(cudaGetLastError());
#else
  return -2;
#endif
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

struct GenNodal {
  static constexpr int kB2 = SSB_CAT(gen_nodal_o, SSB_ORDER)::kB2;
  template <typename R>
  __device__ __forceinline__ static const R* nodalV() { return SSB_CAT(gen_nodal_o, SSB_ORDER)::nodalV<R>(); }
  template <typename R>
  __device__ __forceinline__ static const R* nodalP() { return SSB_CAT(gen_nodal_o, SSB_ORDER)::nodalP<R>(); }
  __device__ __forceinline__ static const double* massDiag() { return SSB_CAT(gen_nodal_o, SSB_ORDER)::massDiag(); }
};

int nodalBcEntry(const NodalBcParams& p, cudaStream_t stream) {
  if (p.m <= 0) {
    return 0;
  }
  constexpr int W = 4;
  k_local_nodal_bc<Real, Gen, GenNodal, W><<<unsigned((p.m + W - 1) / W), 32 * W, 0, stream>>>(p);
  return int(cudaGetLastError());
}

int volumeEnergiesEntry(const VolumeEnergyParams& p, cudaStream_t stream) {
  if (p.n <= 0) {
    return 0;
  }
  constexpr int perBlock = 8 * (sizeof(Real) == 4 ? 2 : 1);   // elements per CTA and iteration
  long long blocks = (p.n + perBlock - 1) / perBlock;
  const long long resident = 8LL * smCount();   // 64 warps per SM stream the elements
  if (blocks > resident) {
    blocks = resident;
  }
  k_volume_energies<Real, Gen, GenNodal><<<unsigned(blocks), 256, 0, stream>>>(p);
  return int(cudaGetLastError());
}

const KernelSet kSet = {Gen::kOrder, int(sizeof(Real)), Gen::kB, Gen::kBP, EL * VL, EN * VL, localEntry, neighEntry, taylorEntry, accumulateEntry, taylorFpEntry,
                        nodalBcEntry, volumeEnergiesEntry};

}  // namespace

const KernelSet* SSB_CAT(kernelset_, SSB_TAG)() { return &kSet; }

}  // namespace ssb
