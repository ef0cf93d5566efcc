// TEST INFRASTRUCTURE ONLY — never linked into or called by the product path.
//
// Thin C-ABI harness around the REFERENCE's own generated CPU kernels
// (`seissol::kernel::*`, emitted by /root/reference/codegen/generate.py with the
// generic C++ loop back-end; see oracle/build_ref.py).  It restates only the
// hand-written wrappers that bind pointers into those kernels:
//
//   ssref_ader    <- Spacetime::computeAder              src/Kernels/LinearCK/Time.cpp:65-140
//   ssref_taylor  <- Time::evaluate                      src/Kernels/LinearCK/Time.cpp:237-255
//   ssref_local   <- Local::computeIntegral              src/Kernels/LinearCK/Local.cpp:124-160
//   ssref_neigh   <- Neighbor::computeNeighborsIntegral  src/Kernels/LinearCK/Neighbor.cpp:66-109
//                    (+ TimeCommon::computeIntegrals      src/Kernels/TimeCommon.cpp:40-87)
//   ssref_dr_interp <- DynamicRupture::spaceTimeInterpolation  src/Kernels/DynamicRupture.cpp:57-108
//   kDivMT negation <- OnHost::negateStiffnessMatrix     src/Memory/GlobalData.cpp:51-62
//   proxy loops   <- ProxyKernelHost*::run               src/Proxy/KernelHost.cpp:39-244
//
// All arrays use the reference's layouts (column-major, leading dimension padded
// as the generator decided; sizes are reported by ssref_info()).
// Built into oracle/_ref/libssref_<cfg>.so; one library per (equation, order,
// precision) exactly like the reference builds one binary per configuration.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "init.h"
#include "kernel.h"
#include "tensor.h"
#include "yateto.h"

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace seissol;

#ifdef SSREF_DOUBLE
typedef double real;
#else
typedef float real;
#endif

namespace {
constexpr int Order = SSREF_ORDER;
constexpr unsigned QSize = tensor::Q::size();
constexpr unsigned ISize = tensor::I::size();
constexpr unsigned DerSize = yateto::computeFamilySize<tensor::dQ>();
constexpr unsigned QInterpSize = tensor::QInterpolated::size();

// Negated copies of kDivMT (GlobalData.cpp:51-62).
std::vector<real> negKDivMT[3];
bool initialised = false;

void ensureInit() {
  if (initialised) {
    return;
  }
  for (unsigned i = 0; i < 3; ++i) {
    const unsigned n = init::kDivMT::size(i);
    negKDivMT[i].assign(init::kDivMT::Values[i], init::kDivMT::Values[i] + n);
    for (auto& v : negKDivMT[i]) {
      v *= static_cast<real>(-1.0);
    }
  }
  initialised = true;
}

inline void aderCell(const real* coeffs, real* dofs, const real* star, real* I, real* derivs) {
  alignas(64) real tmp[DerSize];
  real* buf = derivs != nullptr ? derivs : tmp;
  kernel::derivative krnl;
  for (unsigned i = 0; i < 3; ++i) {
    krnl.kDivMT(i) = negKDivMT[i].data();
    krnl.star(i) = star + 24 * i;
  }
  krnl.dQ(0) = dofs;
  for (unsigned i = 1; i < yateto::numFamilyMembers<tensor::dQ>(); ++i) {
    krnl.dQ(i) = buf + yateto::computeFamilySize<tensor::dQ>(1, i);
  }
  krnl.I = I;
  for (int d = 0; d < Order; ++d) {
    krnl.power(d) = coeffs[d];
  }
  if (derivs != nullptr) {
    std::copy_n(dofs, tensor::dQ::size(0), derivs);
  }
  krnl.execute();
}

inline void taylorCell(const real* coeffs, const real* derivs, real* I) {
  kernel::derivativeTaylorExpansion krnl;
  krnl.I = I;
  for (unsigned i = 0; i < yateto::numFamilyMembers<tensor::dQ>(); ++i) {
    krnl.dQ(i) = derivs + yateto::computeFamilySize<tensor::dQ>(1, i);
    krnl.power(i) = coeffs[i];
  }
  krnl.execute();
}

inline void localCell(
    real* dofs, const real* I, const real* star, const real* nApNm1, const uint8_t* faceTypes) {
  kernel::volume vol;
  vol.Q = dofs;
  vol.I = I;
  for (unsigned i = 0; i < 3; ++i) {
    vol.kDivM(i) = init::kDivM::Values[i];
    vol.star(i) = star + 24 * i;
  }
  vol.execute();
  kernel::localFlux lf;
  for (unsigned i = 0; i < 4; ++i) {
    lf.rDivM(i) = init::rDivM::Values[i];
    lf.fMrT(i) = init::fMrT::Values[i];
  }
  lf.Q = dofs;
  lf.I = I;
  lf._prefetch.I = I;
  lf._prefetch.Q = dofs;
  for (unsigned f = 0; f < 4; ++f) {
    if (faceTypes[f] != 3) { // FaceType::DynamicRupture
      lf.AplusT = nApNm1 + 81 * f;
      lf.execute(f);
    }
  }
}

inline void neighCell(real* dofs,
                      const real* const nbI[4],
                      const real* nAmNm1,
                      const uint8_t* faceRel, // [4][2] = {neighbor face j, orientation h}
                      const uint8_t* faceTypes,
                      const real* const godunov[4],
                      const real* const fluxSolver[4],
                      const uint8_t* drSide,
                      const uint8_t* drRel) {
  for (unsigned f = 0; f < 4; ++f) {
    if (faceTypes[f] == 0) { // Regular
      kernel::neighboringFlux nf;
      for (unsigned i = 0; i < 4; ++i) {
        nf.rDivM(i) = init::rDivM::Values[i];
        nf.rT(i) = init::rT::Values[i];
      }
      for (unsigned i = 0; i < 3; ++i) {
        nf.fP(i) = init::fP::Values[i];
      }
      nf.Q = dofs;
      nf.I = nbI[f];
      nf.AminusT = nAmNm1 + 81 * f;
      nf._prefetch.I = nbI[f];
      nf.execute(faceRel[2 * f + 1], faceRel[2 * f + 0], f);
    } else if (faceTypes[f] == 3) { // DynamicRupture
      dynamicRupture::kernel::nodalFlux dr;
      for (unsigned i = 0; i < 16; ++i) {
        dr.V3mTo2nTWDivM.data[i] = init::V3mTo2nTWDivM::Values[i];
      }
      dr.fluxSolver = fluxSolver[f];
      dr.QInterpolated = godunov[f];
      dr.Q = dofs;
      dr._prefetch.I = godunov[f];
      dr.execute(drSide[f], drRel[f]);
    }
  }
}
} // namespace

extern "C" {

// out[0..7] = order, sizeof(real), Q size, I size, derivative family size,
//             QInterpolated size, #basis functions, padded leading dim of Q
void ssref_info(int64_t* out) {
  out[0] = Order;
  out[1] = sizeof(real);
  out[2] = QSize;
  out[3] = ISize;
  out[4] = DerSize;
  out[5] = QInterpSize;
  out[6] = tensor::Q::Shape[0];
  out[7] = QSize / tensor::Q::Shape[1];
}

// offsets (in reals) of dQ(d) inside the derivative family, and sizes
void ssref_der_layout(int64_t* offsets, int64_t* sizes) {
  for (unsigned i = 0; i < yateto::numFamilyMembers<tensor::dQ>(); ++i) {
    offsets[i] = yateto::computeFamilySize<tensor::dQ>(1, i);
    sizes[i] = tensor::dQ::size(i);
  }
}

// static flop counts: {derivative, volume, localFlux[4], neighboringFlux[48], taylor}
void ssref_flops(int64_t* nz, int64_t* hw) {
  int k = 0;
  nz[k] = kernel::derivative::NonZeroFlops;
  hw[k++] = kernel::derivative::HardwareFlops;
  nz[k] = kernel::volume::NonZeroFlops;
  hw[k++] = kernel::volume::HardwareFlops;
  for (int f = 0; f < 4; ++f) {
    nz[k] = kernel::localFlux::NonZeroFlops[f];
    hw[k++] = kernel::localFlux::HardwareFlops[f];
  }
  for (int f = 0; f < 48; ++f) {
    nz[k] = kernel::neighboringFlux::NonZeroFlops[f];
    hw[k++] = kernel::neighboringFlux::HardwareFlops[f];
  }
  nz[k] = kernel::derivativeTaylorExpansion::NonZeroFlops;
  hw[k++] = kernel::derivativeTaylorExpansion::HardwareFlops;
}

void ssref_ader(int64_t n,
                const real* coeffs,
                real* dofs,
                const real* star, // [n][72]
                real* I,          // [n][ISize]
                real* derivs) {   // nullable, [n][DerSize]
  ensureInit();
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < n; ++c) {
    aderCell(coeffs,
             dofs + c * QSize,
             star + c * 72,
             I + c * ISize,
             derivs != nullptr ? derivs + c * DerSize : nullptr);
  }
}

void ssref_taylor(int64_t n, const real* coeffs, const real* derivs, real* I) {
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < n; ++c) {
    taylorCell(coeffs, derivs + c * DerSize, I + c * ISize);
  }
}

void ssref_local(int64_t n,
                 real* dofs,
                 const real* I,
                 const real* star,
                 const real* nApNm1, // [n][4][81]
                 const uint8_t* faceTypes) {
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < n; ++c) {
    localCell(dofs + c * QSize, I + c * ISize, star + c * 72, nApNm1 + c * 324, faceTypes + 4 * c);
  }
}

// Neighbor indices address rows of `Iall` ([*][ISize]); DR faces address rows of
// godunov ([*][QInterpSize]) / fluxSolver ([*][81]) through drIndex[n][4].
void ssref_neigh(int64_t n,
                 real* dofs,
                 const real* Iall,
                 const int64_t* neighbors, // [n][4]
                 const real* nAmNm1,       // [n][4][81]
                 const uint8_t* faceRel,   // [n][4][2]
                 const uint8_t* faceTypes, // [n][4]
                 const real* godunov,
                 const real* fluxSolver,
                 const int64_t* drIndex, // [n][4] or null
                 const uint8_t* drSide,  // [n][4] or null
                 const uint8_t* drRel) { // [n][4] or null
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < n; ++c) {
    const real* nb[4];
    const real* god[4] = {nullptr, nullptr, nullptr, nullptr};
    const real* fs[4] = {nullptr, nullptr, nullptr, nullptr};
    uint8_t side[4] = {0, 0, 0, 0};
    uint8_t rel[4] = {0, 0, 0, 0};
    for (int f = 0; f < 4; ++f) {
      nb[f] = Iall + neighbors[4 * c + f] * ISize;
      if (drIndex != nullptr && faceTypes[4 * c + f] == 3) {
        god[f] = godunov + drIndex[4 * c + f] * QInterpSize;
        fs[f] = fluxSolver + drIndex[4 * c + f] * 81;
        side[f] = drSide[4 * c + f];
        rel[f] = drRel[4 * c + f];
      }
    }
    neighCell(dofs + c * QSize,
              nb,
              nAmNm1 + c * 324,
              faceRel + 8 * c,
              faceTypes + 4 * c,
              god,
              fs,
              side,
              rel);
  }
}

// spaceTimeInterpolation for nFaces fault faces.  derivPlus/Minus: [nFaces][DerSize];
// tinvT: [nFaces][81]; faceInfo: [nFaces][3] = {plusSide, minusSide, faceRelation};
// coeffs: [Order][Order]; outputs [nFaces][Order][QInterpSize].
void ssref_dr_interp(int64_t nFaces,
                     const real* derivPlus,
                     const real* derivMinus,
                     const real* tinvT,
                     const uint8_t* faceInfo,
                     const real* coeffs,
                     real* qPlus,
                     real* qMinus) {
#pragma omp parallel for schedule(static)
  for (int64_t fc = 0; fc < nFaces; ++fc) {
    alignas(64) real dofsPlus[QSize];
    alignas(64) real dofsMinus[QSize];
    dynamicRupture::kernel::evaluateAndRotateQAtInterpolationPoints krnl;
    for (unsigned i = 0; i < 16; ++i) {
      krnl.V3mTo2n.data[i] = init::V3mTo2n::Values[i];
    }
    for (int t = 0; t < Order; ++t) {
      taylorCell(coeffs + t * Order, derivPlus + fc * DerSize, dofsPlus);
      taylorCell(coeffs + t * Order, derivMinus + fc * DerSize, dofsMinus);
      real* qp = qPlus + (fc * Order + t) * QInterpSize;
      real* qm = qMinus + (fc * Order + t) * QInterpSize;
      krnl.QInterpolated = qp;
      krnl.Q = dofsPlus;
      krnl.TinvT = tinvT + 81 * fc;
      krnl._prefetch.QInterpolated = qp;
      krnl.execute(faceInfo[3 * fc + 0], 0);
      krnl.QInterpolated = qm;
      krnl.Q = dofsMinus;
      krnl.TinvT = tinvT + 81 * fc;
      krnl._prefetch.QInterpolated = qm;
      krnl.execute(faceInfo[3 * fc + 1], faceInfo[3 * fc + 2]);
    }
  }
}

// P x P resampling mat-vec of the generated `resampleParameter` kernel
// (LinearSlipWeakening.cpp:18-26); in/out [n][QInterpSize/9].
void ssref_resample(int64_t n, const real* in, real* out) {
  constexpr unsigned P = QInterpSize / 9;
  for (int64_t i = 0; i < n; ++i) {
    dynamicRupture::kernel::resampleParameter k;
    k.resample = init::resample::Values;
    k.originalQ = in + i * P;
    k.resampledQ = out + i * P;
    k.execute();
  }
}

// Proxy pass (src/Proxy/KernelHost.cpp): kernel = 0 all, 1 local, 2 neigh, 3 ader,
// 4 localwoader.  Runs `steps` passes and returns wall seconds (no warm-up here;
// the caller does its own, as Runner.cpp:93-111 does).
double ssref_proxy(int kernel,
                   int64_t n,
                   int steps,
                   const real* coeffs,
                   real* dofs,
                   real* buffers, // I, [n][ISize]
                   const real* star,
                   const real* nApNm1,
                   const real* nAmNm1,
                   const int64_t* neighbors,
                   const uint8_t* faceRel,
                   const uint8_t* faceTypes) {
  ensureInit();
  const auto t0 = std::chrono::steady_clock::now();
  for (int s = 0; s < steps; ++s) {
    if (kernel == 0 || kernel == 1 || kernel == 3 || kernel == 4) {
#pragma omp parallel for schedule(static)
      for (int64_t c = 0; c < n; ++c) {
        if (kernel != 4) {
          aderCell(coeffs, dofs + c * QSize, star + c * 72, buffers + c * ISize, nullptr);
        }
        if (kernel != 3) {
          localCell(dofs + c * QSize,
                    buffers + c * ISize,
                    star + c * 72,
                    nApNm1 + c * 324,
                    faceTypes + 4 * c);
        }
      }
    }
    if (kernel == 0 || kernel == 2) {
      ssref_neigh(n,
                  dofs,
                  buffers,
                  neighbors,
                  nAmNm1,
                  faceRel,
                  faceTypes,
                  nullptr,
                  nullptr,
                  nullptr,
                  nullptr,
                  nullptr);
    }
  }
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

int ssref_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// torchrun exports OMP_NUM_THREADS=1 to every rank: the benchmark's reference arm asks for all host cores explicitly
void ssref_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) {
    omp_set_num_threads(n);
  }
#else
  (void)n;
#endif
}
// Nodal boundary conditions of the local integral (Local::computeIntegral, src/Kernels/LinearCK/Local.cpp:157-252) for m
// boundary faces, one after the other: projectToNodalBoundaryRotated(face) -> boundary condition at the face nodes ->
// localFluxNodal(face).  kind 4 = FaceType::Dirichlet (createEasiBoundaryGhostCells: INodal(l, a) = sum_b map(a, b, l)
// INodal(l, b) + const(a, l)), kind 2 = FaceType::FreeSurfaceGravity (:168-199: the three normal stresses become
// 2 p - sigma with p = -rho g displacement(l)).  Dense inputs: tinv [m][9][9] row-major Tinv(p, q); easiMap
// [m][9][9][B2] (a, b, l); easiConst [m][9][B2]; displacement [m][B2]; rhoG [m].
int64_t ssref_nodal_points() { return tensor::INodal::Shape[0]; }
void ssref_nodal_bc(int64_t m, real* dofs, const real* Iall, const real* nAmNm1, const int64_t* cell, const uint8_t* face,
                    const uint8_t* kind, const real* tinv, const real* easiMap, const real* easiConst,
                    const real* displacement, const double* rhoG) {
  constexpr unsigned B2 = tensor::INodal::Shape[0];
  for (int64_t i = 0; i < m; ++i) {
    alignas(64) real inodal[tensor::INodal::size()] = {};
    alignas(64) real tinvT[tensor::Tinv::size()] = {};
    auto tv = init::Tinv::view::create(tinvT);
    for (unsigned p = 0; p < 9; ++p) {
      for (unsigned q = 0; q < 9; ++q) {
        tv(p, q) = tinv[81 * i + 9 * p + q];
      }
    }
    kernel::projectToNodalBoundaryRotated pk;
    for (unsigned f = 0; f < 4; ++f) {
      pk.V3mTo2nFace(f) = nodal::init::V3mTo2nFace::Values[f];
    }
    pk.I = Iall + cell[i] * ISize;
    pk.INodal = inodal;
    pk.Tinv = tinvT;
    pk.execute(face[i]);
    auto nv = init::INodal::view::create(inodal);
    if (kind[i] == 4) {
      alignas(64) real mapT[tensor::easiBoundaryMap::size()] = {};
      alignas(64) real constT[tensor::easiBoundaryConstant::size()] = {};
      auto mv = init::easiBoundaryMap::view::create(mapT);
      auto cv = init::easiBoundaryConstant::view::create(constT);
      for (unsigned a = 0; a < 9; ++a) {
        for (unsigned l = 0; l < B2; ++l) {
          cv(a, l) = easiConst[(i * 9 + a) * B2 + l];
          for (unsigned b = 0; b < 9; ++b) {
            mv(a, b, l) = easiMap[((i * 9 + a) * 9 + b) * B2 + l];
          }
        }
      }
      kernel::createEasiBoundaryGhostCells ek;
      ek.easiBoundaryMap = mapT;
      ek.easiBoundaryConstant = constT;
      ek.easiIdentMap = init::easiIdentMap::Values;
      ek.INodal = inodal;
      ek.execute();
    } else if (kind[i] == 2) {
      for (unsigned l = 0; l < B2; ++l) {
        const double pressureAtBnd = -1 * rhoG[i] * displacement[i * B2 + l];
        nv(l, 0) = 2 * pressureAtBnd - nv(l, 0);
        nv(l, 1) = 2 * pressureAtBnd - nv(l, 1);
        nv(l, 2) = 2 * pressureAtBnd - nv(l, 2);
      }
    }
    kernel::localFluxNodal lf;
    for (unsigned f = 0; f < 4; ++f) {
      lf.project2nFaceTo3m(f) = init::project2nFaceTo3m::Values[f];
    }
    lf.INodal = inodal;
    lf.AminusT = nAmNm1 + cell[i] * 324 + 81 * face[i];
    lf.Q = dofs + cell[i] * QSize;
    lf._prefetch.I = pk.I;
    lf._prefetch.Q = lf.Q;
    lf.execute(face[i]);
  }
}
}
