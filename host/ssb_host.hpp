// C++ host side of seissol_b200 above the C ABI (include/seissol_b200.h): the reference's batched kernel classes
// with their names and argument meaning, bound to one `Layer` of device arrays instead of the recorder tables.
//
//   kernels::Spacetime::computeBatchedAder          <- src/Kernels/Spacetime.h:60-67, LinearCK/Time.cpp:142-217
//   kernels::Time::evaluateBatched                  <- src/Kernels/Time.h:56-60,      LinearCK/Time.cpp:257-292
//   kernels::Local::computeBatchedIntegral          <- src/Kernels/Local.h:64-68,     LinearCK/Local.cpp:255-395
//   kernels::Neighbor::computeBatchedNeighborsIntegral <- src/Kernels/Neighbor.h:49-50, LinearCK/Neighbor.cpp:111-188
//
// Error behaviour follows the reference: device errors are fatal (`logError()` / APIWRAP abort) -> std::runtime_error.
// There is no CPU path: constructing a Context without a CUDA device throws.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>

#include "seissol_b200.h"

namespace seissol_b200 {

inline void check(int rc, const char* what) {
  if (rc != 0) {
    throw std::runtime_error(std::string(what) + ": " + ssb_error_string(rc));
  }
}
inline void checkCuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
  }
}

// One (ConvergenceOrder, real, equation) configuration (the reference builds one binary per configuration).
class Context {
  public:
  Context(int order, int realSize, bool referenceLayout = false, int equation = 0, int mechanisms = 0)
      : order_(order), realSize_(realSize) {
    check(ssb_create(&ctx_, order, realSize, equation, mechanisms), "ssb_create");
    if (referenceLayout) {
      check(ssb_layout_reference(ctx_), "ssb_layout_reference");
    }
    int q[24];
    check(ssb_query(ctx_, q), "ssb_query");
    basisFunctions_ = q[0];
    ld_ = q[1];
    dofsSize_ = q[2];
    derivativesSize_ = q[3];
    faultPoints_ = q[22];
    faultPointsPadded_ = q[23];
  }
  ~Context() { ssb_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  ssb_ctx* handle() const { return ctx_; }
  int order() const { return order_; }
  int realSize() const { return realSize_; }
  int basisFunctions() const { return basisFunctions_; }
  int leadingDimension() const { return ld_; }
  std::size_t dofsSize() const { return dofsSize_; }             // reals per Q / I tensor
  std::size_t derivativesSize() const { return derivativesSize_; }
  int faultPoints() const { return faultPoints_; }
  int faultPointsPadded() const { return faultPointsPadded_; }

  private:
  ssb_ctx* ctx_{nullptr};
  int order_, realSize_, basisFunctions_{0}, ld_{0}, faultPoints_{0}, faultPointsPadded_{0};
  std::size_t dofsSize_{0}, derivativesSize_{0};
};

// Device arrays of one layer = the slice of the LTS storage one time cluster works on
// (LTS::{Dofs, Buffers, Derivatives, LocalIntegration, NeighboringIntegration, CellInformation, FaceNeighbors},
// src/Memory/Descriptor/LTS.h:97-119).  Arrays are contiguous per variable like Storage<Varmap> holds them.
struct Layer {
  std::size_t size{0};
  int realSize{4};
  void* dofs{nullptr};
  void* buffers{nullptr};                 // time-integrated DOFs I
  void* derivatives{nullptr};             // optional
  void* localIntegration{nullptr};        // [size][396] star[3][24] | nApNm1[4][81]
  void* neighboringIntegration{nullptr};  // [size][324] nAmNm1[4][81]
  std::uint8_t* faceTypes{nullptr};       // [size][4]
  std::uint8_t* faceRelations{nullptr};   // [size][4][2] {neighbor's face, orientation}
  long long* faceNeighbors{nullptr};      // [size][4] index into buffers (proxy) ...
  const void* const* faceNeighborPtrs{nullptr};  // ... or LTS::FaceNeighborsDevice pointers
  int* permutation{nullptr};              // optional processing order of the neighbor kernel
  // optional face projections F_f = fMrT_f I, [size][4][9][B2P] (include/seissol_b200.h "face-projection path"): when
  // set, Local stores them and Neighbor gathers them (faceNeighbors then index this array) instead of the neighbors' I
  void* faceProjections{nullptr};
  std::size_t faceProjectionsStride{0};   // reals per cell
  std::size_t dofsStride{0}, derivativesStride{0};  // reals per cell
};

inline ssb_array_t strided(const void* base, std::size_t strideBytes) {
  return ssb_array_t{nullptr, base, static_cast<long long>(strideBytes)};
}
inline ssb_array_t none() { return ssb_array_t{nullptr, nullptr, 0}; }

namespace kernels {

class Spacetime {
  public:
  explicit Spacetime(Context& ctx) : ctx_(ctx) {}
  // coeffs: Taylor integration coefficients of [0, dt] (src/Numerical/TimeBasis.h:82-96), host array
  void computeBatchedAder(const double* coeffs, double /*timeStepWidth*/, Layer& layer, cudaStream_t stream) {
    const std::size_t rs = layer.realSize;
    check(ssb_ader(ctx_.handle(), strided(layer.dofs, layer.dofsStride * rs), strided(layer.buffers, layer.dofsStride * rs),
                   layer.derivatives != nullptr ? strided(layer.derivatives, layer.derivativesStride * rs) : none(),
                   strided(layer.localIntegration, 396 * rs), coeffs, static_cast<long long>(layer.size), nullptr, stream),
          "ssb_ader");
  }

  private:
  Context& ctx_;
};

class Time {
  public:
  explicit Time(Context& ctx) : ctx_(ctx) {}
  void evaluateBatched(const double* coeffs, const void* const* timeDerivatives, void* const* timeIntegratedDofs,
                       std::size_t numElements, cudaStream_t stream) {
    check(ssb_taylor(ctx_.handle(), ssb_array_t{timeDerivatives, nullptr, 0},
                     ssb_array_t{const_cast<const void* const*>(timeIntegratedDofs), nullptr, 0}, coeffs,
                     static_cast<long long>(numElements), stream),
          "ssb_taylor");
  }

  private:
  Context& ctx_;
};

class Local {
  public:
  explicit Local(Context& ctx) : ctx_(ctx) {}
  void computeBatchedIntegral(Layer& layer, double /*timeStepWidth*/, cudaStream_t stream) {
    const std::size_t rs = layer.realSize;
    if (layer.faceProjections != nullptr) {
      check(ssb_local_faceproj(ctx_.handle(), strided(layer.dofs, layer.dofsStride * rs),
                               strided(layer.buffers, layer.dofsStride * rs), strided(layer.localIntegration, 396 * rs),
                               strided(layer.faceProjections, layer.faceProjectionsStride * rs), layer.faceTypes,
                               static_cast<long long>(layer.size), nullptr, stream),
            "ssb_local_faceproj");
      return;
    }
    check(ssb_local(ctx_.handle(), strided(layer.dofs, layer.dofsStride * rs), strided(layer.buffers, layer.dofsStride * rs),
                    strided(layer.localIntegration, 396 * rs), layer.faceTypes, static_cast<long long>(layer.size), nullptr,
                    stream),
          "ssb_local");
  }

  // Boundary faces with a nodal boundary condition (the second half of Local::computeBatchedIntegral, Local.cpp:350-395,
  // through DirichletBoundary::evaluateOnDevice): one face list (cell, face, FaceType) for Dirichlet (4) and
  // FreeSurfaceGravity (2) faces; all arrays are device arrays (include/seissol_b200.h, ssb_local_nodal_bc).
  struct BoundaryFaces {
    const long long* cell{nullptr};
    const std::uint8_t* face{nullptr};
    const std::uint8_t* kind{nullptr};
    const void* tinv{nullptr};            // [m][9][9]
    const void* easiMap{nullptr};         // [m][9][9][B2]
    const void* easiConstant{nullptr};    // [m][9][B2]
    const void* displacement{nullptr};    // [m][B2]
    const void* rhoG{nullptr};            // [m]
    std::size_t size{0};
  };
  void computeBatchedNodalBoundaryConditions(Layer& layer, const BoundaryFaces& bc, cudaStream_t stream) {
    const std::size_t rs = layer.realSize;
    check(ssb_local_nodal_bc(ctx_.handle(), strided(layer.dofs, layer.dofsStride * rs),
                             strided(layer.buffers, layer.dofsStride * rs), strided(layer.neighboringIntegration, 324 * rs),
                             bc.cell, bc.face, bc.kind, bc.tinv, bc.easiMap, bc.easiConstant, bc.displacement, bc.rhoG,
                             static_cast<long long>(bc.size), stream),
          "ssb_local_nodal_bc");
  }
  // Local::evaluateBatchedTimeDependentBc (Local.cpp:397-455): FaceType::Analytical faces.  The host loop that evaluates
  // the analytical solution at the face nodes and integrates it over the time step stays with the caller (it runs user
  // callbacks); analyticScratch is its result on the device, [m][9][B2].
  void evaluateBatchedTimeDependentBc(Layer& layer, const long long* cells, const std::uint8_t* faces,
                                      const void* analyticScratch, std::size_t numFaces, cudaStream_t stream) {
    const std::size_t rs = layer.realSize;
    check(ssb_local_nodal_flux(ctx_.handle(), strided(layer.dofs, layer.dofsStride * rs),
                               strided(layer.neighboringIntegration, 324 * rs), cells, faces, analyticScratch,
                               static_cast<long long>(numFaces), stream),
          "ssb_local_nodal_flux");
  }

  private:
  Context& ctx_;
};

class Neighbor {
  public:
  explicit Neighbor(Context& ctx) : ctx_(ctx) {}
  void computeBatchedNeighborsIntegral(Layer& layer, cudaStream_t stream) {
    const std::size_t rs = layer.realSize;
    if (layer.faceProjections != nullptr && layer.faceNeighborPtrs == nullptr) {
      check(ssb_neigh_faceproj(ctx_.handle(), strided(layer.dofs, layer.dofsStride * rs), nullptr, layer.faceNeighbors,
                               strided(layer.faceProjections, layer.faceProjectionsStride * rs),
                               strided(layer.neighboringIntegration, 324 * rs), layer.faceTypes, layer.faceRelations,
                               static_cast<long long>(layer.size), layer.permutation, stream),
            "ssb_neigh_faceproj");
      return;
    }
    check(ssb_neigh(ctx_.handle(), strided(layer.dofs, layer.dofsStride * rs), layer.faceNeighborPtrs, layer.faceNeighbors,
                    strided(layer.buffers, layer.dofsStride * rs), strided(layer.neighboringIntegration, 324 * rs),
                    layer.faceTypes, layer.faceRelations, static_cast<long long>(layer.size), layer.permutation, stream),
          "ssb_neigh");
  }

  private:
  Context& ctx_;
};

// viscoelastic2 (src/Kernels/LinearCKAnelastic/*): the same layer plus the memory variables (LTS::DofsAne) and their time
// integral; localIntegration = star[3][33] | nApNm1[4][135] | E[162] | w[3] | W[3], neighboringIntegration =
// nAmNm1[4][135] | w[3] (src/Equations/viscoelastic2/Model/IntegrationData.h)
struct ViscoLayer : Layer {
  void* dofsAne{nullptr};
  void* buffersAne{nullptr};        // LocalTmp::timeIntegratedAne of the reference, per cell
  std::size_t aneStride{0};         // reals per cell
};

namespace anelastic {
class Spacetime {
  public:
  explicit Spacetime(Context& ctx) : ctx_(ctx) {}
  void computeBatchedAder(const double* coeffs, double, ViscoLayer& l, cudaStream_t stream) {
    const std::size_t rs = l.realSize;
    check(ssb_visco_ader(ctx_.handle(), strided(l.dofs, l.dofsStride * rs), strided(l.dofsAne, l.aneStride * rs),
                         strided(l.buffers, l.dofsStride * rs), strided(l.buffersAne, l.aneStride * rs),
                         strided(l.localIntegration, 807 * rs), coeffs, static_cast<long long>(l.size), stream),
          "ssb_visco_ader");
  }

  private:
  Context& ctx_;
};
class Local {
  public:
  explicit Local(Context& ctx) : ctx_(ctx) {}
  void computeBatchedIntegral(ViscoLayer& l, double, cudaStream_t stream) {
    const std::size_t rs = l.realSize;
    check(ssb_visco_local(ctx_.handle(), strided(l.dofs, l.dofsStride * rs), strided(l.dofsAne, l.aneStride * rs),
                          strided(l.buffers, l.dofsStride * rs), strided(l.buffersAne, l.aneStride * rs),
                          strided(l.localIntegration, 807 * rs), l.faceTypes, static_cast<long long>(l.size), stream),
          "ssb_visco_local");
  }

  private:
  Context& ctx_;
};
class Neighbor {
  public:
  explicit Neighbor(Context& ctx) : ctx_(ctx) {}
  void computeBatchedNeighborsIntegral(ViscoLayer& l, cudaStream_t stream) {
    const std::size_t rs = l.realSize;
    check(ssb_visco_neigh(ctx_.handle(), strided(l.dofs, l.dofsStride * rs), strided(l.dofsAne, l.aneStride * rs),
                          l.faceNeighborPtrs, l.faceNeighbors, strided(l.buffers, l.dofsStride * rs),
                          strided(l.neighboringIntegration, 543 * rs), l.faceTypes, l.faceRelations,
                          static_cast<long long>(l.size), l.permutation, stream),
          "ssb_visco_neigh");
  }

  private:
  Context& ctx_;
};
}  // namespace anelastic

// Fault-side device arrays of one DynamicRupture::Layer (src/Memory/Descriptor/DynamicRupture.h:47-157) + the
// CellDRMapping of the adjacent cells (src/Initializer/Typedefs.h:265-270)
struct FaultLayer {
  std::size_t faces{0};
  const void* const* timeDerivativePlus{nullptr};    // [faces] derivative family of the plus-side cell
  const void* const* timeDerivativeMinus{nullptr};
  void* tinvT{nullptr};                              // [faces][81]
  std::uint8_t* faceInformation{nullptr};            // [faces][3] {plusSide, minusSide, faceRelation}
  void* qInterpolatedPlus{nullptr};                  // [faces][O][9][ppad]
  void* qInterpolatedMinus{nullptr};
  int* permutation{nullptr};
  // per cell and face
  const void* const* godunov{nullptr};               // [cells][4]
  const void* const* fluxSolver{nullptr};            // [cells][4]
  std::uint8_t* sideRelation{nullptr};               // [cells][4][2] {side, faceRelation}
  std::uint8_t* drFaceTypes{nullptr};                // [cells][4]
};

class DynamicRupture {
  public:
  explicit DynamicRupture(Context& ctx) : ctx_(ctx) {}
  // src/Kernels/DynamicRupture.h:48-50; coeffs: [O][O] point-evaluation rows (DynamicRupture.cpp:62-70), host array
  void batchedSpaceTimeInterpolation(FaultLayer& fault, const double* coeffs, cudaStream_t stream) {
    check(ssb_dr_interp(ctx_.handle(), ssb_array_t{fault.timeDerivativePlus, nullptr, 0},
                        ssb_array_t{fault.timeDerivativeMinus, nullptr, 0}, fault.tinvT, fault.faceInformation,
                        fault.qInterpolatedPlus, fault.qInterpolatedMinus, coeffs, static_cast<long long>(fault.faces),
                        fault.permutation, stream),
          "ssb_dr_interp");
  }
  // DynamicRupture branch of Neighbor::computeBatchedNeighborsIntegral (LinearCK/Neighbor.cpp:150-186)
  void computeBatchedNodalFlux(Layer& layer, FaultLayer& fault, cudaStream_t stream) {
    const std::size_t rs = layer.realSize;
    check(ssb_neigh_dr(ctx_.handle(), strided(layer.dofs, layer.dofsStride * rs), fault.godunov, fault.fluxSolver,
                       fault.sideRelation, fault.drFaceTypes, static_cast<long long>(layer.size), stream),
          "ssb_neigh_dr");
  }

  private:
  Context& ctx_;
};

}  // namespace kernels

// Volume part of writer::EnergyOutput::computeVolumeEnergies (src/ResultWriter/EnergyOutput.cpp:347-486): the OpenMP loop
// over the cells of a layer as one launch.  volumes [size] and materials [size][3] = lambda, mu, rho are device arrays of
// doubles; energies is a device double[8] the caller zeroes before the first layer and reads back after the last
// (elastic kinetic, elastic, acoustic kinetic, acoustic, momentum x / y / z, unused).
class EnergyOutput {
  public:
  explicit EnergyOutput(Context& ctx) : ctx_(ctx) {}
  void computeVolumeEnergies(Layer& layer, const double* volumes, const double* materials, double* energies,
                             cudaStream_t stream) {
    const std::size_t rs = layer.realSize;
    check(ssb_volume_energies(ctx_.handle(), strided(layer.dofs, layer.dofsStride * rs), volumes, materials,
                              static_cast<long long>(layer.size), energies, stream),
          "ssb_volume_energies");
  }

  private:
  Context& ctx_;
};

}  // namespace seissol_b200
