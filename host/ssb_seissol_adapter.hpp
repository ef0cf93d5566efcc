// Adapters from the reference's recorder tables to the flat arrays of include/seissol_b200.h.
//
// This header is compiled INSIDE a SeisSol source tree (it includes the reference's headers; nothing in this repository
// includes it, and it is not part of libseissol_b200.so): define SEISSOL_B200_WITH_REFERENCE_HEADERS and add the
// reference's src/ to the include path.  It turns the ConditionalKey tables the batch recorders fill
// (src/Initializer/BatchRecorders/{Local,Neigh,DynamicRupture}IntegrationRecorder.cpp) into ssb_array_t descriptors, so
// that the bodies of the batched kernel classes become one-line calls:
//
//   ConditionalKey(Time || Volume)                          -> {Dofs, Idofs, Derivatives, LocalIntegrationData}   ssb_ader / ssb_local
//   ConditionalKey(Time, WithLtsBuffers)                    -> {Buffers, Idofs}                                    ssb_lts_accumulate
//   ConditionalKey(NeighborFlux, WithGts|LtsDerivatives)    -> {Derivatives, Idofs}                                ssb_taylor
//   ConditionalKey(DrSpaceMap, side, relation)              -> {QInterpolated±, TinvT, Derivatives±}               ssb_dr_interp
//
// The per-(face, relation) NeighborFlux tables (4 x 48 + 4 x 16 keys) are NOT consumed: ssb_neigh takes the layer arrays the
// recorder derives them from (LTS::FaceNeighborsDevice, CellLocalInformation::{faceTypes, faceRelations}), one launch for all.
#pragma once
#ifdef SEISSOL_B200_WITH_REFERENCE_HEADERS

#include <seissol_b200.h>

#include "Initializer/BatchRecorders/DataTypes/ConditionalKey.h"
#include "Initializer/BatchRecorders/DataTypes/ConditionalTable.h"
#include "Initializer/BatchRecorders/DataTypes/EncodedConstants.h"
#include "Kernels/Precision.h"

namespace seissol_b200::adapter {

using seissol::recording::ComputationKind;
using seissol::recording::ConditionalKey;
using seissol::recording::ConditionalPointersToRealsTable;
using seissol::recording::DrConditionalPointersToRealsTable;
using seissol::recording::KernelNames;
namespace inner_keys = seissol::recording::inner_keys;

inline ssb_array_t none() { return ssb_array_t{nullptr, nullptr, 0}; }

// one variable of a recorded batch: the device array of device pointers (`real**`, DataTypes/Table.h:25-61)
template <typename Table, typename Id>
inline ssb_array_t pointers(Table& batch, Id id) {
  return ssb_array_t{reinterpret_cast<const void* const*>(batch.get(id)->getDeviceDataPtr()), nullptr, 0};
}
template <typename Table, typename Id>
inline long long count(Table& batch, Id id) {
  return static_cast<long long>(batch.get(id)->getSize());
}

// Spacetime::computeBatchedAder (src/Kernels/LinearCK/Time.cpp:142-217)
inline int ader(ssb_ctx* ctx, ConditionalPointersToRealsTable& table, const seissol::real* coeffs, int order, void* stream) {
  const ConditionalKey key(*KernelNames::Time || *KernelNames::Volume);
  auto it = table.find(key);
  if (it == table.end()) {
    return 0;
  }
  auto& e = it->second;
  double c[8];
  for (int d = 0; d < order; ++d) {
    c[d] = static_cast<double>(coeffs[d]);
  }
  using Id = inner_keys::Wp::Id;
  // entries of the Derivatives table may be nullptr (cells without LtsSetup::hasDerivatives): the kernel skips them
  return ssb_ader(ctx, pointers(e, Id::Dofs), pointers(e, Id::Idofs), pointers(e, Id::Derivatives),
                  pointers(e, Id::LocalIntegrationData), c, count(e, Id::Dofs), nullptr, stream);
}

// Local::computeBatchedIntegral (src/Kernels/LinearCK/Local.cpp:255-395); faceTypes: uint8 [n][4] device copy of
// CellLocalInformation::faceTypes in the order of the recorded batch
inline int local(ssb_ctx* ctx, ConditionalPointersToRealsTable& table, const uint8_t* faceTypes, void* stream) {
  const ConditionalKey key(*KernelNames::Time || *KernelNames::Volume);
  auto it = table.find(key);
  if (it == table.end()) {
    return 0;
  }
  auto& e = it->second;
  using Id = inner_keys::Wp::Id;
  return ssb_local(ctx, pointers(e, Id::Dofs), pointers(e, Id::Idofs), pointers(e, Id::LocalIntegrationData), faceTypes,
                   count(e, Id::Dofs), nullptr, stream);
}

// TimeCluster::computeLocalIntegrationDevice, buffer accumulation (src/Solver/TimeStepping/TimeCluster.cpp:508-530)
inline int accumulateBuffers(ssb_ctx* ctx, ConditionalPointersToRealsTable& table, bool reset, void* stream) {
  const ConditionalKey key(*KernelNames::Time, *ComputationKind::WithLtsBuffers);
  auto it = table.find(key);
  if (it == table.end()) {
    return 0;
  }
  auto& e = it->second;
  using Id = inner_keys::Wp::Id;
  return ssb_lts_accumulate(ctx, pointers(e, Id::Buffers), pointers(e, Id::Idofs), reset ? 1 : 0, count(e, Id::Buffers),
                            stream);
}

// TimeCommon::computeBatchedIntegrals (src/Kernels/TimeCommon.cpp:89-125): neighbors read through their derivatives
inline int integrateNeighbors(ssb_ctx* ctx, ConditionalPointersToRealsTable& table, const seissol::real* coeffsGts,
                              const seissol::real* coeffsLts, int order, void* stream) {
  using Id = inner_keys::Wp::Id;
  int rc = 0;
  for (int pass = 0; pass < 2 && rc == 0; ++pass) {
    const ConditionalKey key(*KernelNames::NeighborFlux,
                             pass == 0 ? *ComputationKind::WithGtsDerivatives : *ComputationKind::WithLtsDerivatives);
    auto it = table.find(key);
    if (it == table.end()) {
      continue;
    }
    auto& e = it->second;
    const seissol::real* coeffs = pass == 0 ? coeffsGts : coeffsLts;
    double c[8];
    for (int d = 0; d < order; ++d) {
      c[d] = static_cast<double>(coeffs[d]);
    }
    rc = ssb_taylor(ctx, pointers(e, Id::Derivatives), pointers(e, Id::Idofs), c, count(e, Id::Derivatives), stream);
  }
  return rc;
}

// kernels::DynamicRupture::batchedSpaceTimeInterpolation (src/Kernels/DynamicRupture.cpp:110-168).  The recorder files the
// faces under (DrSpaceMap, side, relation) keys (DynamicRuptureRecorder.cpp:62-88); ssb_dr_interp takes the LAYER arrays
// those tables point into (TimeDerivative{Plus,Minus}Device, DRGodunovData::dataTinvT, DRFaceInformation) in one call.
inline int interpolateFault(ssb_ctx* ctx, const seissol::real* const* derivativesPlus, const seissol::real* const* derivativesMinus,
                            const seissol::real* tinvT, const uint8_t* faceInformation, seissol::real* qInterpolatedPlus,
                            seissol::real* qInterpolatedMinus, const double* pointCoefficients, long long faces, void* stream) {
  return ssb_dr_interp(ctx, ssb_array_t{reinterpret_cast<const void* const*>(derivativesPlus), nullptr, 0},
                       ssb_array_t{reinterpret_cast<const void* const*>(derivativesMinus), nullptr, 0}, tinvT, faceInformation,
                       qInterpolatedPlus, qInterpolatedMinus, pointCoefficients, faces, nullptr, stream);
}

}  // namespace seissol_b200::adapter

#endif  // SEISSOL_B200_WITH_REFERENCE_HEADERS
