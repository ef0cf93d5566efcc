// TEST INFRASTRUCTURE ONLY: minimal stand-in for the reference's generated Config.h so that
// /root/reference/src/DynamicRupture/FrictionLaws/RateAndStateCommon.h (included verbatim by
// oracle/ref_rs.cpp, never copied) finds `Config::Precision`.
#pragma once
#include "Common/Real.h"
namespace seissol {
struct Config {
#ifdef SSREF_DOUBLE
  static constexpr RealType Precision = RealType::F64;
#else
  static constexpr RealType Precision = RealType::F32;
#endif
};
}  // namespace seissol
