// TEST INFRASTRUCTURE ONLY.  Exposes the reference's rate-and-state helper functions
// (src/DynamicRupture/FrictionLaws/RateAndStateCommon.h, included from the mounted reference
// tree at build time) through a C ABI so that oracle/friction.py can be pinned against them.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>

#include "DynamicRupture/FrictionLaws/RateAndStateCommon.h"

using seissol::real;
namespace rs = seissol::dr::friction_law::rs;

extern "C" {
void ssref_rs_arsinhexp(int64_t n, const real* x, const real* expLog, const real* exp, real* out, real* dout) {
  for (int64_t i = 0; i < n; ++i) {
    out[i] = rs::arsinhexp<real>(x[i], expLog[i], exp[i]);
    dout[i] = rs::arsinhexpDerivative<real>(x[i], expLog[i], exp[i]);
  }
}
void ssref_rs_cexp(int64_t n, const real* c, real* out) {
  for (int64_t i = 0; i < n; ++i) {
    out[i] = rs::computeCExp<real>(c[i]);
  }
}
void ssref_rs_logsinh(int64_t n, const real* x, const real* c, real* out) {
  for (int64_t i = 0; i < n; ++i) {
    out[i] = rs::logsinh<real>(x[i], c[i]);
  }
}
double ssref_rs_almost_zero() { return static_cast<double>(rs::almostZero()); }
}
