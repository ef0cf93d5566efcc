// Plain parameter blocks handed from the C ABI (capi.cu) to the kernel launchers of the per-configuration translation
// units.  Kept apart from the kernel headers so that a change to one kernel family recompiles only its own units.
#pragma once
#include "common.cuh"

namespace ssb {

constexpr int kVM = 3;  // viscoelastic2: mechanisms of the compiled configuration

// ---------------------------------------------------------------------------------------
struct LocalParams {
  ElemArray dofs;      // Q, in/out                         (LTS::Dofs)
  ElemArray idofs;     // I: out if ader, in if !ader       (LTS::Buffers / Idofs scratch)
  ElemArray derivs;    // optional derivative family, out   (LTS::Derivatives)
  ElemArray localInt;  // LocalIntegrationData: star[3][24], nApNm1[4][81]
  ElemArray faceProj;  // optional out: F_f = fMrT_f . I, [4][9][B2P] per element (kernels_faceproj.cuh)
  const uint8_t* faceTypes;  // [n][4] FaceType, may be null (= all regular)
  const int* perm;           // optional processing order
  long long n;
  int ldg;             // leading dimension of Q/I columns in global memory (reals)
  int derOffset[8];    // offset of dQ(d) inside the derivative family (reals)
  int derLd[8];        // leading dimension of dQ(d) in global memory (reals)
  double coeffs[8];    // Taylor integration coefficients c_d
};

struct NeighParams {
  ElemArray dofs;       // Q in/out
  ElemArray neighInt;   // NeighboringIntegrationData: nAmNm1[4][81]
  const void* const* nbPtrs;  // [n][4] pointers to the neighbors' I (LTS::FaceNeighbors) or null
  const long long* nbIdx;     // else [n][4] indices into idofs
  ElemArray idofs;
  ElemArray faceProj;         // face-projection path: nbIdx indexes this array; nbPtrs point to the neighbors' blocks
  const uint8_t* faceTypes;   // [n][4] or null
  const uint8_t* faceRel;     // [n][4][2] = {neighbor's face j, orientation h}
  const int* perm;
  long long n;
  int ldg;
  int useFaceProj;            // nbPtrs / nbIdx refer to face projections (ssb_neigh_faceproj)
};

struct TaylorParams {
  ElemArray derivs;
  ElemArray idofs;       // out: I (k_taylor) or the face projection fMrT_j I, [9][B2P] (k_taylor_fp)
  const uint8_t* faceJ;  // k_taylor_fp: [n] local face id j of the element whose derivatives are integrated
  long long n;
  int ldg;
  int order;
  int derSize;  // reals of one derivative family (contiguous per element)
  int rows[8];
  int derOffset[8];
  int derLd[8];
  double coeffs[8];
};


struct AccumulateParams {
  ElemArray buffers;
  ElemArray idofs;
  long long n;
  int vec4PerElem;  // row size / 4 reals
  int reset;
};

struct DrInterpParams {
  ElemArray derPlus;        // per FACE: derivative family of the plus-side element (DynamicRupture::TimeDerivativePlus)
  ElemArray derMinus;       // per FACE: minus side
  const void* tinvT;        // [n][81] DRGodunovData::dataTinvT, column-major TinvT[a + 9 b]
  const uint8_t* faceInfo;  // [n][3] DRFaceInformation {plusSide, minusSide, faceRelation}
  void* qPlus;              // [n][O][9][ppad] DynamicRupture::QInterpolatedPlus
  void* qMinus;
  const int* perm;
  long long n;
  int ppad;
  int rawStride;            // shared-memory reals per face (>= derSize, odd multiple of 16 bytes)
  int derSize;
  int derOffset[8];
  int derLd[8];
  double coeffs[8][8];      // [time point][derivative]: tau^d / d!
};

// FrictionLawData of the reference as a structure of device arrays, all [n][ppad] unless noted
// (src/Memory/Descriptor/DynamicRupture.h:47-265).
struct FrictionParams {
  const void* qPlus;        // [n][O][9][ppad]
  const void* qMinus;
  void* imposedPlus;        // [n][9][ppad]
  void* imposedMinus;
  const void* impAndEta;    // [n][11]: zp zs zpNeig zsNeig etaP etaS invEtaS invZp invZs invZpNeig invZsNeig
  const void* initialStress;  // [n][6][ppad]
  void* mu; void* slip1; void* slip2; void* slipRate1; void* slipRate2; void* slipRateMagnitude;
  void* accumulatedSlipMagnitude; void* traction1; void* traction2; void* ruptureTime; void* peakSlipRate;
  void* dynStressTime;
  uint8_t* ruptureTimePending; uint8_t* dynStressTimePending;
  // linear slip weakening
  const void* dC; const void* muS; const void* muD; const void* cohesion; const void* forcedRuptureTime;
  // rate and state, fast velocity weakening
  const void* rsA; const void* rsSl0; const void* rsF0; const void* rsMuW; const void* rsB; const void* rsSrW;
  void* stateVariable; uint8_t* convergenceInner; uint8_t* convergenceOuter;
  const void* resampleT;    // [P][P]: resampleT[j*P + p] = resample[p][j]
  long long n;
  int ppad;
  double deltaT[8];
  double timeWeights[8];
  double fullUpdateTime;
  double etaPDamp;
  double healingThreshold;
  double t0;
  double rsSr0;
  double rsSlipRateTolerance;
  double rsStateTolerance;
  int rsMaxNumberSlipRateUpdates;
  int rsNumberStateVariableUpdates;
  // nucleation (common::adjustInitialStress) and initial pressure; initialStress is updated in place when nucleationCount > 0
  void* initialPressure;            // [n][ppad] or null (= 0, not written)
  const void* nucleationStress;     // [n][nucleationCount][6][ppad]
  const void* nucleationPressure;   // [n][nucleationCount][ppad]
  int nucleationCount;
  double nucleationT0[4];
  double nucleationS0[4];
  // friction energy output (common::computeFrictionEnergy); frictionalEnergy == null: off
  void* energySlip;                 // [n][3][ppad]
  void* energyAccumulatedSlip;      // [n][ppad]
  void* frictionalEnergy;           // [n][ppad]
  const void* spaceWeights;         // [P]
  const void* doubledSurfaceArea;   // [n]
  int energiesFromAcrossFaultVelocities;
  int rsFloatResolutionExit;        // single precision: accept a Newton iterate at the rounding level of its operands
};

struct NeighDrParams {
  ElemArray dofs;
  const void* const* godunov;      // [n][4] imposed state of the fault face (CellDRMapping::godunov) or null
  const void* const* fluxSolver;   // [n][4] 9x9 flux solver (CellDRMapping::fluxSolver), column-major [c + 9 q]
  const uint8_t* sideRel;          // [n][4][2] {side, faceRelation}
  const uint8_t* faceTypes;        // [n][4]
  long long n;
  int ldg;
  int ppad;
};


struct ViscoLocalParams {
  ElemArray dofs, dofsAne, idofs, idofsAne, localInt;
  const uint8_t* faceTypes;
  long long n;
  int ldg;
  double coeffs[8];
};

struct ViscoNeighDrParams {
  ElemArray dofs, dofsAne, neighInt;
  const void* const* godunov;      // [n][4] imposed state of the fault face, [9][ppad] (CellDRMapping::godunov) or null
  const void* const* fluxSolver;   // [n][4] 9 x 15 flux solver of the extended quantities, column-major [c + 9 q]
  const uint8_t* sideRel;          // [n][4][2] {side, faceRelation}
  const uint8_t* faceTypes;        // [n][4]
  long long n;
  int ldg;
  int ppad;
};

struct ViscoNeighParams {
  ElemArray dofs, dofsAne, idofs, neighInt;
  const void* const* nbPtrs;
  const long long* nbIdx;
  const uint8_t* faceTypes;
  const uint8_t* faceRel;
  const int* perm;
  long long n;
  int ldg;
};


}  // namespace ssb
