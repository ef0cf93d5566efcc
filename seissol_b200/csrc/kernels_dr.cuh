// Hand-written sm_100a kernels for the dynamic-rupture part of the element update.
//
//   k_dr_interp   : space-time interpolation of both sides' wave fields at the fault quadrature points
//                   (proxy kernel `godunov_dr`).  Replaces kernels::DynamicRupture::batchedSpaceTimeInterpolation
//                   (src/Kernels/DynamicRupture.cpp:110-168; per-face twin :57-108).
//   k_friction_lsw / k_friction_rsfvw : friction solvers, one thread per fault quadrature point with a fused
//                   Newton solve.  Replace FrictionSolver::evaluate for LinearSlipWeakeningLaw<NoSpecialization>
//                   and RateAndStateBase<FastVelocityWeakeningLaw<NoTP>> (src/DynamicRupture/FrictionLaws/
//                   CpuImpl/{BaseFrictionLaw,LinearSlipWeakening,RateAndState,FastVelocityWeakeningLaw}.h,
//                   FrictionSolverCommon.h; GPU twins under GpuImpl/).  Convergence is decided per FACE like the
//                   CPU implementation (all points of a face iterate until all have converged).
//   k_neigh_dr    : nodal flux of the imposed state back into the elements (proxy kernel `neigh_dr`).  Replaces
//                   the DynamicRupture branch of Neighbor::computeBatchedNeighborsIntegral
//                   (src/Kernels/LinearCK/Neighbor.cpp:150-186; dynamicRupture::kernel::nodalFlux).
#pragma once
#include "common.cuh"
#include "params.cuh"

namespace ssb {

// ---------------------------------------------------------------------------------------
template <typename R, typename G, typename GD, int NF>
struct DrInterpCfg {
  using PT = typename PkDr<R>::T;
  static constexpr int V = PkDr<R>::V;
  static constexpr int O = G::kOrder;
  static constexpr int NSG = (O + V - 1) / V;   // groups of V time points
  static constexpr int TPF = 9 * NSG;           // threads per face
  static constexpr int threads = (NF * TPF + 31) / 32 * 32;
};

template <typename R, typename G, typename GD, int NF>
__global__ void __launch_bounds__((DrInterpCfg<R, G, GD, NF>::threads)) k_dr_interp(const DrInterpParams p) {
  using C = DrInterpCfg<R, G, GD, NF>;
  using PT = typename C::PT;
  constexpr int V = C::V, O = C::O, BP = G::kBP, PP = GD::kPP, TPF = C::TPF;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  const int tid = threadIdx.x;
  const bool active = tid < NF * TPF;
  const int f = active ? tid % NF : 0;
  const int j = active ? tid / NF : 0;  // 0..TPF-1
  const int q = j % 9;
  const int sg = j / 9;
  // two derivative buffers per face: the minus side's family is in flight while the plus side is computed
  R* const raw0 = smem + f * p.rawStride;
  R* const raw1 = smem + (NF + f) * p.rawStride;
  R* const tmat = smem + 2 * NF * p.rawStride + f * 84;
  uint64_t* const bar = reinterpret_cast<uint64_t*>(
      smemRaw + ((size_t(NF) * (2 * p.rawStride + 84) * sizeof(R) + 15) / 16) * 16);   // bar[0]: plus, bar[1]: minus

  if (tid == 0) {
    mbar_init(bar, NF);
    mbar_init(bar + 1, NF);
    mbar_fence_init();
  }
  __syncthreads();

  const long long slot = static_cast<long long>(blockIdx.x) * NF + f;
  const long long face = slot < p.n ? (p.perm != nullptr ? p.perm[slot] : slot) : -1;
  uint8_t info[3] = {0, 0, 0};
  if (face >= 0) {
    info[0] = p.faceInfo[3 * face]; info[1] = p.faceInfo[3 * face + 1]; info[2] = p.faceInfo[3 * face + 2];
    // 9x9 rotation of this face -> shared memory (81 reals, spread over the face's threads)
    const R* tg = static_cast<const R*>(p.tinvT) + 81 * face;
    for (int i = j; i < 81; i += TPF) {
      tmat[i] = tg[i];
    }
  }
  int totalRows = 0;
#pragma unroll
  for (int d = 0; d < O; ++d) {
    totalRows += G::derRows(d);
  }

  if (tid < NF) {  // j == 0: this thread moves both derivative families of face f
    if (face >= 0) {
      mbar_arrive_expect_tx(bar, p.derSize * sizeof(R));
      tma_load(raw0, p.derPlus.at(face), p.derSize * sizeof(R), bar);
      mbar_arrive_expect_tx(bar + 1, p.derSize * sizeof(R));
      tma_load(raw1, p.derMinus.at(face), p.derSize * sizeof(R), bar + 1);
    } else {
      mbar_arrive_expect_tx(bar, 0);
      mbar_arrive_expect_tx(bar + 1, 0);
    }
  }
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    R* const raw = side == 0 ? raw0 : raw1;
    mbar_wait(bar + side, 0);
    __syncthreads();  // (first pass) the face's rotation matrix written above is visible

    // ---- rotate every row of every dQ(d) in place: row <- row . TinvT
    if (active && face >= 0) {
      for (int r = j; r < totalRows; r += TPF) {
        int d = 0, k = r;
        while (k >= G::derRows(d)) {
          k -= G::derRows(d);
          ++d;
        }
        R* row = raw + p.derOffset[d] + k;
        const int ld = p.derLd[d];
        R in[9], out[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          in[c] = row[c * ld];
        }
#pragma unroll
        for (int b = 0; b < 9; ++b) {
          R acc = in[0] * tmat[9 * b];
#pragma unroll
          for (int a = 1; a < 9; ++a) {
            acc = fma(in[a], tmat[a + 9 * b], acc);
          }
          out[b] = acc;
        }
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          row[c * ld] = out[c];
        }
      }
    }
    __syncthreads();

    // ---- Taylor sums of column q at the V time points of group sg (registers)
    PT x[BP];
#pragma unroll
    for (int k = 0; k < BP; ++k) {
      x[k] = pk_zero_of<PT>();
    }
#pragma unroll
    for (int d = 0; d < O; ++d) {
      R cv[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int s = sg * V + v;
        cv[v] = s < O ? static_cast<R>(p.coeffs[s][d]) : R(0);
      }
      const PT cp = pk_make(cv);
      const R* col = raw + p.derOffset[d] + q * p.derLd[d];
#pragma unroll
      for (int k4 = 0; k4 < (G::derRows(d) + 3) / 4; ++k4) {
        R t[4];
        ld4(t, col + 4 * k4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (4 * k4 + i < G::derRows(d)) {
            x[4 * k4 + i] = pk_fma(cp, pk_bcast<PT>(t[i]), x[4 * k4 + i]);
          }
        }
      }
    }

    // ---- project to the fault quadrature points: out = V3mTo2n(side, rel) . x
    PT out[PP];
    const int variant = side == 0 ? info[0] : info[1] + 4 * info[2];
    GD::template V3mTo2n<PT>(variant, out, SrcReg<PT, BP>{x});

    if (active && face >= 0) {
      R* base = static_cast<R*>(side == 0 ? p.qPlus : p.qMinus);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int s = sg * V + v;
        if (s < O) {
          R* dst = base + ((face * O + s) * 9 + q) * static_cast<long long>(p.ppad);
#pragma unroll
          for (int p4 = 0; p4 < PP / 4; ++p4) {
            st4lane(dst + 4 * p4, v, out[4 * p4], out[4 * p4 + 1], out[4 * p4 + 2], out[4 * p4 + 3]);
          }
        }
      }
    }
    __syncthreads();  // the family buffer is overwritten by the other side's bulk copy
  }
}

// ---------------------------------------------------------------------------------------
// Friction.  One CTA per fault face, thread = quadrature point.
template <typename R> __device__ __forceinline__ R rsAlmostZero();
template <> __device__ __forceinline__ float rsAlmostZero<float>() { return 1e-35f; }
template <> __device__ __forceinline__ double rsAlmostZero<double>() { return 1e-45; }

// rs::arsinhexp / arsinhexpDerivative / computeCExp / logsinh (FrictionLaws/RateAndStateCommon.h:48-137)
template <typename R>
__device__ __forceinline__ bool rsBigBranch(R x, R expLog) {
  int xexp;
  (void)frexp(x, &xexp);
  return expLog + R(max(xexp, 0)) * R(0.69314718055994530943) > R(10) || expLog >= R(50);
}
// sqrt(x^2 + e^2) of the big-argument branches.  Double precision: the reference's expression.  Single precision:
// the same expression on operands scaled by 2^64 when both are tiny -- DELIBERATE DEVIATION from the reference's float arithmetic: with expLog = state / a ~ 60 the operands are
// e ~ 1e-26 and x = cLin * slipRate, whose squares underflow in float as soon as a Newton iterate is clamped to
// almostZero (1e-35); the reference's expression then returns log(x) instead of log(e): a NEGATIVE friction
// coefficient and an infinite derivative, the iterate never leaves the clamp again and the point stays locked for the
// rest of the run (found by the N-step parity run against the double-precision oracle, tests/test_gpu_dr.py).
// Where nothing underflows the two agree to an ulp.
__device__ __forceinline__ double rsHypot(double x, double e) { return sqrt(x * x + e * e); }
__device__ __forceinline__ float rsHypot(float x, float e) {
  const bool tiny = fabsf(x) < 1e-18f && fabsf(e) < 1e-18f;
  const float up = tiny ? 18446744073709551616.f : 1.f, down = tiny ? 5.42101086242752217e-20f : 1.f;
  const float xs = x * up, es = e * up;
  return sqrtf(xs * xs + es * es) * down;
}
template <typename R>
__device__ __forceinline__ R rsArsinhexp(R x, R expLog, R e) {
  if (rsBigBranch(x, expLog)) {
    if (expLog <= R(0)) {
      e = R(1) / e;
    }
    const R xa = fabs(x);
    const R xs = x >= R(0) ? R(1) : R(-1);
    return xs * (expLog + log(xa + rsHypot(xa, e)));
  }
  if (expLog > R(0)) {
    e = R(1) / e;
  }
  return asinh(e * x);
}
template <typename R>
__device__ __forceinline__ R rsArsinhexpDerivative(R x, R expLog, R e) {
  if (rsBigBranch(x, expLog)) {
    if (expLog <= R(0)) {
      e = R(1) / e;
    }
    return R(1) / rsHypot(x, e);
  }
  if (expLog > R(0)) {
    e = R(1) / e;
  }
  const R v = e * x;
  return e / sqrt(R(1) + v * v);
}
template <typename R>
__device__ __forceinline__ R rsLogsinh(R x, R c) {
  const R sign = c >= R(0) ? R(1) : R(-1);
  const R absC = fabs(c);
  return absC + log(x / R(2) * -sign * expm1(R(-2) * absC));
}

template <typename R>
struct FaultPoint {  // per-point view of the per-face arrays
  long long idx;     // face * ppad + point
  template <typename T> __device__ __forceinline__ T& at(void* a) const { return static_cast<T*>(a)[idx]; }
  template <typename T> __device__ __forceinline__ T get(const void* a) const { return static_cast<const T*>(a)[idx]; }
};

// resampled[p] = sum_j resample[p][j] x[j]  (dynamicRupture::kernel::resampleParameter); block-wide
template <typename R, int P>
__device__ __forceinline__ R resamplePoint(const R* __restrict__ resampleT, R* xs, R x, int pt, bool active) {
  __syncthreads();
  if (active) {
    xs[pt] = x;
  }
  __syncthreads();
  R acc = R(0);
  if (active) {
#pragma unroll 7
    for (int jj = 0; jj < P; ++jj) {
      acc = fma(resampleT[jj * P + pt], xs[jj], acc);
    }
  }
  return acc;
}

enum class Law { LinearSlipWeakening, RateAndStateFVW };

template <typename R, int O, int P, Law LAW>
__global__ void __launch_bounds__((P + 31) / 32 * 32) k_friction(const FrictionParams p) {
  __shared__ R xs[(P + 31) / 32 * 32];
  const int pt = threadIdx.x;
  const bool active = pt < P;
  const long long face = blockIdx.x;
  const FaultPoint<R> fp{face * p.ppad + (active ? pt : 0)};
  const R* imp = static_cast<const R*>(p.impAndEta) + 11 * face;
  const R etaP = imp[4] * static_cast<R>(p.etaPDamp), etaS = imp[5], invEtaS = imp[6];
  const R invZp = imp[7], invZs = imp[8], invZpNeig = imp[9], invZsNeig = imp[10];
  const R* resT = static_cast<const R*>(p.resampleT);

  // ---- precomputeStressFromQInterpolated (FrictionSolverCommon.h:160-189); the interpolated fields stay
  //      in registers for the post-compute
  R qpN[O], qpT1[O], qpT2[O], qpU[O], qpV[O], qpW[O], qmN[O], qmT1[O], qmT2[O], qmU[O], qmV[O], qmW[O];
  R ns[O], t1[O], t2[O], tr1[O], tr2[O];
  {
    const R* qp = static_cast<const R*>(p.qPlus) + face * O * 9 * p.ppad + (active ? pt : 0);
    const R* qm = static_cast<const R*>(p.qMinus) + face * O * 9 * p.ppad + (active ? pt : 0);
#pragma unroll
    for (int o = 0; o < O; ++o) {
      const long long b = static_cast<long long>(o) * 9 * p.ppad;
      qpN[o] = qp[b]; qpT1[o] = qp[b + 3 * p.ppad]; qpT2[o] = qp[b + 5 * p.ppad];
      qpU[o] = qp[b + 6 * p.ppad]; qpV[o] = qp[b + 7 * p.ppad]; qpW[o] = qp[b + 8 * p.ppad];
      qmN[o] = qm[b]; qmT1[o] = qm[b + 3 * p.ppad]; qmT2[o] = qm[b + 5 * p.ppad];
      qmU[o] = qm[b + 6 * p.ppad]; qmV[o] = qm[b + 7 * p.ppad]; qmW[o] = qm[b + 8 * p.ppad];
      ns[o] = etaP * (qmU[o] - qpU[o] + qpN[o] * invZp + qmN[o] * invZpNeig);
      t1[o] = etaS * (qmV[o] - qpV[o] + qpT1[o] * invZs + qmT1[o] * invZsNeig);
      t2[o] = etaS * (qmW[o] - qpW[o] + qpT2[o] * invZs + qmT2[o] * invZsNeig);
      tr1[o] = R(0);
      tr2[o] = R(0);
    }
  }
  // initial stress in fault coordinates; with nucleation it is ramped up in place, time point by time point
  // (common::adjustInitialStress, FrictionSolverCommon.h:503-531, called from BaseFrictionLaw.h:106-118)
  R* const stress = static_cast<R*>(const_cast<void*>(p.initialStress)) + face * 6 * p.ppad + (active ? pt : 0);
  R s0 = stress[0], s3 = stress[3 * p.ppad], s5 = stress[5 * p.ppad];
  const int nuc = p.nucleationCount;
  R s1 = R(0), s2 = R(0), s4 = R(0);
  if (nuc > 0) {
    s1 = stress[1 * p.ppad]; s2 = stress[2 * p.ppad]; s4 = stress[4 * p.ppad];
  }
  R pressure = p.initialPressure != nullptr ? fp.template get<R>(p.initialPressure) : R(0);
  auto smoothStep = [](R t, R t0) -> R {   // src/Numerical/GaussianNucleationFunction.h:22-32
    if (t <= R(0)) {
      return R(0);
    }
    if (t < t0) {
      const R tau = t - t0;
      return exp(tau * tau / (t * (t - R(2) * t0)));
    }
    return R(1);
  };

  R mu = fp.template get<R>(p.mu);
  R slip1 = fp.template get<R>(p.slip1), slip2 = fp.template get<R>(p.slip2);
  R sr1 = fp.template get<R>(p.slipRate1), sr2 = fp.template get<R>(p.slipRate2);
  R srm = fp.template get<R>(p.slipRateMagnitude);
  R accSlip = fp.template get<R>(p.accumulatedSlipMagnitude);
  R trac1 = fp.template get<R>(p.traction1), trac2 = fp.template get<R>(p.traction2);
  R ruptureTime = fp.template get<R>(p.ruptureTime), peak = fp.template get<R>(p.peakSlipRate);
  R dynStressTime = fp.template get<R>(p.dynStressTime);
  bool rtPending = p.ruptureTimePending[fp.idx] != 0, dsPending = p.dynStressTimePending[fp.idx] != 0;

  // law parameters
  R dC = R(1), muS = R(0), muD = R(0), cohesion = R(0), forcedTime = R(0);
  R rsA = R(1), rsSl0 = R(1), rsF0 = R(0), rsMuW = R(0), rsB = R(0), rsSrW = R(1), stateVar = R(0);
  bool convInner = true, convOuter = true;
  if constexpr (LAW == Law::LinearSlipWeakening) {
    dC = fp.template get<R>(p.dC); muS = fp.template get<R>(p.muS); muD = fp.template get<R>(p.muD);
    cohesion = fp.template get<R>(p.cohesion); forcedTime = fp.template get<R>(p.forcedRuptureTime);
  } else {
    rsA = fp.template get<R>(p.rsA); rsSl0 = fp.template get<R>(p.rsSl0); rsF0 = fp.template get<R>(p.rsF0);
    rsMuW = fp.template get<R>(p.rsMuW); rsB = fp.template get<R>(p.rsB); rsSrW = fp.template get<R>(p.rsSrW);
    stateVar = fp.template get<R>(p.stateVariable);
    convInner = p.convergenceInner[fp.idx] != 0;
    convOuter = p.convergenceOuter[fp.idx] != 0;
  }
  R stateBuf = LAW == Law::RateAndStateFVW ? stateVar : R(0);  // preHook

  const R fullUpdateTime = static_cast<R>(p.fullUpdateTime);
  R updateTime = fullUpdateTime;
#pragma unroll
  for (int o = 0; o < O; ++o) {
    const R dt = static_cast<R>(p.deltaT[o]);
    const R startTime = updateTime;
    updateTime += dt;
    for (int i = 0; i < nuc; ++i) {
      const R nt0 = static_cast<R>(p.nucleationT0[i]), ns0 = static_cast<R>(p.nucleationS0[i]);
      if (updateTime <= nt0 + ns0 && updateTime >= ns0) {
        const R gNuc = smoothStep(updateTime - ns0, nt0) - smoothStep(updateTime - ns0 - dt, nt0);
        const R* nst = static_cast<const R*>(p.nucleationStress) + (face * nuc + i) * 6 * p.ppad + (active ? pt : 0);
        s0 += nst[0] * gNuc; s1 += nst[1 * p.ppad] * gNuc; s2 += nst[2 * p.ppad] * gNuc;
        s3 += nst[3 * p.ppad] * gNuc; s4 += nst[4 * p.ppad] * gNuc; s5 += nst[5 * p.ppad] * gNuc;
        if (p.nucleationPressure != nullptr) {
          pressure += static_cast<const R*>(p.nucleationPressure)[(face * nuc + i) * p.ppad + (active ? pt : 0)] * gNuc;
        }
      }
    }
    const R tt1 = s3 + t1[o];
    const R tt2 = s5 + t2[o];
    if constexpr (LAW == Law::LinearSlipWeakening) {
      // calcStrengthHook, calcSlipRateAndTraction (LinearSlipWeakening.h:141-163, 58-96)
      const R strength = -cohesion - mu * min(s0 + ns[o] + pressure, R(0));   // + initialPressure, :147-156
      const R absT = hypot(tt1, tt2);
      srm = max(R(0), (absT - strength) * invEtaS);
      const R divisor = strength + etaS * srm;
      sr1 = srm * tt1 / divisor;
      sr2 = srm * tt2 / divisor;
      tr1[o] = t1[o] - etaS * sr1;
      tr2[o] = t2[o] - etaS * sr2;
      trac1 = tr1[o];
      trac2 = tr2[o];
      slip1 += sr1 * dt;
      slip2 += sr2 * dt;
      // calcStateVariableHook (:165-199)
      const R resampled = resamplePoint<R, P>(resT, xs, srm, pt, active);
      R time = fullUpdateTime;
#pragma unroll
      for (int i = 0; i <= o; ++i) {
        time += static_cast<R>(p.deltaT[i]);
      }
      accSlip += resampled * dt;
      R state = min(fabs(accSlip) / dC, R(1));
      R f2;
      if (p.t0 == 0.0) {
        f2 = time >= forcedTime ? R(1) : R(0);
      } else {
        f2 = min(max((time - forcedTime) / static_cast<R>(p.t0), R(0)), R(1));
      }
      state = max(state, f2);
      // frictionFunctionHook (:103-116)
      mu = muS - (muS - muD) * state;
      const R heal = static_cast<R>(p.healingThreshold);
      if (peak > heal && srm < heal) {
        mu = muS;
        state = R(0);
      }
      stateBuf = state;
    } else {
      // ---- RateAndStateBase::updateFrictionAndSlip (RateAndState.h:40-79)
      const R sr0 = static_cast<R>(p.rsSr0);
      const R stateRef = stateBuf;
      const R normal = min(R(0), ns[o] + s0 + pressure);  // updateNormalStress (:418-433), no thermal pressurisation
      const R shear = hypot(tt1, tt2);
      srm = max(rsAlmostZero<R>(), hypot(sr1, sr2));
      R localSr = srm;
      // FastVelocityWeakeningLaw::updateStateVariable (:44-77)
      auto updateState = [&](R slipRate) {
        const R low = max(R(0), rsF0 - (rsB - rsA) * log(slipRate / sr0));
        R x8 = slipRate / rsSrW;
        x8 = x8 * x8; x8 = x8 * x8; x8 = x8 * x8;
        const R steadyMu = rsMuW + (low - rsMuW) / pow(R(1) + x8, R(1.0 / 8.0));
        const R steadyState = rsA * rsLogsinh<R>(sr0 / slipRate * R(2), steadyMu / rsA);
        const R pre = -slipRate * (dt / rsSl0);
        return steadyState * (-expm1(pre)) + exp(pre) * stateRef;
      };
      const R cLin = R(0.5) / sr0;
      const R acLin = rsA * cLin;
      bool outerPre = false;
      R state = stateBuf;
      for (int it = 0; it < p.rsNumberStateVariableUpdates; ++it) {
        state = updateState(localSr);
        // invertSlipRateIterative (:340-404)
        const R cExpLog = state / rsA;
        const R cExp = cExpLog > R(0) ? exp(-cExpLog) : exp(cExpLog);
        R test = srm;
        R g = R(0);
        R resolution = R(0);
        bool faceConverged = false;
        for (int i = 0; i < p.rsMaxNumberSlipRateUpdates; ++i) {
          const R muF = rsA * rsArsinhexp<R>(cLin * test, cExpLog, cExp);
          g = -invEtaS * (fabs(normal) * muF - shear) - test;
          // (rsFloatResolutionExit, single precision only: g = invEtaS (shear - |normal| mu) - V is a difference of
          //  stresses; it cannot be resolved below the rounding level of its terms, and |dG| >= 1 carries that level over
          //  to V.  The reference's float build keeps iterating among neighbouring floats until the cap.)
          resolution = R(8) * R(1.1920929e-7) * (invEtaS * (fabs(normal) * muF + shear) + test);
          const bool atResolution = p.rsFloatResolutionExit != 0 && fabs(g) <= resolution;
          const bool ok = !active || fabs(g) < static_cast<R>(p.rsSlipRateTolerance) || atResolution;
          if (__syncthreads_and(ok)) {
            mu = muF;
            faceConverged = true;
            break;
          }
          const R dMuF = acLin * rsArsinhexpDerivative<R>(cLin * test, cExpLog, cExp);
          const R dG = -invEtaS * (fabs(normal) * dMuF) - R(1);
          test = max(rsAlmostZero<R>(), test - g / dG);
        }
        if (!faceConverged) {
          convInner = convInner && (fabs(g) < static_cast<R>(p.rsSlipRateTolerance) ||
                                    (p.rsFloatResolutionExit != 0 && fabs(g) <= resolution));
        }
        localSr = R(0.5) * (srm + test);
        outerPre = fabs(test - srm) < static_cast<R>(p.rsStateTolerance) ||
                   (p.rsFloatResolutionExit != 0 && fabs(test - srm) <= R(2) * resolution);
        srm = test;
        if (__syncthreads_and(!active || outerPre)) {
          break;
        }
      }
      convOuter = convOuter && outerPre;
      // calcSlipRateAndTraction (:216-275)
      state = updateState(localSr);
      {
        const R cExpLog = state / rsA;
        const R cExp = cExpLog > R(0) ? exp(-cExpLog) : exp(cExpLog);
        mu = rsA * rsArsinhexp<R>(cLin * srm, cExpLog, cExp);
      }
      const R strength = -mu * normal;
      const R divisor = strength + etaS * srm;
      sr1 = srm * tt1 / divisor;
      sr2 = srm * tt2 / divisor;
      tr1[o] = t1[o] - etaS * sr1;
      tr2[o] = t2[o] - etaS * sr2;
      trac1 = tr1[o];
      trac2 = tr2[o];
      accSlip += srm * dt;
      slip1 += sr1 * dt;
      slip2 += sr2 * dt;
      stateBuf = state;
    }
    // saveRuptureFrontOutput, saveDynamicStressOutput, savePeakSlipRateOutput
    if (rtPending && srm > R(0.001)) {
      ruptureTime = startTime;
      rtPending = false;
    }
    if constexpr (LAW == Law::LinearSlipWeakening) {
      if (dsPending && fabs(accSlip) >= dC) {
        dynStressTime = startTime;
        dsPending = false;
      }
    } else {
      if (ruptureTime > R(0) && ruptureTime <= startTime && dsPending &&
          mu <= rsMuW + R(0.05) * (rsF0 - rsMuW)) {
        dynStressTime = startTime;
        dsPending = false;
      }
    }
    peak = max(peak, srm);
  }
  if constexpr (LAW == Law::RateAndStateFVW) {
    // postHook: resampleStateVar (FastVelocityWeakeningLaw.h:140-162)
    const R res = resamplePoint<R, P>(resT, xs, stateBuf - stateVar, pt, active);
    stateVar = stateVar + res;
  }

  if (!active) {
    return;
  }
  // ---- postcomputeImposedStateFromNewStress (FrictionSolverCommon.h:360-407)
  R iN = R(0), iT1 = R(0), iT2 = R(0), pU = R(0), pV = R(0), pW = R(0), mU = R(0), mV = R(0), mW = R(0);
#pragma unroll
  for (int o = 0; o < O; ++o) {
    const R w = static_cast<R>(p.timeWeights[o]);
    iN += w * ns[o];
    iT1 += w * tr1[o];
    iT2 += w * tr2[o];
    mU += w * (qmU[o] - invZpNeig * (ns[o] - qmN[o]));
    mV += w * (qmV[o] - invZsNeig * (tr1[o] - qmT1[o]));
    mW += w * (qmW[o] - invZsNeig * (tr2[o] - qmT2[o]));
    pU += w * (qpU[o] + invZp * (ns[o] - qpN[o]));
    pV += w * (qpV[o] + invZs * (tr1[o] - qpT1[o]));
    pW += w * (qpW[o] + invZs * (tr2[o] - qpT2[o]));
  }
  R* ip = static_cast<R*>(p.imposedPlus) + face * 9 * p.ppad + pt;
  R* im = static_cast<R*>(p.imposedMinus) + face * 9 * p.ppad + pt;
  const R valsP[9] = {iN, R(0), R(0), iT1, R(0), iT2, pU, pV, pW};
  const R valsM[9] = {iN, R(0), R(0), iT1, R(0), iT2, mU, mV, mW};
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    ip[c * p.ppad] = valsP[c];
    im[c * p.ppad] = valsM[c];
  }
  // ---- common::computeFrictionEnergy (FrictionSolverCommon.h:632-722), from the interpolated fields still in registers
  if (p.frictionalEnergy != nullptr) {
    const R bPlus = etaS * invZs, bMinus = etaS * invZsNeig;
    const R areaN = -static_cast<const R*>(p.doubledSurfaceArea)[face];
    const R spaceW = static_cast<const R*>(p.spaceWeights)[pt];
    R* const eSlip = static_cast<R*>(p.energySlip) + face * 3 * p.ppad + pt;
    R eAcc = fp.template get<R>(p.energyAccumulatedSlip);
    R eFric = fp.template get<R>(p.frictionalEnergy);
    R e1 = eSlip[0], e2 = eSlip[p.ppad], e3 = eSlip[2 * p.ppad];
#pragma unroll
    for (int o = 0; o < O; ++o) {
      const R w = static_cast<R>(p.timeWeights[o]);
      const R d1 = qmU[o] - qpU[o], d2 = qmV[o] - qpV[o], d3 = qmW[o] - qpW[o];
      eAcc += w * (p.energiesFromAcrossFaultVelocities ? sqrt(d1 * d1 + d2 * d2 + d3 * d3) : srm);
      e1 += w * d1; e2 += w * d2; e3 += w * d3;
      const R t12 = bPlus * qmT1[o] + bMinus * qpT1[o];
      const R t13 = bPlus * qmT2[o] + bMinus * qpT2[o];
      eFric += (w * spaceW * areaN) * (t12 * d2 + t13 * d3);
    }
    eSlip[0] = e1; eSlip[p.ppad] = e2; eSlip[2 * p.ppad] = e3;
    fp.template at<R>(p.energyAccumulatedSlip) = eAcc;
    fp.template at<R>(p.frictionalEnergy) = eFric;
  }
  if (nuc > 0) {
    stress[0] = s0; stress[1 * p.ppad] = s1; stress[2 * p.ppad] = s2;
    stress[3 * p.ppad] = s3; stress[4 * p.ppad] = s4; stress[5 * p.ppad] = s5;
    if (p.initialPressure != nullptr) {
      fp.template at<R>(p.initialPressure) = pressure;
    }
  }
  fp.template at<R>(p.mu) = mu;
  fp.template at<R>(p.slip1) = slip1; fp.template at<R>(p.slip2) = slip2;
  fp.template at<R>(p.slipRate1) = sr1; fp.template at<R>(p.slipRate2) = sr2;
  fp.template at<R>(p.slipRateMagnitude) = srm;
  fp.template at<R>(p.accumulatedSlipMagnitude) = accSlip;
  fp.template at<R>(p.traction1) = trac1; fp.template at<R>(p.traction2) = trac2;
  fp.template at<R>(p.ruptureTime) = ruptureTime; fp.template at<R>(p.peakSlipRate) = peak;
  fp.template at<R>(p.dynStressTime) = dynStressTime;
  p.ruptureTimePending[fp.idx] = rtPending ? 1 : 0;
  p.dynStressTimePending[fp.idx] = dsPending ? 1 : 0;
  if constexpr (LAW == Law::RateAndStateFVW) {
    fp.template at<R>(p.stateVariable) = stateVar;
    p.convergenceInner[fp.idx] = convInner ? 1 : 0;
    p.convergenceOuter[fp.idx] = convOuter ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------
// Nodal flux of dynamic-rupture faces: Q += V3mTo2nTWDivM(side, rel) . (godunov . fluxSolver).
//
// V3mTo2nTWDivM is the one DENSE operator of the path (B x P, 56 x 49 at order 6, 16 (side, relation) variants),
// and the variant differs from element to element, so immediate-operand code would diverge 16-fold inside a
// warp.  Instead: ONE WARP PER ELEMENT, the operator read from a dense table in global memory
// ([variant][point][64], basis index fastest: coalesced, L1/L2 resident, 200 KB in single precision), lane l
// owns the rows l and l + 32 of all nine quantity columns (18 accumulators).  Per DR face:
//   1. T[p][q] = sum_c godunov[c][p] fluxSolver[c][q]   (lane = point, coalesced reads of the imposed state)
//   2. acc[k][q] += sum_p W[p][k] T[p][q]                (T broadcast from shared memory, W coalesced)
// and finally Q[k][q] += acc[k][q] with coalesced read-modify-writes.
template <typename R, typename G, typename GD, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_neigh_dr(const NeighDrParams p) {
  constexpr int B = G::kB, P = GD::kP;
  constexpr int TS = 12;  // reals per point of T in shared memory (9 padded to 12: three 128-bit words)
  __shared__ __align__(16) R sT[WARPS][2 * 32 * TS];
  __shared__ __align__(16) R sFs[WARPS][84];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long el = static_cast<long long>(blockIdx.x) * WARPS + warp;
  if (el >= p.n) {
    return;
  }
  const R* const table = GD::template nodalFluxTable<R>();
  const uchar4 ft = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * el);
  R acc[2][9];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      acc[r][q] = R(0);
    }
  }
  bool any = false;
#pragma unroll 1
  for (int f = 0; f < 4; ++f) {
    const uint8_t ty = f == 0 ? ft.x : (f == 1 ? ft.y : (f == 2 ? ft.z : ft.w));
    const R* god = static_cast<const R*>(p.godunov[4 * el + f]);
    if (ty != 3 || god == nullptr) {
      continue;  // warp-uniform
    }
    any = true;
    const R* fs = static_cast<const R*>(p.fluxSolver[4 * el + f]);
    const int variant = p.sideRel[8 * el + 2 * f] + 4 * p.sideRel[8 * el + 2 * f + 1];
    __syncwarp();
    for (int i = lane; i < 81; i += 32) {
      sFs[warp][i] = fs[i];
    }
    __syncwarp();
    // ---- step 1: T = godunov . fluxSolver for the points lane, lane + 32
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int pt = lane + 32 * half;
      R t[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        t[q] = R(0);
      }
      if (pt < P) {
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          const R gv = god[c * p.ppad + pt];
#pragma unroll
          for (int q = 0; q < 9; ++q) {
            t[q] = fma(gv, sFs[warp][c + 9 * q], t[q]);
          }
        }
      }
      R* dst = &sT[warp][pt * TS];
      st4(dst, t[0], t[1], t[2], t[3]);
      st4(dst + 4, t[4], t[5], t[6], t[7]);
      st4(dst + 8, t[8], R(0), R(0), R(0));
    }
    __syncwarp();
    // ---- step 2: acc += W(variant) . T
    const R* w = table + static_cast<size_t>(variant) * P * 64 + lane;
#pragma unroll 7
    for (int pt = 0; pt < P; ++pt) {
      const R w0 = w[pt * 64];
      const R w1 = w[pt * 64 + 32];
      R t[12];
      {
        R u[4];
        ld4(u, &sT[warp][pt * TS]);     t[0] = u[0]; t[1] = u[1]; t[2] = u[2]; t[3] = u[3];
        ld4(u, &sT[warp][pt * TS + 4]); t[4] = u[0]; t[5] = u[1]; t[6] = u[2]; t[7] = u[3];
        ld4(u, &sT[warp][pt * TS + 8]); t[8] = u[0];
      }
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        acc[0][q] = fma(w0, t[q], acc[0][q]);
        acc[1][q] = fma(w1, t[q], acc[1][q]);
      }
    }
  }
  if (!any) {
    return;
  }
  R* const dofs = static_cast<R*>(const_cast<void*>(p.dofs.at(el)));
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int k = lane + 32 * r;
    if (k < B) {
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        dofs[q * p.ldg + k] += acc[r][q];
      }
    }
  }
}

}  // namespace ssb
