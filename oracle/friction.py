"""CPU ORACLE (numpy) for the dynamic-rupture friction solvers — TEST INFRASTRUCTURE ONLY.

Plain restatement, face by face and vectorised over the fault quadrature points, of the
reference's CPU friction path (every function cites the file:line it follows):

  precompute_stress      <- common::precomputeStressFromQInterpolated  FrictionSolverCommon.h:146-189
  postcompute_imposed    <- common::postcomputeImposedStateFromNewStress FrictionSolverCommon.h:333-407
  rupture/peak outputs   <- saveRuptureFrontOutput / savePeakSlipRateOutput  :546-592
  evaluate (driver)      <- BaseFrictionLaw<D>::evaluate               CpuImpl/BaseFrictionLaw.h:49-189
  LinearSlipWeakening    <- LinearSlipWeakeningLaw<NoSpecialization>   CpuImpl/LinearSlipWeakening.h:28-250
  RateAndStateFVW        <- RateAndStateBase<FastVelocityWeakeningLaw> CpuImpl/RateAndState.h:40-433,
                            CpuImpl/FastVelocityWeakeningLaw.h:44-165, RateAndStateCommon.h:22-141
  delta_t                <- FrictionSolver::computeDeltaT              FrictionSolver.cpp:20-37

Parity status: the transcendental helpers (arsinhexp, arsinhexpDerivative, computeCExp,
logsinh) are PINNED against the reference header compiled verbatim into oracle/_ref
(tests/test_oracle_friction.py); pre/post-compute are pinned by the closed-form checks the
reference's own tests use (tests/DynamicRupture/FrictionLaws); the composed LSW / RS-FVW
update has no golden vectors in the reference tree ("parity unpinned" at that granularity,
SURVEY.md §8c) — the restatement follows the CPU control flow statement by statement,
including its per-FACE convergence semantics (all points of a face iterate until all converge).

Layout: per-face arrays are [nFaces, P] (P = (O+1)^2 points, no padding); QInterpolated is
[nFaces, O, P, 9]; imposed states [nFaces, P, 9].  Quantity indices as
DynamicRupture/Misc.h:217-234: N=0, T1=3, T2=5, U=6, V=7, W=8.
Nucleation: adjustInitialStress (FrictionSolverCommon.h:503-531) per time point and the initial pressure in the normal
stress (CpuImpl/LinearSlipWeakening.h:147-156, CpuImpl/RateAndState.h:418-433); smooth_step_increment is pinned by the
closed form of the reference's own test (tests/Numerical/GaussianNucleation.t.h).  Thermal pressurisation is not part of
the path (NoTP).
"""
import math

import numpy as np

N_, T1_, T2_, U_, V_, W_ = 0, 3, 5, 6, 7, 8


def impedances(rho_p, cp_p, cs_p, rho_m, cp_m, cs_m, dtype):
    """ImpedancesAndEta (DynamicRupture/Typedefs.h:23-35) from the materials on both sides
    (Initializer/Model/DynamicRuptureMatrices.cpp:341-359)."""
    zp, zs = np.asarray(rho_p * cp_p, dtype=np.float64), np.asarray(rho_p * cs_p, dtype=np.float64)
    zpn, zsn = np.asarray(rho_m * cp_m, dtype=np.float64), np.asarray(rho_m * cs_m, dtype=np.float64)
    d = {"zp": zp, "zs": zs, "zpNeig": zpn, "zsNeig": zsn,
         "invZp": 1 / zp, "invZs": 1 / zs, "invZpNeig": 1 / zpn, "invZsNeig": 1 / zsn}
    d["etaP"] = 1.0 / (1.0 / zp + 1.0 / zpn)
    d["etaS"] = 1.0 / (1.0 / zs + 1.0 / zsn)
    d["invEtaS"] = 1.0 / d["etaS"]
    return {k: v.astype(dtype) for k, v in d.items()}


IMP_FIELDS = ["zp", "zs", "zpNeig", "zsNeig", "etaP", "etaS", "invEtaS", "invZp", "invZs", "invZpNeig", "invZsNeig"]


class Params:
    """FrictionLawParameters (DynamicRupture/Misc.h:254-287), the members used on this path."""

    def __init__(self, **kw):
        self.healingThreshold = -1.0
        self.rsSr0 = 1e-6
        self.t0 = 0.0
        self.etaDamp = 1.0
        self.etaDampEnd = math.inf
        self.nucleationCount = 0
        self.nucleationT0 = []
        self.nucleationS0 = []
        self.rsMaxNumberSlipRateUpdates = 60
        self.rsNumberStateVariableUpdates = 10
        self.rsSlipRateTolerance = 1e-8
        self.rsStateTolerance = 1e-8
        for k, v in kw.items():
            if not hasattr(self, k):
                raise KeyError(k)
            setattr(self, k, v)


# ----------------------------------------------------------------------------- RateAndStateCommon.h
def smooth_step(t, t0):
    """gaussianNucleationFunction::smoothStep (src/Numerical/GaussianNucleationFunction.h:22-32), scalar of t's type."""
    if t <= 0:
        return type(t)(0)
    if t < t0:
        tau = t - t0
        return type(t)(np.exp(tau * tau / (t * (t - type(t)(2) * t0))))
    return type(t)(1)


def smooth_step_increment(t, dt, t0):
    """smoothStepIncrement (:37-40)."""
    return smooth_step(t, t0) - smooth_step(t - dt, t0)


def adjust_initial_stress(st, f, p, update_time, dt):
    """common::adjustInitialStress (FrictionSolverCommon.h:503-531) for every nucleation of face f, in place on
    st['initialStress'][f] ([6, P]) and st['initialPressure'][f]."""
    for i in range(p.nucleationCount):
        t0, s0 = update_time.__class__(p.nucleationT0[i]), update_time.__class__(p.nucleationS0[i])
        if update_time <= t0 + s0 and update_time >= s0:
            g = smooth_step_increment(update_time - s0, dt, t0)
            st["initialStress"][f] += st["nucleationStress"][f, i] * g
            st["initialPressure"][f] += st["nucleationPressure"][f, i] * g


def almost_zero(dtype):
    return dtype(1e-45) if dtype == np.float64 else dtype(1e-35)


def arsinhexp(x, exp_log, exp):
    """rs::arsinhexp (RateAndStateCommon.h:48-71): asinh(x * e^c) without overflow."""
    dt = x.dtype.type
    _, xexp = np.frexp(x)
    big = (exp_log + np.maximum(xexp, 0).astype(x.dtype) * dt(0.69314718055994530943) > dt(10)) | (exp_log >= dt(50))
    e_big = np.where(exp_log <= 0, dt(1) / exp, exp)
    xa = np.abs(x)
    xs = np.where(x >= 0, dt(1), dt(-1))
    with np.errstate(all="ignore"):
        r_big = xs * (exp_log + np.log(xa + np.sqrt(xa * xa + e_big * e_big)))
        e_small = np.where(exp_log > 0, dt(1) / exp, exp)
        r_small = np.arcsinh(e_small * x)
    return np.where(big, r_big, r_small).astype(x.dtype)


def arsinhexp_derivative(x, exp_log, exp):
    """rs::arsinhexpDerivative (RateAndStateCommon.h:92-113)."""
    dt = x.dtype.type
    _, xexp = np.frexp(x)
    big = (exp_log + np.maximum(xexp, 0).astype(x.dtype) * dt(0.69314718055994530943) > dt(10)) | (exp_log >= dt(50))
    with np.errstate(all="ignore"):
        e_big = np.where(exp_log <= 0, dt(1) / exp, exp)
        r_big = dt(1) / np.sqrt(x * x + e_big * e_big)
        e_small = np.where(exp_log > 0, dt(1) / exp, exp)
        v = e_small * x
        r_small = e_small / np.sqrt(dt(1) + v * v)
    return np.where(big, r_big, r_small).astype(x.dtype)


def compute_c_exp(c_exp_log):
    """rs::computeCExp (RateAndStateCommon.h:77-86)."""
    return np.where(c_exp_log > 0, np.exp(-c_exp_log), np.exp(c_exp_log)).astype(c_exp_log.dtype)


def logsinh(x, c):
    """rs::logsinh (RateAndStateCommon.h:131-137): log(x * sinh(c))."""
    dt = x.dtype.type
    sign = np.where(c >= 0, dt(1), dt(-1))
    abs_c = np.abs(c)
    with np.errstate(all="ignore"):
        return (abs_c + np.log(x / dt(2) * -sign * np.expm1(dt(-2) * abs_c))).astype(x.dtype)


# ----------------------------------------------------------------------------- common pre/post
def precompute_stress(imp, f, qip, qim, eta_p_damp):
    """precomputeStressFromQInterpolated, elastic branch (FrictionSolverCommon.h:160-189).
    qip/qim: [O, P, 9] of face f.  Returns normalStress, traction1, traction2: [O, P]."""
    dt = qip.dtype.type
    eta_p = imp["etaP"][f] * dt(eta_p_damp)
    eta_s = imp["etaS"][f]
    ns = eta_p * (qim[:, :, U_] - qip[:, :, U_] + qip[:, :, N_] * imp["invZp"][f] + qim[:, :, N_] * imp["invZpNeig"][f])
    t1 = eta_s * (qim[:, :, V_] - qip[:, :, V_] + qip[:, :, T1_] * imp["invZs"][f] + qim[:, :, T1_] * imp["invZsNeig"][f])
    t2 = eta_s * (qim[:, :, W_] - qip[:, :, W_] + qip[:, :, T2_] * imp["invZs"][f] + qim[:, :, T2_] * imp["invZsNeig"][f])
    return ns, t1, t2


def postcompute_imposed(imp, f, ns, tr1, tr2, qip, qim, time_weights):
    """postcomputeImposedStateFromNewStress, elastic branch (FrictionSolverCommon.h:360-407).
    ns: fault normal stress [O,P]; tr1/tr2: traction RESULTS [O,P].  Returns (plus, minus) [P,9]."""
    P = qip.shape[1]
    plus = np.zeros((P, 9), dtype=qip.dtype)
    minus = np.zeros((P, 9), dtype=qip.dtype)
    for o in range(qip.shape[0]):
        w = qip.dtype.type(time_weights[o])
        minus[:, N_] += w * ns[o]
        minus[:, T1_] += w * tr1[o]
        minus[:, T2_] += w * tr2[o]
        minus[:, U_] += w * (qim[o, :, U_] - imp["invZpNeig"][f] * (ns[o] - qim[o, :, N_]))
        minus[:, V_] += w * (qim[o, :, V_] - imp["invZsNeig"][f] * (tr1[o] - qim[o, :, T1_]))
        minus[:, W_] += w * (qim[o, :, W_] - imp["invZsNeig"][f] * (tr2[o] - qim[o, :, T2_]))
        plus[:, N_] += w * ns[o]
        plus[:, T1_] += w * tr1[o]
        plus[:, T2_] += w * tr2[o]
        plus[:, U_] += w * (qip[o, :, U_] + imp["invZp"][f] * (ns[o] - qip[o, :, N_]))
        plus[:, V_] += w * (qip[o, :, V_] + imp["invZs"][f] * (tr1[o] - qip[o, :, T1_]))
        plus[:, W_] += w * (qip[o, :, W_] + imp["invZs"][f] * (tr2[o] - qip[o, :, T2_]))
    return plus, minus


def delta_t(time_points, dtype=np.float64):
    """FrictionSolver::computeDeltaT (FrictionSolver.cpp:20-37)."""
    d = np.empty(len(time_points), dtype=np.float64)
    d[0] = time_points[0]
    d[1:] = np.diff(time_points)
    d[-1] += time_points[0]
    return d.astype(dtype)


# ----------------------------------------------------------------------------- friction laws
class LinearSlipWeakening:
    """LinearSlipWeakeningLaw<NoSpecialization> (CpuImpl/LinearSlipWeakening.h)."""
    extra = ["dC", "muS", "muD", "cohesion", "forcedRuptureTime"]

    def __init__(self, params, resample):
        self.p = params
        self.resample = resample

    def pre_hook(self, st, f):
        return np.zeros_like(st["mu"][f])

    def post_hook(self, st, f, state_buf):
        pass

    def update(self, st, imp, f, ns, t1, t2, tr1, tr2, state_buf, time_index, dts, full_update_time):
        dt = st["mu"].dtype.type
        d_t = dt(dts[time_index])
        # calcStrengthHook (:141-163); initialPressure and fluidPressure are zero on this path
        total_normal = st["initialStress"][f, 0] + ns[time_index]
        if "initialPressure" in st:      # + initialPressure (+ fluidPressure = 0 for elastic materials), :147-156
            total_normal = total_normal + st["initialPressure"][f]
        strength = -st["cohesion"][f] - st["mu"][f] * np.minimum(total_normal, dt(0))
        # calcSlipRateAndTraction (:58-96)
        tt1 = st["initialStress"][f, 3] + t1[time_index]
        tt2 = st["initialStress"][f, 5] + t2[time_index]
        absolute = np.hypot(tt1, tt2)  # misc::magnitude == std::hypot
        srm = np.maximum(dt(0), (absolute - strength) * imp["invEtaS"][f])
        st["slipRateMagnitude"][f] = srm
        divisor = strength + imp["etaS"][f] * srm
        st["slipRate1"][f] = srm * tt1 / divisor
        st["slipRate2"][f] = srm * tt2 / divisor
        tr1[time_index] = t1[time_index] - imp["etaS"][f] * st["slipRate1"][f]
        tr2[time_index] = t2[time_index] - imp["etaS"][f] * st["slipRate2"][f]
        st["traction1"][f] = tr1[time_index]
        st["traction2"][f] = tr2[time_index]
        st["slip1"][f] += st["slipRate1"][f] * d_t
        st["slip2"][f] += st["slipRate2"][f] * d_t
        # calcStateVariableHook (:165-199): resampled slip rate (LinearSlipWeakening.cpp:18-26)
        resampled = (self.resample @ srm.astype(self.resample.dtype)).astype(srm.dtype)
        time = dt(full_update_time)
        for i in range(time_index + 1):
            time = dt(time + dt(dts[i]))
        st["accumulatedSlipMagnitude"][f] += resampled * d_t
        state = np.minimum(np.abs(st["accumulatedSlipMagnitude"][f]) / st["dC"][f], dt(1))
        if self.p.t0 == 0:
            f2 = (time >= st["forcedRuptureTime"][f]).astype(srm.dtype)
        else:
            f2 = np.clip((time - st["forcedRuptureTime"][f]) / dt(self.p.t0), dt(0), dt(1))
        state = np.maximum(state, f2)
        # frictionFunctionHook (:103-116)
        st["mu"][f] = st["muS"][f] - (st["muS"][f] - st["muD"][f]) * state
        heal = (st["peakSlipRate"][f] > dt(self.p.healingThreshold)) & (srm < dt(self.p.healingThreshold))
        st["mu"][f] = np.where(heal, st["muS"][f], st["mu"][f])
        state = np.where(heal, dt(0), state)
        state_buf[:] = state

    def save_dynamic_stress(self, st, f, time):
        """saveDynamicStressOutput (:118-129)."""
        hit = st["dynStressTimePending"][f] & (np.abs(st["accumulatedSlipMagnitude"][f]) >= st["dC"][f])
        st["dynStressTime"][f] = np.where(hit, st["mu"].dtype.type(time), st["dynStressTime"][f])
        st["dynStressTimePending"][f] &= ~hit


class RateAndStateFastVelocityWeakening:
    """RateAndStateBase<FastVelocityWeakeningLaw<NoTP>, NoTP> (CpuImpl/RateAndState.h,
    CpuImpl/FastVelocityWeakeningLaw.h)."""
    extra = ["a", "sl0", "stateVariable", "f0", "muW", "b", "srW", "convergenceInner", "convergenceOuter"]

    def __init__(self, params, resample):
        self.p = params
        self.resample = resample

    def pre_hook(self, st, f):
        return st["stateVariable"][f].copy()

    def post_hook(self, st, f, state_buf):
        """resampleStateVar (FastVelocityWeakeningLaw.h:140-162)."""
        delta = state_buf - st["stateVariable"][f]
        st["stateVariable"][f] = st["stateVariable"][f] + (self.resample @ delta.astype(self.resample.dtype)).astype(delta.dtype)

    # FastVelocityWeakeningLaw::updateStateVariable (:44-77)
    def update_state_variable(self, st, f, ref, d_t, slip_rate):
        dt = ref.dtype.type
        mu_w, sr_w, a, sl0 = st["muW"][f], st["srW"][f], st["a"][f], st["sl0"][f]
        sr0 = dt(self.p.rsSr0)
        with np.errstate(all="ignore"):
            low = np.maximum(dt(0), st["f0"][f] - (st["b"][f] - a) * np.log(slip_rate / sr0))
            x8 = (slip_rate / sr_w) ** 2
            x8 = x8 * x8
            x8 = x8 * x8
            steady_mu = mu_w + (low - mu_w) / np.power(dt(1) + x8, dt(1.0 / 8.0))
            steady_state = a * logsinh(sr0 / slip_rate * dt(2), steady_mu / a)
            pre = -slip_rate * (d_t / sl0)
            return (steady_state * (-np.expm1(pre)) + np.exp(pre) * ref).astype(ref.dtype)

    def mu_details(self, st, f, state):
        dt = state.dtype.type
        a = st["a"][f]
        c_lin = dt(0.5) / dt(self.p.rsSr0)
        c_exp_log = state / a
        return {"a": a, "cLin": c_lin, "cExpLog": c_exp_log, "cExp": compute_c_exp(c_exp_log), "acLin": a * c_lin}

    @staticmethod
    def update_mu(sr, d):
        return d["a"] * arsinhexp(d["cLin"] * sr, d["cExpLog"], d["cExp"])

    @staticmethod
    def update_mu_derivative(sr, d):
        return d["acLin"] * arsinhexp_derivative(d["cLin"] * sr, d["cExpLog"], d["cExp"])

    def normal_stress(self, st, f, ns, time_index):
        """updateNormalStress (RateAndState.h:406-419), no fluid pressure."""
        total = ns[time_index] + st["initialStress"][f, 0]
        if "initialPressure" in st:      # RateAndState.h:418-433 (no thermal pressurisation)
            total = total + st["initialPressure"][f]
        return np.minimum(ns.dtype.type(0), total)

    def invert_slip_rate(self, st, imp, f, state, normal, shear):
        """invertSlipRateIterative (RateAndState.h:340-404): Newton on g(V); every point of the
        face iterates until ALL points satisfy |g| < tol (CPU semantics)."""
        dt = state.dtype.type
        details = self.mu_details(st, f, state)
        test = st["slipRateMagnitude"][f].copy()
        tol = dt(self.p.rsSlipRateTolerance)
        g = np.zeros_like(test)
        for _ in range(self.p.rsMaxNumberSlipRateUpdates):
            mu_f = self.update_mu(test, details)
            g = -imp["invEtaS"][f] * (np.abs(normal) * mu_f - shear) - test
            if np.all(np.abs(g) < tol):
                st["mu"][f] = mu_f
                return True, test
            d_mu = self.update_mu_derivative(test, details)
            d_g = -imp["invEtaS"][f] * (np.abs(normal) * d_mu) - dt(1)
            test = np.maximum(almost_zero(type(tol)), test - g / d_g)
        st["convergenceInner"][f] &= np.abs(g) < tol
        return False, test

    def update(self, st, imp, f, ns, t1, t2, tr1, tr2, state_buf, time_index, dts, full_update_time):
        dt = st["mu"].dtype.type
        d_t = dt(dts[time_index])
        # calcInitialVariables (:117-146)
        state_ref = state_buf.copy()
        normal = self.normal_stress(st, f, ns, time_index)
        tt1 = st["initialStress"][f, 3] + t1[time_index]
        tt2 = st["initialStress"][f, 5] + t2[time_index]
        shear = np.hypot(tt1, tt2)
        srm = np.hypot(st["slipRate1"][f], st["slipRate2"][f])
        srm = np.maximum(almost_zero(dt), srm)
        st["slipRateMagnitude"][f] = srm
        local_sr = srm.copy()
        # updateStateVariableIterative (:148-214)
        outer_pre = np.zeros(srm.shape, dtype=bool)
        state = state_buf
        for _ in range(self.p.rsNumberStateVariableUpdates):
            state = self.update_state_variable(st, f, state_ref, d_t, local_sr)
            normal = self.normal_stress(st, f, ns, time_index)
            _, test = self.invert_slip_rate(st, imp, f, state, normal, shear)
            local_sr = dt(0.5) * (st["slipRateMagnitude"][f] + test)
            outer_pre = np.abs(test - st["slipRateMagnitude"][f]) < dt(self.p.rsStateTolerance)
            st["slipRateMagnitude"][f] = test
            if np.all(outer_pre):
                break
        st["convergenceOuter"][f] &= outer_pre
        normal = self.normal_stress(st, f, ns, time_index)
        # calcSlipRateAndTraction (:216-275)
        state = self.update_state_variable(st, f, state_ref, d_t, local_sr)
        details = self.mu_details(st, f, state)
        srm = st["slipRateMagnitude"][f]
        st["mu"][f] = self.update_mu(srm, details)
        strength = -st["mu"][f] * normal
        divisor = strength + imp["etaS"][f] * srm
        st["slipRate1"][f] = srm * tt1 / divisor
        st["slipRate2"][f] = srm * tt2 / divisor
        tr1[time_index] = t1[time_index] - imp["etaS"][f] * st["slipRate1"][f]
        tr2[time_index] = t2[time_index] - imp["etaS"][f] * st["slipRate2"][f]
        st["traction1"][f] = tr1[time_index]
        st["traction2"][f] = tr2[time_index]
        st["accumulatedSlipMagnitude"][f] += srm * d_t
        st["slip1"][f] += st["slipRate1"][f] * d_t
        st["slip2"][f] += st["slipRate2"][f] * d_t
        state_buf[:] = state

    def save_dynamic_stress(self, st, f, time):
        """saveDynamicStressOutput (RateAndState.h:277-292)."""
        dt = st["mu"].dtype.type
        hit = ((st["ruptureTime"][f] > dt(0)) & (st["ruptureTime"][f] <= dt(time)) & st["dynStressTimePending"][f]
               & (st["mu"][f] <= st["muW"][f] + dt(0.05) * (st["f0"][f] - st["muW"][f])))
        st["dynStressTime"][f] = np.where(hit, dt(time), st["dynStressTime"][f])
        st["dynStressTimePending"][f] &= ~hit


COMMON_STATE = ["mu", "slip1", "slip2", "slipRate1", "slipRate2", "slipRateMagnitude", "accumulatedSlipMagnitude",
                "traction1", "traction2", "ruptureTime", "peakSlipRate", "dynStressTime"]
COMMON_FLAGS = ["ruptureTimePending", "dynStressTimePending"]


def compute_friction_energy(energy, f, imp, qip_f, qim_f, time_weights, slip_rate_magnitude_f):
    """common::computeFrictionEnergy (FrictionSolverCommon.h:632-722) for face f.  energy: dict(slip [nF, 3, P],
    accumulatedSlip [nF, P], frictionalEnergy [nF, P], spaceWeights [P], doubledSurfaceArea [nF],
    fromAcrossFaultVelocities bool), accumulated in place."""
    dt = qip_f.dtype.type
    b_plus = imp["etaS"][f] * imp["invZs"][f]
    b_minus = imp["etaS"][f] * imp["invZsNeig"][f]
    area_n = -dt(energy["doubledSurfaceArea"][f])
    for o in range(qip_f.shape[0]):
        w = dt(time_weights[o])
        d1 = qim_f[o, :, U_] - qip_f[o, :, U_]
        d2 = qim_f[o, :, V_] - qip_f[o, :, V_]
        d3 = qim_f[o, :, W_] - qip_f[o, :, W_]
        if energy["fromAcrossFaultVelocities"]:
            energy["accumulatedSlip"][f] += w * np.sqrt(d1 * d1 + d2 * d2 + d3 * d3)
        else:
            energy["accumulatedSlip"][f] += w * slip_rate_magnitude_f
        energy["slip"][f, 0] += w * d1
        energy["slip"][f, 1] += w * d2
        energy["slip"][f, 2] += w * d3
        t12 = b_plus * qim_f[o, :, T1_] + b_minus * qip_f[o, :, T1_]
        t13 = b_plus * qim_f[o, :, T2_] + b_minus * qip_f[o, :, T2_]
        energy["frictionalEnergy"][f] += (w * energy["spaceWeights"].astype(qip_f.dtype) * area_n) * (t12 * d2 + t13 * d3)


def evaluate(law, st, imp, qip, qim, full_update_time, dts, time_weights, energy=None):
    """BaseFrictionLaw<D>::evaluate (CpuImpl/BaseFrictionLaw.h:49-189) over all faces.
    st: dict of state arrays (updated in place); qip/qim: [nF, O, P, 9]; energy: optional friction energy output
    (compute_friction_energy, drParameters.isFrictionEnergyRequired).
    Returns imposedStatePlus, imposedStateMinus: [nF, P, 9]."""
    nf, order, P, _ = qip.shape
    dt = qip.dtype.type
    plus = np.zeros((nf, P, 9), dtype=qip.dtype)
    minus = np.zeros((nf, P, 9), dtype=qip.dtype)
    for f in range(nf):
        eta_damp = law.p.etaDamp if law.p.etaDampEnd > full_update_time else 1.0
        ns, t1, t2 = precompute_stress(imp, f, qip[f], qim[f], eta_damp)
        state_buf = law.pre_hook(st, f)
        tr1 = np.zeros_like(t1)
        tr2 = np.zeros_like(t2)
        update_time = dt(full_update_time)
        for ti in range(order):
            start_time = update_time
            update_time = dt(update_time + dt(dts[ti]))
            adjust_initial_stress(st, f, law.p, update_time, dt(dts[ti]))     # BaseFrictionLaw.h:106-118
            law.update(st, imp, f, ns, t1, t2, tr1, tr2, state_buf, ti, dts, full_update_time)
            # saveRuptureFrontOutput (FrictionSolverCommon.h:546-566)
            hit = st["ruptureTimePending"][f] & (st["slipRateMagnitude"][f] > dt(0.001))
            st["ruptureTime"][f] = np.where(hit, start_time, st["ruptureTime"][f])
            st["ruptureTimePending"][f] &= ~hit
            law.save_dynamic_stress(st, f, start_time)
            # savePeakSlipRateOutput (:575-592)
            st["peakSlipRate"][f] = np.maximum(st["peakSlipRate"][f], st["slipRateMagnitude"][f])
        law.post_hook(st, f, state_buf)
        plus[f], minus[f] = postcompute_imposed(imp, f, ns, tr1, tr2, qip[f], qim[f], time_weights)
        if energy is not None:
            compute_friction_energy(energy, f, imp, qip[f], qim[f], time_weights, st["slipRateMagnitude"][f])
    return plus, minus


# ----------------------------------------------------------------------------- synthetic fault data
def fault_state(law_name, n_faces, order, dtype, seed=7, shear=75e6):
    """Bounded synthetic fault data: TPV5-like LSW, TPV104-like RS-FVW."""
    rng = np.random.default_rng(seed)
    P = (order + 1) ** 2
    z = lambda: np.zeros((n_faces, P), dtype=dtype)  # noqa: E731
    st = {k: z() for k in COMMON_STATE}
    st["ruptureTimePending"] = np.ones((n_faces, P), dtype=bool)
    st["dynStressTimePending"] = np.ones((n_faces, P), dtype=bool)
    stress = np.zeros((n_faces, 6, P), dtype=dtype)
    stress[:, 0] = (-120e6 * (1 + 0.05 * rng.uniform(-1, 1, (n_faces, P)))).astype(dtype)
    stress[:, 3] = (shear * (1 + 0.1 * rng.uniform(-1, 1, (n_faces, P)))).astype(dtype)
    stress[:, 5] = (5e6 * rng.uniform(-1, 1, (n_faces, P))).astype(dtype)
    st["initialStress"] = stress
    rho = 2670.0
    cp = 6000.0 * (1 + 0.05 * rng.uniform(-1, 1, n_faces))
    cs = 3464.0 * (1 + 0.05 * rng.uniform(-1, 1, n_faces))
    cpm = 6000.0 * (1 + 0.05 * rng.uniform(-1, 1, n_faces))
    csm = 3464.0 * (1 + 0.05 * rng.uniform(-1, 1, n_faces))
    imp = impedances(rho, cp, cs, rho, cpm, csm, dtype)
    if law_name == "lsw":
        st["mu"] += dtype(0.677)
        st["muS"] = z() + dtype(0.677)
        st["muD"] = z() + dtype(0.525)
        st["dC"] = (0.4 * (1 + 0.2 * rng.uniform(-1, 1, (n_faces, P)))).astype(dtype)
        st["cohesion"] = z()
        st["forcedRuptureTime"] = z() + dtype(1e9)
        st["forcedRuptureTime"][:, : P // 4] = dtype(0.0)  # a nucleation patch ruptures immediately
        params = Params(healingThreshold=-1.0, t0=0.0)
    else:
        st["a"] = (0.01 * (1 + 0.3 * rng.uniform(0, 1, (n_faces, P)))).astype(dtype)
        st["b"] = z() + dtype(0.014)
        st["sl0"] = z() + dtype(0.4)
        st["f0"] = z() + dtype(0.6)
        st["muW"] = z() + dtype(0.2)
        st["srW"] = z() + dtype(0.1)
        st["mu"] += dtype(0.6)
        st["stateVariable"] = (0.6 * (1 + 0.05 * rng.uniform(-1, 1, (n_faces, P)))).astype(dtype)
        ini = (10.0 ** rng.uniform(-16, -1, (n_faces, P)))
        st["slipRate1"] = ini.astype(dtype)
        st["slipRate2"] = (0.1 * ini).astype(dtype)
        st["convergenceInner"] = np.ones((n_faces, P), dtype=bool)
        st["convergenceOuter"] = np.ones((n_faces, P), dtype=bool)
        params = Params(rsSr0=1e-6)
    return st, imp, params


def fault_qinterpolated(n_faces, order, dtype, seed=11):
    """Synthetic interpolated wave fields at the fault: stresses O(1e6..1e7) Pa, velocities O(0.1..1) m/s."""
    rng = np.random.default_rng(seed)
    P = (order + 1) ** 2
    q = np.zeros((2, n_faces, order, P, 9))
    q[..., :6] = 3e6 * rng.uniform(-1, 1, q[..., :6].shape)
    q[..., 6:] = 0.5 * rng.uniform(-1, 1, q[..., 6:].shape)
    return q[0].astype(dtype), q[1].astype(dtype)
