"""CPU tests of the friction oracle (oracle/friction.py): the rate-and-state helper functions are pinned
against the reference header compiled verbatim into oracle/_ref; pre-/post-compute against the closed
forms the reference's own unit tests use (tests/DynamicRupture/FrictionLaws/FrictionSolverCommon.t.h
style: Uphoff thesis eq. 4.53/4.60); the composed laws are checked for physical invariants."""
import ctypes

import numpy as np
import pytest

from oracle import aderdg as O
from oracle import friction as F
from oracle import ref_lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("cfg,dt", [("elastic_o6_f32", np.float32), ("elastic_o6_f64", np.float64)])
def test_rs_helpers_match_reference_header(cfg, dt):
    if not ref_lib.available(cfg):
        pytest.skip("oracle/_ref not built here")
    lib = ref_lib.RefLib(cfg).lib
    lib.ssref_rs_almost_zero.restype = ctypes.c_double
    assert float(F.almost_zero(dt)) == pytest.approx(lib.ssref_rs_almost_zero(), rel=1e-6)
    rng = np.random.default_rng(0)
    n = 4000
    x = (10.0 ** rng.uniform(-30, 6, n) * rng.choice([-1, 1], n)).astype(dt)
    c = rng.uniform(-80, 80, n).astype(dt)
    e = F.compute_c_exp(c)
    e_ref = np.zeros_like(c)
    lib.ssref_rs_cexp(ctypes.c_int64(n), _p(c), _p(e_ref))
    tol = 20 * np.finfo(dt).eps
    assert np.allclose(e, e_ref, rtol=tol, atol=0)
    out, dout = np.zeros_like(x), np.zeros_like(x)
    lib.ssref_rs_arsinhexp(ctypes.c_int64(n), _p(x), _p(c), _p(e_ref), _p(out), _p(dout))
    # the large-exponent branch computes c + log(...): rounding is relative to |c|, not to the result
    assert np.all(np.abs(F.arsinhexp(x, c, e_ref) - out) <= tol * np.maximum(np.abs(out), np.abs(c)) + np.finfo(dt).tiny)
    assert np.allclose(F.arsinhexp_derivative(x, c, e_ref), dout, rtol=tol, atol=np.finfo(dt).tiny)
    xs = (10.0 ** rng.uniform(-6, 12, n)).astype(dt)
    cs = rng.uniform(0.5, 90, n).astype(dt)
    ls = np.zeros_like(xs)
    lib.ssref_rs_logsinh(ctypes.c_int64(n), _p(xs), _p(cs), _p(ls))
    assert np.allclose(F.logsinh(xs, cs), ls, rtol=tol, atol=tol)
    # identity the helpers exist for: asinh(x e^c) where it does not overflow
    m = np.abs(c) < 20
    direct = np.arcsinh(x[m].astype(np.float64) * np.exp(c[m].astype(np.float64)))
    big = np.abs(direct) > 1e-3   # tiny results underflow in single precision
    assert np.all(np.abs(out[m][big] - direct[big]) <= 50 * np.finfo(dt).eps * np.maximum(np.abs(direct[big]), np.abs(c[m][big])))


def test_resample_matrix_orientation():
    cfg = "elastic_o6_f64"
    if not ref_lib.available(cfg):
        pytest.skip("oracle/_ref not built here")
    R = ref_lib.RefLib(cfg)
    tab = O.load_tables(6, np.float64)
    x = np.random.default_rng(1).uniform(-1, 1, (5, 49))
    assert np.allclose(R.resample(x), x @ tab["resample"].T, rtol=1e-13, atol=1e-14)
    # resampling reproduces constants (partition of unity of the nodal basis)
    assert np.allclose(tab["resample"] @ np.ones(49), 1.0, atol=1e-12)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_pre_and_postcompute_closed_forms(dt):
    order, nf = 4, 3
    P = (order + 1) ** 2
    st, imp, _ = F.fault_state("lsw", nf, order, dt)
    qip, qim = F.fault_qinterpolated(nf, order, dt)
    f = 1
    ns, t1, t2 = F.precompute_stress(imp, f, qip[f], qim[f], 1.0)
    # eq. 4.53: theta = eta * (v- - v+ + sigma+/Z+ + sigma-/Z-), in float64 from the definitions
    q64p, q64m = qip[f].astype(np.float64), qim[f].astype(np.float64)
    zp, zpn = float(imp["zp"][f]), float(imp["zpNeig"][f])
    want = (zp * zpn / (zp + zpn)) * (q64m[..., 6] - q64p[..., 6] + q64p[..., 0] / zp + q64m[..., 0] / zpn)
    assert np.allclose(ns, want, rtol=50 * np.finfo(dt).eps, atol=1e-3 if dt == np.float32 else 1e-9)
    # identical states on both sides and matched impedances: no jump -> theta = sigma, and with the
    # tractions returned unchanged the imposed state equals the time average of the input state
    imp2 = F.impedances(2670.0, np.full(nf, 6000.0), np.full(nf, 3464.0), 2670.0, np.full(nf, 6000.0), np.full(nf, 3464.0), dt)
    ns, t1, t2 = F.precompute_stress(imp2, f, qip[f], qip[f], 1.0)
    assert np.allclose(ns, qip[f][..., 0], rtol=1e-5) and np.allclose(t1, qip[f][..., 3], rtol=1e-5)
    w = np.full(order, 1.0 / order)
    plus, minus = F.postcompute_imposed(imp2, f, ns, t1, t2, qip[f], qip[f], w)
    mean = qip[f].astype(np.float64).mean(axis=0)
    for q in (0, 3, 5, 6, 7, 8):
        assert np.allclose(plus[:, q], mean[:, q], rtol=1e-4, atol=1e-4 * np.abs(mean[:, q]).max())
        assert np.allclose(minus[:, q], mean[:, q], rtol=1e-4, atol=1e-4 * np.abs(mean[:, q]).max())
    assert np.all(plus[:, [1, 2, 4]] == 0)
    assert P == plus.shape[0]


@pytest.mark.parametrize("law_name", ["lsw", "rsfvw"])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_laws_run_and_respect_invariants(law_name, dt):
    order, nf = 4, 6
    tab = O.load_tables(order, dt)
    st, imp, params = F.fault_state(law_name, nf, order, dt)
    qip, qim = F.fault_qinterpolated(nf, order, dt)
    law = (F.LinearSlipWeakening if law_name == "lsw" else F.RateAndStateFastVelocityWeakening)(params, tab["resample"])
    pts, wts = O.gauss_legendre_shifted(order, 0.0, 1e-3)
    dts = F.delta_t(pts)
    assert np.isclose(dts.sum(), 1e-3)
    before = {k: v.copy() for k, v in st.items()}
    plus, minus = F.evaluate(law, st, imp, qip, qim, 0.0, dts, wts / 1e-3)
    for k, v in st.items():
        assert np.all(np.isfinite(v.astype(np.float64))), k
    assert np.all(np.isfinite(plus)) and np.all(np.isfinite(minus))
    # tractions are continuous across the fault, slip rate is non-negative, slip accumulates
    assert np.array_equal(plus[..., [0, 3, 5]], minus[..., [0, 3, 5]])
    assert np.all(st["slipRateMagnitude"] >= 0)
    assert np.all(st["peakSlipRate"] >= st["slipRateMagnitude"] - 1e-30)
    assert np.all(np.abs(st["accumulatedSlipMagnitude"]) >= np.abs(before["accumulatedSlipMagnitude"]) - 1e-12)
    if law_name == "lsw":
        assert np.all((st["mu"] <= st["muS"] + 1e-6) & (st["mu"] >= st["muD"] - 1e-6))
        # the forced-rupture patch is fully weakened
        assert np.allclose(st["mu"][:, : (order + 1) ** 2 // 4], st["muD"][:, : (order + 1) ** 2 // 4])
    else:
        if dt == np.float64:
            # single precision cannot meet the reference's default |g| < 1e-8 on a slipping fault
            # (and x*x + exp*exp underflows at the almostZero slip-rate floor: the reference formula itself)
            assert st["convergenceInner"].all() and st["convergenceOuter"].all()
            assert np.all(st["mu"] > 0) and np.all(st["mu"] < 1.5)


def test_gaussian_nucleation_known_answers():
    """The reference's own test of the nucleation function (tests/Numerical/GaussianNucleation.t.h:18-50): the increment of
    the smooth step against the closed form, for rise times 0.8 and 1.0, dt = 0.01, i = -10..110, both precisions."""
    for dt_ in (np.float32, np.float64):
        step_w = dt_(0.01)
        for rise in (dt_(0.8), dt_(1.0)):
            def closed(t):
                if t < 0:
                    return dt_(0)
                if t > rise:
                    return dt_(1)
                if t == 0:
                    return dt_(0)
                return dt_(np.exp((t - rise) * (t - rise) / (t * (t - 2 * rise))))
            for i in range(-10, 111):
                t = dt_(i) * step_w
                got = F.smooth_step_increment(t, step_w, rise)
                assert abs(got - (closed(t) - closed(t - step_w))) < 1e-4
    # the increments of one rise sum to the full step
    tot = sum(F.smooth_step_increment(np.float64(i * 0.01), np.float64(0.01), np.float64(0.8)) for i in range(0, 200))
    assert abs(tot - 1.0) < 1e-12


def test_adjust_initial_stress_ramps_to_the_nucleation_stress():
    """common::adjustInitialStress (FrictionSolverCommon.h:503-531): over the rise time the initial stress gains exactly
    the nucleation stress, the initial pressure the nucleation pressure (up to the last increment); outside [s0, s0 + t0] nothing changes."""
    rng = np.random.default_rng(0)
    P = 25
    st = {"initialStress": np.zeros((1, 6, P)), "initialPressure": np.zeros((1, P)),
          "nucleationStress": rng.uniform(-1, 1, (1, 2, 6, P)), "nucleationPressure": rng.uniform(-1, 1, (1, 2, P))}
    p = F.Params()
    p.nucleationCount, p.nucleationT0, p.nucleationS0 = 2, [0.4, 0.2], [0.0, 0.5]
    dt = np.float64(0.005)
    t = np.float64(0.0)
    F.adjust_initial_stress(st, 0, p, np.float64(-1.0), dt)
    assert np.all(st["initialStress"] == 0)
    for _ in range(300):
        t = t + dt
        F.adjust_initial_stress(st, 0, p, t, dt)
    # (the reference's window test `time <= t0 + s0` drops the last increment when the accumulated time lands a rounding
    # error above the end of the rise: the sum misses at most that increment)
    assert np.allclose(st["initialStress"][0], st["nucleationStress"][0].sum(0), atol=5e-3)
    assert np.allclose(st["initialPressure"][0], st["nucleationPressure"][0].sum(0), atol=5e-3)
    assert np.abs(st["initialStress"][0]).max() > 0.5


def test_friction_energy_closed_form():
    """common::computeFrictionEnergy (FrictionSolverCommon.h:632-722) on a constant slip-rate field: the slip grows by
    the across-fault velocity x dt, the accumulated slip by its magnitude, the frictional energy by
    -area x space weight x (traction . slip rate) x dt (the closed form of the reference's definition)."""
    dt_ = np.float64
    nf, order, P = 2, 4, 25
    st, imp, params = F.fault_state("lsw", nf, order, dt_)
    qip = np.zeros((nf, order, P, 9))
    qim = np.zeros((nf, order, P, 9))
    qim[..., F.U_], qim[..., F.V_], qim[..., F.W_] = 0.5, 2.0, -1.0     # constant velocity jump
    qip[..., F.T1_], qim[..., F.T1_] = 3e6, 1e6
    qip[..., F.T2_], qim[..., F.T2_] = -2e6, 4e6
    step = 1e-3
    _, w = O.gauss_legendre_shifted(order, 0.0, step)
    sw = np.linspace(0.5, 1.5, P)
    energy = {"slip": np.zeros((nf, 3, P)), "accumulatedSlip": np.zeros((nf, P)), "frictionalEnergy": np.zeros((nf, P)),
              "spaceWeights": sw, "doubledSurfaceArea": np.array([2.0, 3.0]), "fromAcrossFaultVelocities": True}
    for f in range(nf):
        F.compute_friction_energy(energy, f, imp, qip[f], qim[f], w, st["slipRateMagnitude"][f])
    assert np.allclose(energy["slip"][:, 0], 0.5 * step) and np.allclose(energy["slip"][:, 1], 2.0 * step)
    assert np.allclose(energy["accumulatedSlip"], np.sqrt(0.25 + 4 + 1) * step)
    for f in range(nf):
        bp, bm = imp["etaS"][f] * imp["invZs"][f], imp["etaS"][f] * imp["invZsNeig"][f]
        t12, t13 = bp * 1e6 + bm * 3e6, bp * 4e6 + bm * (-2e6)
        want = -energy["doubledSurfaceArea"][f] * sw * (t12 * 2.0 + t13 * (-1.0)) * step
        assert np.allclose(energy["frictionalEnergy"][f], want, rtol=1e-12)
    # the other definition of the accumulated slip: the friction law's slip-rate magnitude
    energy2 = dict(energy, accumulatedSlip=np.zeros((nf, P)), slip=np.zeros((nf, 3, P)), frictionalEnergy=np.zeros((nf, P)),
                   fromAcrossFaultVelocities=False)
    st["slipRateMagnitude"][:] = 0.25
    F.compute_friction_energy(energy2, 0, imp, qip[0], qim[0], w, st["slipRateMagnitude"][0])
    assert np.allclose(energy2["accumulatedSlip"][0], 0.25 * step)
