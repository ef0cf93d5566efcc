"""CPU tests of the viscoelastic2 oracle (oracle/visco.py): against the golden vectors produced by the reference's own
generated kernels (tests/golden/visco_o5_*.npz) and against oracle/_ref itself when the prebuilt library is present.
Tolerance: summation order only (100 eps relative L2, the bound of the reference's generated kernel tests)."""
import os

import numpy as np
import pytest

from conftest import rel_l2
from oracle import aderdg as O
from oracle import ref_lib
from oracle import visco as V

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = [("visco_o5_f32", np.float32), ("visco_o5_f64", np.float64)]


def tol(dt):
    return 100 * np.finfo(dt).eps


@pytest.mark.parametrize("cfg,dt", CASES)
def test_visco_oracle_matches_golden(cfg, dt):
    g = np.load(os.path.join(GOLDEN, f"{cfg}.npz"))
    tab = O.load_tables(5, dt)
    star, E = V.star_dense(g["star"]), V.e_dense(g["E"], 3)
    I, Ia = V.ader(tab, g["Q"], g["Qane"], star, E, g["w"], g["W"], g["coeffs"])
    assert rel_l2(I, g["I"]) < tol(dt) and rel_l2(Ia, g["Iane"]) < tol(dt)
    ql, qal = V.local(tab, g["Q"], g["Qane"], g["I"], g["Iane"], star, V.flux_dense(g["Aplus"]), E, g["w"], g["W"],
                      g["face_types"])
    assert rel_l2(ql, g["Q_local"]) < tol(dt) and rel_l2(qal, g["Qane_local"]) < tol(dt)
    assert rel_l2(ql - g["Q"], g["Q_local"] - g["Q"]) < 10 * tol(dt)
    qn, qan = V.neighbor(tab, g["Q_local"], g["Qane_local"], g["I"], g["neighbors"], V.flux_dense(g["Aminus"]), g["w"],
                         g["face_rel"], g["face_types"])
    assert rel_l2(qn, g["Q_all"]) < tol(dt) and rel_l2(qan, g["Qane_all"]) < tol(dt)
    assert rel_l2(qn - g["Q_local"], g["Q_all"] - g["Q_local"]) < 10 * tol(dt)
    # static non-zero flop counts of the generated kernels
    assert list(g["flops"][:2]) == [120489, 15924] and int(g["flops"][3]) == 14175 and int(g["flops"][5]) == 1575


def test_visco_reduces_to_elastic_without_memory_variables():
    """With E = 0 the elastic columns of the viscoelastic update equal the elastic update with the 24-entry star."""
    dt = np.float64
    tab = O.load_tables(5, dt)
    d = V.proxy_data(5, 7, dt, seed=3)
    d["E"][:] = 0
    coeffs = O.time_integrate_coeffs(5, 0.0, 1.0, dt)
    star = V.star_dense(d["star"])
    I, _ = V.ader(tab, d["Q"], d["Qane"], star, V.e_dense(d["E"], 3), d["w"], d["W"], coeffs)
    I_el = O.ader(tab, d["Q"], star[:, :, :, :9], coeffs)
    assert rel_l2(I, I_el) < 1e-14
    ql, _ = V.local(tab, d["Q"], d["Qane"], I, np.zeros_like(d["Qane"]), star, V.flux_dense(d["Aplus"]),
                    V.e_dense(d["E"], 3), d["w"], d["W"], d["face_types"])
    ql_el = O.local(tab, d["Q"], I_el, star[:, :, :, :9], V.flux_dense(d["Aplus"])[:, :, :, :9], d["face_types"])
    assert rel_l2(ql, ql_el) < 1e-14


@pytest.mark.parametrize("cfg,dt", CASES)
def test_visco_oracle_matches_reference_library(cfg, dt):
    if not ref_lib.available(cfg):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    R = ref_lib.ViscoRefLib(cfg)
    tab = O.load_tables(5, dt)
    d = V.proxy_data(5, 33, dt, seed=5)
    d["face_types"][::4, 2] = 3
    coeffs = O.time_integrate_coeffs(5, 0.0, 1.0, dt)
    star, E = V.star_dense(d["star"]), V.e_dense(d["E"], 3)
    I, Ia = V.ader(tab, d["Q"], d["Qane"], star, E, d["w"], d["W"], coeffs)
    Ir, Iar = R.ader(d["Q"], d["Qane"], d["star"], d["E"], d["w"], d["W"], coeffs)
    assert rel_l2(I, Ir) < tol(dt) and rel_l2(Ia, Iar) < tol(dt)
    ql, qal = V.local(tab, d["Q"], d["Qane"], I, Ia, star, V.flux_dense(d["Aplus"]), E, d["w"], d["W"], d["face_types"])
    qlr, qalr = R.local(d["Q"], d["Qane"], Ir, Iar, d["star"], d["Aplus"], d["E"], d["w"], d["W"], d["face_types"])
    assert rel_l2(ql, qlr) < tol(dt) and rel_l2(qal, qalr) < tol(dt)
    qn, qan = V.neighbor(tab, ql, qal, I, d["neighbors"], V.flux_dense(d["Aminus"]), d["w"], d["face_rel"], d["face_types"])
    qnr, qanr = R.neigh(qlr, qalr, Ir, d["neighbors"], d["Aminus"], d["w"], d["face_rel"], d["face_types"])
    assert rel_l2(qn, qnr) < tol(dt) and rel_l2(qan, qanr) < tol(dt)


def visco_dr_case(rng, n, m, P, dt):
    """m fault faces attached to random (cell, face) slots: imposed states, 9 x 15 flux solvers, side / relation."""
    face_types = np.zeros((n, 4), dtype=np.uint8)
    index = np.zeros((n, 4), dtype=np.int64)
    slots = rng.choice(4 * n, size=m, replace=False)
    face_types.ravel()[slots] = 3
    index.ravel()[slots] = np.arange(m)
    return face_types, {
        "godunov": rng.uniform(-1, 1, (m, P, 9)).astype(dt),
        "fluxSolver": (rng.uniform(-1, 1, (m, 9, 15)) * 1e-2).astype(dt),
        "index": index,
        "side": rng.integers(0, 4, (n, 4)).astype(np.uint8),
        "rel": rng.integers(0, 4, (n, 4)).astype(np.uint8),
    }


@pytest.mark.parametrize("cfg,dt", CASES)
def test_visco_dr_nodal_flux_matches_reference_library(cfg, dt):
    """The DynamicRupture branch of the viscoelastic neighbor integral (LinearCKAnelastic/Neighbor.cpp:107-115): the
    oracle against the reference's generated dynamicRupture::kernel::nodalFlux + kernel::neighbor."""
    if not ref_lib.available(cfg):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    R = ref_lib.ViscoRefLib(cfg)
    tab = O.load_tables(5, dt)
    rng = np.random.default_rng(12)
    n, m, P = 29, 40, 36
    d = V.proxy_data(5, n, dt, seed=6)
    ft, dr = visco_dr_case(rng, n, m, P, dt)
    I = rng.uniform(-1, 1, d["Q"].shape).astype(dt)
    qn, qan = V.neighbor(tab, d["Q"], d["Qane"], I, d["neighbors"], V.flux_dense(d["Aminus"]), d["w"], d["face_rel"], ft, dr=dr)
    qnr, qanr = R.neigh_dr(d["Q"], d["Qane"], I, d["neighbors"], d["Aminus"], d["w"], d["face_rel"], ft, dr)
    assert rel_l2(qn, qnr) < tol(dt) and rel_l2(qan, qanr) < tol(dt)
    q0, qa0 = V.neighbor(tab, d["Q"], d["Qane"], I, d["neighbors"], V.flux_dense(d["Aminus"]), d["w"], d["face_rel"], ft)
    assert rel_l2(qn - q0, qnr - q0) < 10 * tol(dt) and rel_l2(qan - qa0, qanr - qa0) < 10 * tol(dt)   # the DR part itself
    assert np.linalg.norm(qn - q0) > 1e-3 * np.linalg.norm(q0)
