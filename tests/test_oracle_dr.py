"""CPU tests pinning the dynamic-rupture branch of the numpy oracle:
`oracle.aderdg.dr_interpolate` and `oracle.aderdg.neighbor(dr=...)` against
(a) golden vectors written by the reference's own generated kernels (tests/golden/dr_*.npz,
    tests/golden/make_golden_dr.py) and
(b) oracle/_ref itself when the prebuilt library is present (RefLib.dr_interpolate / RefLib.neighbor(dr=...):
    the reference's dynamicRupture::kernel::evaluateAndRotateQAtInterpolationPoints and nodalFlux).
Tolerance: summation order only, 100*eps relative L2 (the bound of the reference's generated kernel tests)."""
import os
import sys

import numpy as np
import pytest

from conftest import rel_l2
from oracle import aderdg as O
from oracle import ref_lib

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
from make_golden_dr import CASES, dr_inputs  # noqa: E402


def tol(dt):
    return 100 * np.finfo(dt).eps


@pytest.mark.parametrize("cfg,order,dt", CASES)
def test_dr_oracle_matches_golden(cfg, order, dt):
    g = np.load(os.path.join(GOLDEN, f"dr_{cfg}.npz"))
    tab = O.load_tables(order, dt)
    dp = [g[f"dp{d}"] for d in range(order)]
    dm = [g[f"dm{d}"] for d in range(order)]
    qp, qm = O.dr_interpolate(tab, dp, dm, g["tinv"], g["info"], g["coeffs"])
    assert qp.shape == g["qip"].shape == (dp[0].shape[0], order, (order + 1) ** 2, 9)
    assert rel_l2(qp, g["qip"]) < tol(dt)
    assert rel_l2(qm, g["qim"]) < tol(dt)
    # every time point on its own (a wrong coefficient row would hide in the global norm)
    for s in range(order):
        assert rel_l2(qp[:, s], g["qip"][:, s]) < tol(dt)
        assert rel_l2(qm[:, s], g["qim"][:, s]) < tol(dt)
    dr = {k[3:]: g[k] for k in g.files if k.startswith("dr_")}
    qn = O.neighbor(tab, g["Q"], g["I"], g["neighbors"], O.flux_dense(g["Aminus"]), g["face_rel"], g["ft"], dr=dr)
    assert rel_l2(qn, g["Q_neigh"]) < tol(dt)
    assert rel_l2(qn - g["Q"], g["Q_neigh"] - g["Q"]) < tol(dt)
    # the DR contribution is O(1), not hidden behind Q
    no_dr = O.neighbor(tab, g["Q"], g["I"], g["neighbors"], O.flux_dense(g["Aminus"]), g["face_rel"], g["ft"])
    assert rel_l2(no_dr, g["Q_neigh"]) > 0.1


@pytest.mark.parametrize("cfg,order,dt", CASES)
def test_dr_oracle_matches_reference_library(cfg, order, dt):
    from oracle import ref_lib
    if not ref_lib.available(cfg):
        pytest.skip("oracle/_ref not built here")
    R = ref_lib.RefLib(cfg)
    tab = O.load_tables(order, dt)
    x = dr_inputs(order, dt, seed=99, faces=48, cells=33)
    # all 4 plus sides and all 16 (minus side, relation) operator variants are exercised
    x["info"][:16, 1] = np.arange(16) % 4
    x["info"][:16, 2] = np.arange(16) // 4
    x["info"][:4, 0] = np.arange(4)
    qp, qm = O.dr_interpolate(tab, x["dp"], x["dm"], x["tinv"], x["info"], x["coeffs"])
    rp, rm = R.dr_interpolate(x["dp"], x["dm"], x["tinv"], x["info"], x["coeffs"])
    assert rel_l2(qp, rp) < tol(dt)
    assert rel_l2(qm, rm) < tol(dt)
    c = x["cell"]
    got = O.neighbor(tab, c["Q"], x["I"], c["neighbors"], O.flux_dense(c["Aminus"]), c["face_rel"], x["ft"], dr=x["dr"])
    want = R.neighbor(c["Q"], x["I"], c["neighbors"], c["Aminus"], c["face_rel"], x["ft"], dr=x["dr"])
    assert rel_l2(got, want) < tol(dt)
    assert rel_l2(got - c["Q"], want - c["Q"]) < tol(dt)


def nodal_bc_case(rng, n, m, B2, dt):
    """m boundary faces on random (cell, face) slots: Dirichlet (easi map + constant) and FreeSurfaceGravity faces with
    a random rotation Tinv."""
    slots = rng.choice(4 * n, size=m, replace=False)
    kind = np.where(rng.uniform(size=m) < 0.5, 4, 2).astype(np.uint8)
    return {
        "cell": (slots // 4).astype(np.int64), "face": (slots % 4).astype(np.uint8), "kind": kind,
        "tinv": rng.uniform(-1, 1, (m, 9, 9)).astype(dt),
        "easiMap": rng.uniform(-1, 1, (m, 9, 9, B2)).astype(dt),
        "easiConst": rng.uniform(-1, 1, (m, 9, B2)).astype(dt),
        "displacement": rng.uniform(-1, 1, (m, B2)).astype(dt),
        "rhoG": rng.uniform(1.0, 3.0, m),
    }


@pytest.mark.parametrize("cfg,order,dt", [("elastic_o4_f32", 4, np.float32), ("elastic_o6_f32", 6, np.float32),
                                          ("elastic_o6_f64", 6, np.float64)])
def test_nodal_boundary_conditions_match_reference_library(cfg, order, dt):
    """O.local_nodal_bc against the reference's generated projectToNodalBoundaryRotated, createEasiBoundaryGhostCells
    and localFluxNodal kernels (Local.cpp:157-252)."""
    if not ref_lib.available(cfg):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    R = ref_lib.RefLib(cfg)
    tab = O.load_tables(order, dt)
    rng = np.random.default_rng(21)
    n, m = 17, 23
    B2 = order * (order + 1) // 2
    d = O.proxy_data(order, n, dt, seed=9)
    I = rng.uniform(-1, 1, d["Q"].shape).astype(dt)
    bc = nodal_bc_case(rng, n, m, B2, dt)
    want = R.nodal_bc(d["Q"], I, d["Aminus"], bc)
    got = O.local_nodal_bc(tab, d["Q"], I, O.flux_dense(d["Aminus"]), bc)
    tol = 100 * np.finfo(dt).eps
    assert rel_l2(got, want) < tol
    assert rel_l2(got - d["Q"], want - d["Q"]) < 10 * tol
    assert np.linalg.norm(got - d["Q"]) > 1e-3 * np.linalg.norm(d["Q"])


def plane_wave_solution(nodes, t):
    """A smooth stand-in for physics::InitialField::evaluate: nine quantities at the nodes [B2, 3] at time t."""
    k = np.array([0.7, -0.4, 0.9])
    phase = nodes @ k - 1.3 * t
    amp = np.linspace(0.5, 1.5, 9)
    return np.sin(phase)[:, None] * amp[None, :] + 0.1 * t * np.cos(2 * phase)[:, None]


@pytest.mark.parametrize("cfg,order,dt", [("elastic_o4_f32", 4, np.float32), ("elastic_o6_f64", 6, np.float64)])
def test_analytical_boundary_faces_match_reference_library(cfg, order, dt):
    """O.local_nodal_flux (FaceType::Analytical, Local.cpp:233-246) against the reference's generated localFluxNodal: the
    reference library applies it to GIVEN nodal values when the easi map of a Dirichlet face is zero and the easi
    constant holds the values.  The time integration of the callback (DirichletBoundary.h:121-165) is checked against
    the exact integral of the test function."""
    if not ref_lib.available(cfg):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    R = ref_lib.RefLib(cfg)
    tab = O.load_tables(order, dt)
    rng = np.random.default_rng(5)
    n, m = 13, 19
    B2 = order * (order + 1) // 2
    d = O.proxy_data(order, n, dt, seed=4)
    bc = nodal_bc_case(rng, n, m, B2, dt)
    nodes = rng.uniform(-1, 1, (m, B2, 3))
    inodal = O.analytical_inodal(plane_wave_solution, nodes, 0.25, 0.05, order).astype(dt)      # [m, B2, 9]
    bc["kind"][:] = 4
    bc["easiMap"][:] = 0
    bc["easiConst"] = np.ascontiguousarray(inodal.transpose(0, 2, 1))
    want = R.nodal_bc(d["Q"], np.zeros_like(d["Q"]), d["Aminus"], bc)
    got = O.local_nodal_flux(tab, d["Q"], O.flux_dense(d["Aminus"]), bc["cell"], bc["face"], inodal)
    tol = 100 * np.finfo(dt).eps
    assert rel_l2(got, want) < tol
    assert rel_l2(got - d["Q"], want - d["Q"]) < 10 * tol
    # the quadrature of order O integrates the smooth callback to ~ (k dt)^(2 O)
    fine = O.analytical_inodal(plane_wave_solution, nodes[:2], 0.25, 0.05, 12)
    assert rel_l2(O.analytical_inodal(plane_wave_solution, nodes[:2], 0.25, 0.05, order), fine) < 1e-12


def test_analytical_inodal_host_helper_equals_oracle():
    """seissol_b200.boundary.analytical_inodal (product host code) = the oracle's restatement, in the ABI layout."""
    from seissol_b200 import boundary
    rng = np.random.default_rng(8)
    nodes = rng.uniform(-1, 1, (7, 21, 3))
    got = boundary.analytical_inodal(plane_wave_solution, nodes, 1.5, 0.01, 6)
    want = O.analytical_inodal(plane_wave_solution, nodes, 1.5, 0.01, 6)
    assert got.shape == (7, 9, 21)
    assert rel_l2(got.transpose(0, 2, 1), want) < 1e-14
