"""CPU tests: the numpy oracle against (a) golden vectors produced by the reference's own
generated kernels (tests/golden/*.npz) and (b) oracle/_ref itself when the prebuilt
library is present.  Tolerance: summation-order differences only (100*eps relative L2, the
bound the reference's own generated kernel tests use: yateto KernelTest)."""
import os

import numpy as np
import pytest

from conftest import rel_l2
from oracle import aderdg as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = [("elastic_o4_f32", 4, np.float32), ("elastic_o6_f32", 6, np.float32),
         ("elastic_o6_f64", 6, np.float64)]


def tol(dt):
    return 100 * np.finfo(dt).eps


@pytest.mark.parametrize("cfg,order,dt", CASES)
def test_oracle_matches_golden(cfg, order, dt):
    g = np.load(os.path.join(GOLDEN, f"{cfg}.npz"))
    tab = O.load_tables(order, dt)
    star = O.star_dense(g["star"], tab["star_pattern"])
    I, ders = O.ader(tab, g["Q"], star, g["coeffs"], True)
    assert rel_l2(I, g["I"]) < tol(dt)
    for d in range(order):
        ref = g[f"dQ{d}"]
        assert rel_l2(ders[d], ref) < tol(dt) or np.abs(ref).max() == 0
        # dQ(d) keeps only its first B(O-d) rows
        assert np.all(ders[d][:, O.nbasis(order - d):] == 0)
    assert rel_l2(O.taylor([g[f"dQ{d}"] for d in range(order)], g["coeffs2"]), g["taylor"]) < tol(dt)
    ql = O.local(tab, g["Q"], g["I"], star, O.flux_dense(g["Aplus"]), g["face_types"])
    assert rel_l2(ql, g["Q_local"]) < tol(dt)
    qn = O.neighbor(tab, g["Q_local"], g["I"], g["neighbors"], O.flux_dense(g["Aminus"]),
                    g["face_rel"], g["face_types_neigh"])
    assert rel_l2(qn, g["Q_all"]) < tol(dt)
    # the updates themselves (not just Q + small update) agree
    assert rel_l2(ql - g["Q"], g["Q_local"] - g["Q"]) < 10 * tol(dt)
    assert rel_l2(qn - g["Q_local"], g["Q_all"] - g["Q_local"]) < 10 * tol(dt)


@pytest.mark.parametrize("cfg,order,dt", CASES)
def test_oracle_matches_reference_library(cfg, order, dt):
    from oracle import ref_lib
    if not ref_lib.available(cfg):
        pytest.skip("oracle/_ref not built here")
    R = ref_lib.RefLib(cfg)
    tab = O.load_tables(order, dt)
    d = O.proxy_data(order, 257, dt, seed=7)
    c = O.time_integrate_coeffs(order, 0.0, 1.0, dt)
    star = O.star_dense(d["star"], tab["star_pattern"])
    I = O.ader(tab, d["Q"], star, c)
    assert rel_l2(I, R.ader(d["Q"], d["star"], c)) < tol(dt)
    assert rel_l2(O.local(tab, d["Q"], I, star, O.flux_dense(d["Aplus"]), d["face_types"]),
                  R.local(d["Q"], I, d["star"], d["Aplus"], d["face_types"])) < tol(dt)
    assert rel_l2(O.neighbor(tab, d["Q"], I, d["neighbors"], O.flux_dense(d["Aminus"]), d["face_rel"],
                             d["face_types"]),
                  R.neighbor(d["Q"], I, d["neighbors"], d["Aminus"], d["face_rel"], d["face_types"])) < tol(dt)
    # reference static flop counts used for the GFLOP/s metric
    from seissol_b200 import proxy
    f = R.flops()
    assert f["derivative"][0] == proxy.FLOPS[order]["derivative"][0]
    assert f["volume"][0] == proxy.FLOPS[order]["volume"][0]
    assert int(f["localFlux"][0].sum()) == proxy.FLOPS[order]["localFlux"][0]
    assert abs(4 * f["neighboringFlux"][0].mean() - proxy.FLOPS[order]["neighboringFlux"][0]) < 1e-6


def test_time_basis_known_answers():
    """tests/Numerical/TimeBasis.t.h style known answers: integrate / point coefficients."""
    c = O.time_integrate_coeffs(4, 0.0, 2.0)
    assert np.allclose(c, [2.0, 2.0, 8 / 6, 16 / 24])
    c = O.time_integrate_coeffs(3, 1.0, 3.0)
    assert np.allclose(c, [2.0, (9 - 1) / 2, (27 - 1) / 6])
    assert np.allclose(O.time_point_coeffs(4, 0.5), [1, 0.5, 0.125, 0.125 / 6])
    pts, w = O.gauss_legendre_shifted(3, 0.0, 1.0)
    assert np.isclose(w.sum(), 1.0) and np.isclose((w * pts ** 4).sum(), 1 / 5)
    dts = O.friction_delta_t(pts, 1.0)
    assert np.isclose(dts.sum(), 1.0)


def test_linearity_and_zero_star():
    """Size-independent properties: the update is linear in Q; zero star/flux matrices give I = c0*Q."""
    tab = O.load_tables(5, np.float64)
    d = O.proxy_data(5, 33, np.float64, seed=3)
    c = O.time_integrate_coeffs(5, 0.0, 0.7)
    star = O.star_dense(d["star"], tab["star_pattern"])
    a = O.ader(tab, d["Q"], star, c)
    b = O.ader(tab, 2.5 * d["Q"], star, c)
    assert rel_l2(b, 2.5 * a) < 1e-14
    z = O.ader(tab, d["Q"], np.zeros_like(star), c)
    assert rel_l2(z, c[0] * d["Q"]) < 1e-15


def test_energy_output_quadrature_matches_the_reference_mass_matrix():
    """The quadrature + basis of the energy-output oracle (O.volume_energies) reproduce the reference's modal mass
    matrix M3 (codegen/matrices, tables o*.npz), which is what the CUDA kernel integrates with; and the energies of a
    constant state have their closed form."""
    for order in (4, 6):
        tab = O.load_tables(order, np.float64)
        pts, w = O.tetrahedron_quadrature(order + 1)
        assert abs(w.sum() - 1.0 / 6.0) < 1e-14
        phi = O.dubiner_basis(order, pts)
        mass = np.einsum("m,mk,ml->kl", w, phi, phi)
        assert np.abs(mass - tab["M3"]).max() < 1e-13
    B = 56
    Q = np.zeros((2, B, 9))
    Q[:, 0, 6:9] = [2.0, -1.0, 0.5]       # constant velocity (phi_0 = 1)
    Q[:, 0, 0:3] = 3.0                    # isotropic stress
    vol = np.array([0.25, 0.5])
    mat = np.array([[2.0, 1.0, 3.0], [4.0, 0.0, 2.0]])   # second element acoustic
    e = O.volume_energies(6, Q, vol, mat)
    assert abs(e[0] - 0.25 * 0.5 * 3.0 * 5.25) < 1e-13
    assert abs(e[1] - 0.25 * 0.5 * 27.0 / (3 * 2.0 + 2 * 1.0)) < 1e-13     # 0.5 tr(s)^2 / (3 (3 lambda + 2 mu)) * 3 ... = 0.5 * 27 / 8
    assert abs(e[2] - 0.5 * 0.5 * 2.0 * 5.25) < 1e-13
    assert abs(e[3] - 0.5 * 9.0 / 8.0) < 1e-13
    assert np.allclose(e[4:7], 0.25 * 3.0 * np.array([2.0, -1.0, 0.5]), atol=1e-13)
