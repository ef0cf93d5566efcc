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
