#!/usr/bin/env python3
"""Golden vectors for the dynamic-rupture part of the path (SURVEY.md §8 a4 DR branch, a5), produced by the
REFERENCE's own generated kernels (oracle/_ref: dynamicRupture::kernel::evaluateAndRotateQAtInterpolationPoints and
nodalFlux behind the wrappers restated in oracle/ref_harness.cpp from src/Kernels/DynamicRupture.cpp:57-108 and
src/Kernels/LinearCK/Neighbor.cpp:90-106).  Run where /root/reference is mounted:
    python tests/golden/make_golden_dr.py
Writes tests/golden/dr_<cfg>.npz: inputs in the oracle's logical layout + the reference outputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import aderdg as O  # noqa: E402
from oracle import build_ref  # noqa: E402
from oracle.ref_lib import RefLib  # noqa: E402

CASES = [("elastic_o4_f32", 4, np.float32), ("elastic_o6_f32", 6, np.float32),
         ("elastic_o6_f64", 6, np.float64)]


def dr_inputs(order, dt, seed, faces=9, cells=7):
    """Seeded inputs shared by this script and tests/test_oracle_dr.py."""
    rng = np.random.default_rng(seed)
    B, P = O.nbasis(order), (order + 1) ** 2
    dp, dm = [], []
    for d in range(order):
        rows = O.nbasis(order - d)
        for lst in (dp, dm):
            dq = np.zeros((faces, B, 9), dtype=dt)
            dq[:, :rows] = rng.uniform(0, 1, (faces, rows, 9))
            lst.append(dq)
    info = np.stack([rng.integers(0, 4, faces), rng.integers(0, 4, faces), rng.integers(0, 4, faces)], 1).astype(np.uint8)
    # every (side, relation) pair of the minus side appears at least once when faces >= 16
    tinv = rng.uniform(-1, 1, (faces, 9, 9)).astype(dt)
    pts, _ = O.gauss_legendre_shifted(order, 0.0, 0.6)
    coeffs = np.stack([O.time_point_coeffs(order, tau, dt) for tau in pts]).astype(dt)
    # nodal flux: cells with a mix of regular and DR faces
    d = O.proxy_data(order, cells, dt, seed=seed + 1)
    ft = np.zeros((cells, 4), dtype=np.uint8)
    ft[rng.uniform(size=(cells, 4)) < 0.5] = 3
    ft[0] = 3
    ft[1] = 0
    m = int((ft == 3).sum())
    index = np.zeros((cells, 4), dtype=np.int64)
    index[ft == 3] = rng.permutation(m)
    dr = {"godunov": rng.uniform(-1, 1, (m, P, 9)).astype(dt),
          "fluxSolver": rng.uniform(-1, 1, (m, 9, 9)).astype(dt),
          "index": index,
          "side": rng.integers(0, 4, (cells, 4)).astype(np.uint8),
          "rel": rng.integers(0, 4, (cells, 4)).astype(np.uint8)}
    I = rng.uniform(-1, 1, (cells, B, 9)).astype(dt)
    return dict(dp=dp, dm=dm, info=info, tinv=tinv, coeffs=coeffs, cell=d, ft=ft, dr=dr, I=I)


def main():
    for cfg, order, dt in CASES:
        build_ref.compile_cfg("/root/reference", cfg)
        R = RefLib(cfg)
        x = dr_inputs(order, dt, 4242 + order, faces=20)
        qp, qm = R.dr_interpolate(x["dp"], x["dm"], x["tinv"], x["info"], x["coeffs"])
        c = x["cell"]
        qn = R.neighbor(c["Q"], x["I"], c["neighbors"], c["Aminus"], c["face_rel"], x["ft"], dr=x["dr"])
        out = {"info": x["info"], "tinv": x["tinv"], "coeffs": x["coeffs"], "qip": qp, "qim": qm,
               "Q": c["Q"], "I": x["I"], "neighbors": c["neighbors"], "Aminus": c["Aminus"], "face_rel": c["face_rel"],
               "ft": x["ft"], "Q_neigh": qn}
        for k, v in x["dr"].items():
            out["dr_" + k] = v
        for d in range(order):
            out[f"dp{d}"] = x["dp"][d]
            out[f"dm{d}"] = x["dm"][d]
        path = os.path.join(ROOT, "tests", "golden", f"dr_{cfg}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
