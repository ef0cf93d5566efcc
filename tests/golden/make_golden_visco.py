#!/usr/bin/env python3
"""Golden vectors of the viscoelastic2 element update (order 5, 3 mechanisms): inputs in the oracle's logical layout
and the outputs of the REFERENCE's own generated CPU kernels (oracle/_ref/libssref_visco_o5_*.so, oracle/build_ref.py).
Run where /root/reference is mounted:  python tests/golden/make_golden_visco.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import aderdg as O  # noqa: E402
from oracle import build_ref  # noqa: E402
from oracle import visco as V  # noqa: E402
from oracle.ref_lib import ViscoRefLib  # noqa: E402

CASES = [("visco_o5_f32", np.float32), ("visco_o5_f64", np.float64)]


def main():
    for cfg, dt in CASES:
        build_ref.compile_cfg("/root/reference", cfg)
        R = ViscoRefLib(cfg)
        n = 10
        d = V.proxy_data(5, n, dt, seed=77)
        d["face_types"][2, 1] = 3   # DR face: no local flux, no neighbor flux
        d["face_types"][6, 3] = 5   # outflow: local flux only
        c = O.time_integrate_coeffs(5, 0.0, 0.5, dt)
        I, Ia = R.ader(d["Q"], d["Qane"], d["star"], d["E"], d["w"], d["W"], c)
        ql, qal = R.local(d["Q"], d["Qane"], I, Ia, d["star"], d["Aplus"], d["E"], d["w"], d["W"], d["face_types"])
        qn, qan = R.neigh(ql, qal, I, d["neighbors"], d["Aminus"], d["w"], d["face_rel"], d["face_types"])
        out = dict(d)
        out.update(coeffs=c, I=I, Iane=Ia, Q_local=ql, Qane_local=qal, Q_all=qn, Qane_all=qan,
                   flops=np.array([R.flops()["derivative"], R.flops()["volumeExt"], sum(R.flops()["localFluxExt"]),
                                   R.flops()["local"], sum(R.flops()["neighborFluxExt"]), R.flops()["neighbor"]]))
        path = os.path.join(ROOT, "tests", "golden", f"{cfg}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
