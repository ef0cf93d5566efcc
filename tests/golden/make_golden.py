#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own
generated CPU kernels (oracle/_ref, see oracle/build_ref.py) on small seeded inputs.
Run in a container where /root/reference is mounted:  python tests/golden/make_golden.py
The .npz files store inputs (oracle logical layout) and the reference outputs."""
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


def main():
    for cfg, order, dt in CASES:
        build_ref.compile_cfg("/root/reference", cfg)
        R = RefLib(cfg)
        n = 12
        d = O.proxy_data(order, n, dt, seed=1234 + order)
        # exercise DR-face skipping in the local kernel on two cells
        d["face_types"][3, 1] = 3
        d["face_types"][7, 0] = 3
        c = O.time_integrate_coeffs(order, 0.0, 0.5, dt)
        I, ders = R.ader(d["Q"], d["star"], c, True)
        c2 = O.time_integrate_coeffs(order, 0.1, 0.35, dt)
        tay = R.taylor(ders, c2)
        ql = R.local(d["Q"], I, d["star"], d["Aplus"], d["face_types"])
        # neighbor stage: face 1 of cell 3 stays DR-less (boundary-like, type 1 = no neighbor flux)
        ftn = np.zeros_like(d["face_types"])
        ftn[3, 1] = 1
        qn = R.neighbor(ql, I, d["neighbors"], d["Aminus"], d["face_rel"], ftn)
        d["face_types_neigh"] = ftn
        out = dict(d)
        out.update(coeffs=c, coeffs2=c2, I=I, taylor=tay, Q_local=ql, Q_all=qn)
        for k, dq in enumerate(ders):
            out[f"dQ{k}"] = dq
        path = os.path.join(ROOT, "tests", "golden", f"{cfg}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
