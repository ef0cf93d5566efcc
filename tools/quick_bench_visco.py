#!/usr/bin/env python3
"""Quick per-kernel timing (CUDA events) of the viscoelastic2 proxy kernels (order 5, 3 mechanisms); development aid."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from seissol_b200 import api, proxy  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--real", type=int, default=4)
    ap.add_argument("--cells", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--kernels", default="ader,localwoader,neigh,all")
    ap.add_argument("--cpu", action="store_true", help="also time the reference's CPU kernels on the host cores")
    a = ap.parse_args()
    ctx = api.Context(5, a.real, equation=1, mechanisms=3)
    d = proxy.ViscoProxyData.synthetic(ctx, a.cells)
    d.sort_by_face_relation()
    k = proxy.ViscoProxyKernels(ctx, d)
    print(f"viscoelastic2 order 5 real {a.real} cells {a.cells} resident {d.bytes_resident() / 2**30:.2f} GiB")
    for name in a.kernels.split(","):
        for _ in range(2):
            k.run(name)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
        ev[0].record()
        for i in range(a.steps):
            k.run(name)
            ev[i + 1].record()
        torch.cuda.synchronize()
        t = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)])) * 1e-3
        by = proxy.visco_algorithmic_bytes_per_element(a.real, name) * a.cells
        fl = proxy.visco_nz_flops_per_element(name) * a.cells
        print(f"{name:12s} {t * 1e3:9.3f} ms  {a.cells / t / 1e6:8.1f} Melem/s  {by / t / 1e9:8.1f} GB/s(alg)  "
              f"{fl / t / 1e12:6.2f} NZ-TFLOP/s  finite={bool(torch.isfinite(d.dofs).all() and torch.isfinite(d.dofs_ane).all())}")


def cpu_reference(real):
    """The reference's generated viscoelastic2 CPU kernels (oracle/_ref) on the host cores: proxy `all` on a bounded sample
    (reported next to the GPU numbers; test infrastructure used as the baseline only)."""
    from oracle import aderdg as O
    from oracle import ref_lib
    from oracle import visco as V
    cfg = "visco_o5_f32" if real == 4 else "visco_o5_f64"
    if not ref_lib.available(cfg):
        print("cpu reference: oracle/_ref not built")
        return
    R = ref_lib.ViscoRefLib(cfg)
    dt = np.float32 if real == 4 else np.float64
    d = V.proxy_data(5, 60000, dt, seed=2)
    coeffs = O.time_integrate_coeffs(5, 0.0, 1.0, dt)
    R.proxy_all(d, coeffs, 1)
    t, _, _ = R.proxy_all(d, coeffs, 2)
    print(f"cpu reference (generated kernels, {R.threads()} host threads): 60000 elements x 2 passes of `all` in {t:.2f} s = "
          f"{60000 * 2 / t / 1e6:.4f} Melem/s")


if __name__ == "__main__":
    main()
    if "--cpu" in sys.argv:
        cpu_reference(8 if "--real" in sys.argv and sys.argv[sys.argv.index("--real") + 1] == "8" else 4)
