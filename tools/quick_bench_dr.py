#!/usr/bin/env python3
"""Quick timing (CUDA events) of the dynamic-rupture proxy kernels godunov_dr, neigh_dr, all_dr and of the two
friction laws on proxy-shaped data (faces = 4 x cells, src/Proxy/Allocator.cpp:193); development aid."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from seissol_b200 import api, proxy  # noqa: E402


def friction_data(ctx, faces, law, dev, dt, q_plus, q_minus):
    """TPV5-like (LSW) / TPV104-like (RS-FVW) parameters, SURVEY.md §8(d)."""
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    pp = ctx.ppad

    def full(v):
        return torch.full((faces, pp), v, device=dev, dtype=dt)

    d = {"q_plus": q_plus, "q_minus": q_minus,
         "imposed_plus": torch.zeros((faces, 9, pp), device=dev, dtype=dt),
         "imposed_minus": torch.zeros((faces, 9, pp), device=dev, dtype=dt)}
    zp, zs, zpn, zsn = 2670.0 * 6000.0, 2670.0 * 3464.0, 2670.0 * 6000.0, 2670.0 * 3464.0
    eta_p, eta_s = 1.0 / (1.0 / zp + 1.0 / zpn), 1.0 / (1.0 / zs + 1.0 / zsn)
    imp = torch.tensor([zp, zs, zpn, zsn, eta_p, eta_s, 1 / zp, 1 / zs, 1 / zpn, 1 / zsn, 0.0], device=dev, dtype=dt)
    d["imp_and_eta"] = imp.repeat(faces, 1).contiguous()
    stress = torch.zeros((faces, 6, pp), device=dev, dtype=dt)
    stress[:, 0] = -120e6
    stress[:, 3] = 70e6 + 11e6 * torch.rand((faces, pp), generator=g, device=dev, dtype=dt)
    d["initial_stress"] = stress
    for k in ("slip1", "slip2", "slip_rate2", "accumulated_slip_magnitude", "traction1", "traction2", "rupture_time",
              "peak_slip_rate", "dyn_stress_time"):
        d[k] = full(0.0)
    d["slip_rate1"] = full(1e-16)
    d["slip_rate_magnitude"] = full(1e-16)
    d["mu"] = full(0.677 if law == "lsw" else 0.6)
    d["rupture_time_pending"] = torch.ones((faces, pp), device=dev, dtype=torch.uint8)
    d["dyn_stress_time_pending"] = torch.ones((faces, pp), device=dev, dtype=torch.uint8)
    if law == "lsw":
        d.update(d_c=full(0.4), mu_s=full(0.677), mu_d=full(0.525), cohesion=full(0.0), forced_rupture_time=full(1e9))
    else:
        d.update(rs_a=full(0.01), rs_sl0=full(0.4), rs_f0=full(0.6), rs_mu_w=full(0.2), rs_b=full(0.014),
                 rs_sr_w=full(0.1), state_variable=full(0.6),
                 convergence_inner=torch.zeros((faces, pp), device=dev, dtype=torch.uint8),
                 convergence_outer=torch.zeros((faces, pp), device=dev, dtype=torch.uint8))
    P = ctx.P
    d["resample_t"] = torch.eye(P, device=dev, dtype=dt).contiguous()
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--order", type=int, default=6)
    ap.add_argument("--real", type=int, default=4)
    ap.add_argument("--cells", type=int, default=50_000)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    ctx = api.Context(a.order, a.real)
    d = proxy.ProxyData.synthetic(ctx, a.cells)
    d.sort_by_face_relation()
    dr = proxy.DrProxyData(ctx, d)
    k = proxy.DrProxyKernels(ctx, d, dr)
    faces = dr.faces
    es = a.real
    O, B, P = a.order, ctx.B, ctx.P
    der_reals = sum(9 * proxy.nbasis(O - i) for i in range(O))
    by = {"godunov_dr": (2 * der_reals + 81 + 2 * O * 9 * P) * es * faces,
          "neigh_dr": (2 * 9 * B + 4 * (9 * P + 81)) * es * a.cells,
          "friction": (2 * O * 9 * P + 30 * P + 2 * 9 * P) * es * faces}
    print(f"order {a.order} real {a.real} cells {a.cells} faces {faces}")
    k.godunov_dr()
    fps, _, _ = proxy.friction_params(a.order, 1e-3, full_update_time=0.0)
    fd = {law: friction_data(ctx, faces, law, d.device, d.dtype, dr.q_plus * 0.0, dr.q_minus * 0.0)
          for law in ("lsw", "rsfvw")}
    runs = {"godunov_dr": k.godunov_dr, "neigh_dr": k.neigh_dr, "all_dr": k.all_dr,
            "friction_lsw": lambda: ctx.friction("lsw", fd["lsw"], fps, faces),
            "friction_rsfvw": lambda: ctx.friction("rsfvw", fd["rsfvw"], fps, faces)}
    for name, fn in runs.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
        ev[0].record()
        for i in range(a.steps):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        t = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)])) * 1e-3
        key = "friction" if name.startswith("friction") else name
        gbs = by[key] / t / 1e9 if key in by else float("nan")
        unit = faces if name != "neigh_dr" and name != "all_dr" else a.cells
        print(f"{name:16s} {t * 1e3:9.3f} ms  {unit / t / 1e6:8.2f} M{'faces' if unit == faces else 'cells'}/s  "
              f"{gbs:8.1f} GB/s(alg)")


if __name__ == "__main__":
    main()
