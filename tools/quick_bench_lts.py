#!/usr/bin/env python3
"""Clustered LTS (rate 2) on one GPU: cube mesh, horizontal time-cluster bands; element-updates/s of the LTS schedule
(seissol_b200/lts.py) next to global time stepping of the same mesh; development aid / profiles."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from seissol_b200 import api, lts, partition  # noqa: E402
from seissol_b200.distributed import element_values_torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--order", type=int, default=6)
    ap.add_argument("--real", type=int, default=4)
    ap.add_argument("--dims", type=int, nargs=3, default=[40, 70, 72])
    ap.add_argument("--clusters", type=int, default=3)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    dtype = torch.float32 if a.real == 4 else torch.float64
    ctx = api.Context(a.order, a.real)
    for ncl in (1, a.clusters):
        t0 = time.time()
        layout = partition.build_layout(*a.dims, 0, 1, ncl)
        plan = lts.build_plan(layout, ncl)
        gid = torch.from_numpy(plan.gid).to(dev)
        q = element_values_torch(torch, gid, 9 * ctx.B, 0, 1.0, dtype).reshape(plan.n, 9, ctx.B)
        dofs = torch.zeros((plan.n, 9, ctx.ld), dtype=dtype, device=dev)
        dofs[:, :, :ctx.B] = q
        dofs = dofs.reshape(plan.n, 9 * ctx.ld).contiguous()
        li = element_values_torch(torch, gid, 396, 1, 2e-3, dtype).contiguous()
        ni = element_values_torch(torch, gid, 324, 2, 2e-3, dtype).contiguous()
        be = lts.CudaLtsBackend(torch, ctx, plan, dofs, li, ni)
        solver = lts.LtsSolver(plan, be, a.order, dt0=1.0 / 2 ** (ncl - 1))
        setup = time.time() - t0
        for _ in range(2):
            solver.macro_step()
        torch.cuda.synchronize()
        be.launches = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            solver.macro_step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        # the same macro step captured once into a CUDA graph (the entry points are capture safe) and replayed
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                solver.macro_step()
        torch.cuda.current_stream().wait_stream(side)
        graph.replay()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(a.steps):
            graph.replay()
        g1.record()
        torch.cuda.synchronize()
        ms_graph = g0.elapsed_time(g1) / a.steps
        be.launches = be.launches * a.steps // (a.steps + 1)
        upd = plan.updates_per_macro_step()
        sizes = [r.end - r.start for r in plan.ranges]
        print(f"clusters {ncl}: cells {plan.n} per cluster {sizes} accumulating {plan.n_acc} with derivatives {plan.n_der} "
              f"taylor faces {plan.n_scratch} (setup {setup:.1f} s)")
        print(f"  macro step as one CUDA graph: {ms_graph:8.3f} ms  {plan.updates_per_macro_step() / ms_graph / 1e3:8.1f} M element-updates/s")
        print(f"  macro step {ms:8.3f} ms  updates/macro step {upd}  {upd / ms / 1e3:8.1f} M element-updates/s  "
              f"launches/macro step {be.launches // a.steps}  finite={bool(torch.isfinite(dofs).all())}")
        if ncl == 1:
            gts_ms_per_update = ms / upd
        else:
            gts_equiv = plan.n * 2 ** (ncl - 1) * gts_ms_per_update
            print(f"  same simulated time with global time stepping at dt_0: {gts_equiv:8.3f} ms  => LTS speed-up {gts_equiv / ms:.2f}x "
                  f"(update-count ratio {plan.n * 2 ** (ncl - 1) / upd:.2f}x)")


if __name__ == "__main__":
    main()
