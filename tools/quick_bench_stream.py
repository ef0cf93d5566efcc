#!/usr/bin/env python3
"""Timing of the two pure streaming kernels of the LTS path: k_taylor (I = sum_d c_d dQ(d), Time::evaluateBatched) and
k_lts_accumulate (buffers += I); both are HBM-bound, reported against the algorithmic bytes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from seissol_b200 import api, proxy  # noqa: E402


def timeit(fn, steps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e-3


def main():
    n = 1_000_000
    for real in (4, 8):
        ctx = api.Context(6, real)
        dt = torch.float32 if real == 4 else torch.float64
        derivs = torch.rand((n, ctx.dersize), device="cuda", dtype=dt)
        idofs = torch.zeros((n, ctx.qsize), device="cuda", dtype=dt)
        buffers = torch.zeros_like(idofs)
        coeffs = proxy.integrate_coeffs(6, 0.25, 0.75)
        t = timeit(lambda: ctx.taylor(api.strided(derivs), api.strided(idofs), coeffs, n))
        by = (1134 + 504) * real * n
        print(f"real {real}: taylor          {t * 1e3:7.3f} ms  {n / t / 1e6:8.1f} Melem/s  {by / t / 1e9:8.1f} GB/s(alg)")
        t = timeit(lambda: ctx.lts_accumulate(api.strided(buffers), api.strided(idofs), False, n))
        by = 3 * 504 * real * n
        print(f"real {real}: lts_accumulate  {t * 1e3:7.3f} ms  {n / t / 1e6:8.1f} Melem/s  {by / t / 1e9:8.1f} GB/s(alg)")
        vol = torch.rand(n, device="cuda", dtype=torch.float64) + 0.5
        mat = torch.rand((n, 3), device="cuda", dtype=torch.float64) + 1.0
        out = torch.zeros(8, device="cuda", dtype=torch.float64)
        t = timeit(lambda: ctx.volume_energies(api.strided(idofs), vol, mat, n, out))
        by = (504 * real + 4 * 8) * n
        print(f"real {real}: volume_energies {t * 1e3:7.3f} ms  {n / t / 1e6:8.1f} Melem/s  {by / t / 1e9:8.1f} GB/s(alg)")


if __name__ == "__main__":
    main()
