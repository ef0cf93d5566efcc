#!/usr/bin/env python3
"""Benchmark of the ADER-DG element-update hot path (SeisSol proxy kernel `all`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the element update (`all` = ader + local + neigh) over the whole element batch.
N = 1: BASELINE.json configs[1], the proxy (`proxy <cells> <steps> all`, src/Proxy/Runner.cpp:93-111), elastic
ORDER=6 single precision, 1M synthetic elements with the proxy's random fake topology.
N > 1: ONE cube mesh of ~1M tetrahedra PER GPU partitioned across the GPUs (BASELINE.json configs[4] at the
per-GPU size of configs[1]; weak scaling): copy-layer buffers travel to the neighbouring rank's ghost layer over
NCCL/NVLink while the interior elements are updated (seissol_b200/distributed.py).
`--replicas` runs N independent proxy replicas instead.
Prints ONE JSON line (rank 0).  Metric: element-updates/s (cells x steps / time), the
reference's own proxy metric (src/Proxy/Runner.cpp:122-144), plus non-zero GFLOP/s.

`--impl reference` times the reference's own CPU implementation of the same path (the
generated CPU kernels under oracle/_ref, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ORDER = 6
REAL = 4
KERNEL = "all"
CELLS = 1_000_000
CFG_NAME = {4: "elastic_o6_f32", 8: "elastic_o6_f64"}
# DRAM bytes per element per launch (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture at
# 296 000 elements, profiles/r1_v3b_ncu_summary_f32.txt): traffic ~ algorithmic bytes, i.e. no wasted re-reads
NCU_DRAM_BYTES_PER_ELEMENT = {4: {"k_local<ader>": (0.715280e9 + 559.575552e6) / 296000,
                                  "k_local<localwoader>": (1.663648e9 + 562.966016e6) / 296000,
                                  "k_neigh": (3.397901e9 + 666.347008e6) / 296000}}
FMA_PEAK = {4: 74.2, 8: 37.1}  # TFLOP/s, vector FMA, measured on this pool (tools/ubench/gen_ffma2.py); f64 = half rate


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(statistics.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons)}


def cpu_reference(real, budget_s, cells_hint=None):
    """The reference's generated CPU kernels (oracle/_ref) on the host cores: proxy `all` on a
    bounded sample.  Returns (elem-updates/s, cores, sample description, seconds, cells, steps)."""
    from oracle import aderdg as O
    from oracle import ref_lib
    cfg = CFG_NAME[real]
    if not ref_lib.available(cfg):
        return None
    R = ref_lib.RefLib(cfg)
    dt = np.float32 if real == 4 else np.float64
    coeffs = O.time_integrate_coeffs(ORDER, 0.0, 1.0, dt)
    cores = R.threads()
    # calibrate on a small batch, then size the sample for ~budget_s of CPU work
    cal = O.proxy_data(ORDER, 256 * cores, dt, seed=3)
    R.proxy(KERNEL, cal, coeffs, 1)
    t, _, _ = R.proxy(KERNEL, cal, coeffs, 1)
    rate = cal["Q"].shape[0] / max(t, 1e-9)
    steps = 2
    cells = int(max(1024, min(200_000, rate * budget_s / (steps + 1))))
    if cells_hint:
        cells = min(cells, cells_hint)
    d = O.proxy_data(ORDER, cells, dt, seed=4)
    R.proxy(KERNEL, d, coeffs, 1)                     # warm-up pass (Runner.cpp:93-95)
    t, _, _ = R.proxy(KERNEL, d, coeffs, steps)
    return cells * steps / t, cores, f"{cells} elements x {steps} passes of proxy `{KERNEL}`", t, cells, steps


def run_reference(args):
    """`--impl reference`: rank 0 alone times the CPU reference; other ranks exit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from seissol_b200 import proxy
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    from oracle import aderdg as O
    from oracle import ref_lib
    cfg = CFG_NAME[args.real]
    if not ref_lib.available(cfg):
        print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref/libssref_{cfg}.so not built"}))
        return 0
    R = ref_lib.RefLib(cfg)
    dt = np.float32 if args.real == 4 else np.float64
    coeffs = O.time_integrate_coeffs(ORDER, 0.0, 1.0, dt)
    cores = R.threads()
    cal = O.proxy_data(ORDER, 256 * cores, dt, seed=3)
    R.proxy(KERNEL, cal, coeffs, 1)
    t, _, _ = R.proxy(KERNEL, cal, coeffs, 1)
    rate = cal["Q"].shape[0] / max(t, 1e-9)
    # whole run (warm-up + steps) bounded to ~2 minutes
    cells = int(max(1024, min(args.cells, rate * 120.0 / (steps + warmup))))
    d = O.proxy_data(ORDER, cells, dt, seed=4)
    if warmup:
        R.proxy(KERNEL, d, coeffs, warmup)
    t, _, _ = R.proxy(KERNEL, d, coeffs, steps)
    value = cells * steps / t
    nz = proxy.nz_flops_per_element(ORDER, KERNEL)
    sample = f"{cells} elements per step (bounded sample of the {args.cells}-element workload)"
    line = {
        "impl": "reference", "metric": "element-updates/s", "value": value, "unit": "element-updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.real == 4 else "f64", "data": "synthetic",
        "gflops_nz": value * nz * 1e-9,
        "config": {"workload": f"proxy elastic ORDER={ORDER} {'single' if args.real == 4 else 'double'} kernel={KERNEL}",
                   "cells_per_gpu": args.cells, "sample_cells": cells},
        "cpu_baseline": {"value": value, "unit": "element-updates/s", "cores": cores, "kind": "reference",
                         "sample": sample,
                         "note": "reference generated kernels, yateto generic C++ back-end (LIBXSMM/PSpaMM unavailable offline)"},
        "e2e": {"value": value, "unit": "element-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from seissol_b200 import api, proxy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    ctx = api.Context(ORDER, args.real)
    data = proxy.ProxyData.synthetic(ctx, args.cells, device, seed=42 + rank)
    data.sort_by_face_relation()
    k = proxy.ProxyKernels(ctx, data)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        k.all()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # events: [start, after ader, after localwoader, after neigh] per step, all on the launching stream
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_begin.record(stream)
    for s in range(steps):
        ev[s][0].record(stream)
        k.ader()
        ev[s][1].record(stream)
        k.localwoader()
        ev[s][2].record(stream)
        k.neigh()
        ev[s][3].record(stream)
    t_end.record(stream)
    barrier()
    sampler.stop_flag.set()
    sampler.join()
    total_ms = t_begin.elapsed_time(t_end)
    t_ader = statistics.mean(ev[s][0].elapsed_time(ev[s][1]) for s in range(steps))
    t_local = statistics.mean(ev[s][1].elapsed_time(ev[s][2]) for s in range(steps))
    t_neigh = statistics.mean(ev[s][2].elapsed_time(ev[s][3]) for s in range(steps))
    finite = bool(torch.isfinite(data.dofs).all())
    if world > 1:
        t = torch.tensor([total_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # ---- end to end: host buffers in, host buffers out, every step (pinned memory).  The DOFs of a step are
    #      triple-buffered on the device and the three legs run on their own streams, so that the device->host copy of
    #      step s, the kernels of step s+1 and the host->device copy of step s+2 overlap (PCIe is full duplex); every
    #      step still moves its complete input and its complete result across PCIe inside the timed region.
    import copy
    e2e_steps = max(3, min(steps, 9))
    h_in = torch.empty(data.dofs.shape, dtype=data.dofs.dtype, pin_memory=True)
    h_in.copy_(data.dofs)
    n_slots = 3
    h_out = [torch.empty_like(h_in, pin_memory=True) for _ in range(n_slots)]
    slots = []
    for i in range(n_slots):
        d_i = copy.copy(data)
        d_i.dofs = data.dofs if i == 0 else torch.empty_like(data.dofs)
        slots.append((d_i, proxy.ProxyKernels(ctx, d_i)))
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    ev_in = [torch.cuda.Event() for _ in range(n_slots)]
    ev_done = [torch.cuda.Event() for _ in range(n_slots)]
    ev_out = [torch.cuda.Event() for _ in range(n_slots)]

    def e2e_run(n_steps):
        for st in range(n_steps):
            i = st % n_slots
            d_i, k_i = slots[i]
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_out[i])          # the previous result of this slot has left the device
                d_i.dofs.copy_(h_in, non_blocking=True)
                ev_in[i].record(s_in)
            stream.wait_event(ev_in[i])
            k_i.all(stream)
            ev_done[i].record(stream)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[i])
                h_out[i].copy_(d_i.dofs, non_blocking=True)
                ev_out[i].record(s_out)
        stream.wait_stream(s_out)
        stream.wait_stream(s_in)

    for ev_ in ev_out:
        ev_.record(stream)
    e2e_run(n_slots)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    s_in.wait_stream(stream)
    e2e_run(e2e_steps)
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    io_bytes = data.dofs.numel() * data.dofs.element_size()

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        ms_per_step = total_ms / steps
        value = world * args.cells / (ms_per_step * 1e-3)
        nz = proxy.nz_flops_per_element(ORDER, KERNEL)
        # dominant kernel = the longest of the three launches of a step
        names = {"k_local<ader>": (t_ader, "ader"), "k_local<localwoader>": (t_local, "localwoader"),
                 "k_neigh": (t_neigh, "neigh")}
        dom = max(names, key=lambda n: names[n][0])
        dom_ms = names[dom][0]
        dom_bytes = proxy.algorithmic_bytes_per_element(ORDER, args.real, names[dom][1])
        dom_flops = proxy.nz_flops_per_element(ORDER, names[dom][1])
        achieved = dom_bytes * args.cells / (dom_ms * 1e-3) / 1e9
        all_bytes = proxy.algorithmic_bytes_per_element(ORDER, args.real, KERNEL)
        line = {
            "metric": "element-updates/s", "value": value, "unit": "element-updates/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.real == 4 else "f64", "data": "synthetic",
            "gflops_nz": value * nz * 1e-9,
            "config": {"workload": f"proxy elastic ORDER={ORDER} {'single' if args.real == 4 else 'double'} kernel={KERNEL}, "
                                   f"{args.cells} elements per GPU (BASELINE.json configs[1])",
                       "cells_per_gpu": args.cells, "l2": "inputs exceed L2 (%.1f GiB resident per GPU)" % (data.bytes_resident() / 2 ** 30),
                       "parallelism": f"replicas x{world}" if world > 1 else "single GPU", "finite": finite},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"],
                         "traffic": (NCU_DRAM_BYTES_PER_ELEMENT[args.real][dom] * args.cells
                                     if args.real in NCU_DRAM_BYTES_PER_ELEMENT else None),
                         "traffic_note": "bytes per launch, scaled per element from the ncu --set full capture under profiles/",
                         "algorithmic_bytes": dom_bytes * args.cells, "peak_kind": peak_kind,
                         "ms_per_launch": dom_ms,
                         "fp32_fma": {"achieved_tflops": dom_flops * args.cells / (dom_ms * 1e-3) / 1e12,
                                      "peak_tflops": FMA_PEAK[args.real], "frac": dom_flops * args.cells / (dom_ms * 1e-3) / 1e12 / FMA_PEAK[args.real],
                                      "note": "non-zero flops of the dominant kernel vs the vector-FMA peak measured with tools/ubench (DESIGN.md §5)"},
                         "step": {"k_ader_ms": t_ader, "k_localwoader_ms": t_local, "k_neigh_ms": t_neigh,
                                  "all_algorithmic_GBs": all_bytes * args.cells / (ms_per_step * 1e-3) / 1e9,
                                  "all_nz_tflops": args.cells / (ms_per_step * 1e-3) * nz * 1e-12}},
            "e2e": {"value": world * args.cells * e2e_steps / (e2e_ms * 1e-3), "unit": "element-updates/s",
                    "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes, "steps": e2e_steps,
                    "what": "per step: DOFs pinned host -> device, proxy `all` through the C ABI, DOFs device -> pinned host; "
                            "three device slots, copy-in / kernels / copy-out of consecutive steps overlap on three streams"},
            "gpu_launches": 3 * steps,
            "clocks": sampler.result(),
        }
        if world == 1 and not args.no_cpu:
            cpu = cpu_reference(args.real, 12.0)
            if cpu is not None:
                v, cores, sample, secs, _, _ = cpu
                line["cpu_baseline"] = {"value": v, "unit": "element-updates/s", "cores": cores, "kind": "reference",
                                        "sample": sample, "seconds": secs,
                                        "note": "reference generated kernels, yateto generic C++ back-end "
                                                "(LIBXSMM/PSpaMM unavailable offline), OpenMP over all host threads"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "element-updates/s", "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref not built"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_partitioned(args):
    """N > 1: ONE cube mesh partitioned across the GPUs (BASELINE.json configs[4] at a per-GPU size that matches
    the single-GPU workload): slabs along x, copy/interior/ghost layers, halo exchange of the copy-layer buffers
    over NCCL/NVLink overlapped with the interior element update (seissol_b200/distributed.py).  Weak scaling:
    ~1M elements per GPU, global time stepping."""
    import torch
    import torch.distributed as dist
    from seissol_b200 import api, proxy
    from seissol_b200 import distributed as D

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    ctx = api.Context(ORDER, args.real)
    dtype = torch.float32 if args.real == 4 else torch.float64
    # slab of nxl x ny x nz cubes (5 tetrahedra each) per GPU
    ny, nz = 70, 72
    nxl = max(2, int(round(args.cells / (5.0 * ny * nz))))
    nxl += nxl % 2 if (nxl * world) % 2 else 0
    nx = nxl * world
    if args.lts > 1:
        return run_partitioned_lts(args, torch, dist, ctx, nx, ny, nz, rank, world, device, dtype)
    layout, data = D.build_partitioned(torch, dist, nx, ny, nz, rank, world, ctx.B, ctx.ld, dtype, device, scale=2e-3)
    data.sort_by_face_relation()
    coeffs = proxy.integrate_coeffs(ORDER, 0.0, proxy.TIMESTEP)
    backend = D.CudaBackend(ctx, data, coeffs)
    exchange = D.HaloExchange(torch, dist, layout, data.buffers)
    solver = D.PartitionedSolver(torch, layout, data, backend, exchange, overlap=not args.no_overlap)
    stream = torch.cuda.current_stream()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        solver.step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_begin.record(stream)
    for _ in range(steps):
        solver.step()
    t_end.record(stream)
    barrier()
    sampler.stop_flag.set()
    sampler.join()
    total_ms = t_begin.elapsed_time(t_end)
    finite = bool(torch.isfinite(data.dofs).all())

    # exchange alone (not overlapped), for the report
    x0 = torch.cuda.Event(enable_timing=True)
    x1 = torch.cuda.Event(enable_timing=True)
    barrier()
    x0.record(stream)
    for _ in range(5):
        exchange.exchange()
    x1.record(stream)
    barrier()
    exch_ms = x0.elapsed_time(x1) / 5

    # ---- end to end: DOFs of the own cells host -> device, one step, device -> host, every step
    e2e_steps = max(1, min(steps, 5))
    h_in = torch.empty(data.dofs.shape, dtype=data.dofs.dtype, pin_memory=True)
    h_in.copy_(data.dofs)
    h_out = torch.empty_like(h_in, pin_memory=True)
    for _ in range(2):
        data.dofs.copy_(h_in, non_blocking=True)
        solver.step()
        h_out.copy_(data.dofs, non_blocking=True)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        data.dofs.copy_(h_in, non_blocking=True)
        solver.step()
        h_out.copy_(data.dofs, non_blocking=True)
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    red = torch.tensor([total_ms, e2e_ms, exch_ms], device=device, dtype=torch.float64)
    dist.all_reduce(red, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, exch_ms = (float(v) for v in red.tolist())
    cnt = torch.tensor([layout.n_own, layout.n_copy, layout.n_ghost], device=device, dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    cells_total, copy_total, ghost_total = (int(v) for v in cnt.tolist())
    io_bytes = data.dofs.numel() * data.dofs.element_size()

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        ms_per_step = total_ms / steps
        value = cells_total / (ms_per_step * 1e-3)
        nz = proxy.nz_flops_per_element(ORDER, KERNEL)
        all_bytes = proxy.algorithmic_bytes_per_element(ORDER, args.real, KERNEL)
        achieved = all_bytes * cells_total / world / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": "element-updates/s", "value": value, "unit": "element-updates/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.real == 4 else "f64", "data": "synthetic",
            "gflops_nz": value * nz * 1e-9,
            "config": {"workload": f"cube mesh {nx}x{ny}x{nz} cubes x 5 tetrahedra = {cells_total} elements, elastic ORDER={ORDER} "
                                   f"{'single' if args.real == 4 else 'double'}, element update `all` (ader+local+neigh), "
                                   f"global time stepping, x-slab partition over {world} GPUs with NVLink halo exchange "
                                   "(BASELINE.json configs[4] at the per-GPU size of configs[1])",
                       "cells_per_gpu": cells_total // world, "copy_cells": copy_total, "ghost_cells": ghost_total,
                       "halo_bytes_per_step_per_gpu": exchange.bytes_per_exchange,
                       "halo_exchange_ms_alone": exch_ms, "overlap": solver.overlap,
                       "l2": "inputs exceed L2 (%.1f GiB resident per GPU)" % (
                           sum(x.numel() * x.element_size() for x in (data.dofs, data.buffers, data.local_int, data.neigh_int)) / 2 ** 30),
                       "parallelism": f"mesh partition x{world}, NCCL send/recv of copy-layer buffers", "finite": finite},
            "roofline": {"bound": "hbm", "kernel": "step (k_local<ader> + k_local<localwoader> + k_neigh per layer)",
                         "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "traffic": None, "peak_kind": peak_kind, "per": "GPU"},
            "e2e": {"value": cells_total * e2e_steps / (e2e_ms * 1e-3), "unit": "element-updates/s",
                    "h2d_bytes_per_step": io_bytes * world, "d2h_bytes_per_step": io_bytes * world, "steps": e2e_steps,
                    "what": "per step and GPU: DOFs pinned host -> device, partitioned element update through the C ABI "
                            "(incl. halo exchange), DOFs device -> pinned host"},
            "gpu_launches": backend.launches_per_step * steps,
            "clocks": sampler.result(),
        }
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0


def run_partitioned_lts(args, torch, dist, ctx, nx, ny, nz, rank, world, device, dtype):
    """`--lts C` (N > 1): BASELINE.json configs[4] proper -- the partitioned cube mesh with C rate-2 time clusters (bands in
    z), per-cluster halo exchange (seissol_b200/lts.py).  One step = one macro step (every cluster reaches t + dt_{C-1});
    value = element updates of all ranks per second."""
    from seissol_b200 import lts, partition, proxy
    from seissol_b200.distributed import element_values_torch
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    layout = partition.build_layout(nx, ny, nz, rank, world, args.lts)
    plan = lts.build_plan(layout, args.lts)
    gid = torch.from_numpy(plan.gid).to(device)
    dofs = torch.zeros((plan.n, 9, ctx.ld), dtype=dtype, device=device)
    dofs[:, :, :ctx.B] = element_values_torch(torch, gid, 9 * ctx.B, 0, 1.0, dtype).reshape(plan.n, 9, ctx.B)
    dofs = dofs.reshape(plan.n, 9 * ctx.ld).contiguous()
    li = element_values_torch(torch, gid, 396, 1, 2e-3, dtype).contiguous()
    ni = element_values_torch(torch, gid, 324, 2, 2e-3, dtype).contiguous()
    be = lts.CudaLtsBackend(torch, ctx, plan, dofs, li, ni, dist=dist)
    solver = lts.LtsSolver(plan, be, ORDER, dt0=proxy.TIMESTEP / 2 ** (args.lts - 1))
    stream = torch.cuda.current_stream()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        solver.macro_step()
    barrier()
    be.launches = 0
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        solver.macro_step()
    e1.record(stream)
    barrier()
    sampler.stop_flag.set()
    sampler.join()
    red = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    dist.all_reduce(red, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([plan.updates_per_macro_step(), plan.n, plan.n_acc, plan.n_der, plan.n_scratch], device=device,
                       dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    finite = bool(torch.isfinite(dofs).all())
    if rank == 0:
        ms = float(red.item()) / steps
        upd, cells, n_acc, n_der, n_tay = (int(v) for v in cnt.tolist())
        peaks, peak_kind = measured_peaks()
        all_bytes = proxy.algorithmic_bytes_per_element(ORDER, args.real, KERNEL)
        achieved = all_bytes * upd / world / (ms * 1e-3) / 1e9
        line = {
            "metric": "element-updates/s", "value": upd / (ms * 1e-3), "unit": "element-updates/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.real == 4 else "f64", "data": "synthetic",
            "gflops_nz": upd / (ms * 1e-3) * proxy.nz_flops_per_element(ORDER, KERNEL) * 1e-9,
            "config": {"workload": f"cube mesh {nx}x{ny}x{nz} cubes x 5 tetrahedra = {cells} elements, elastic ORDER={ORDER}, "
                                   f"clustered LTS rate 2 with {args.lts} clusters (bands in z), x-slab partition over {world} GPUs, "
                                   "per-cluster NCCL halo exchange (BASELINE.json configs[4])",
                       "cells_per_gpu": cells // world, "updates_per_macro_step": upd,
                       "gts_updates_for_the_same_simulated_time": cells * 2 ** (args.lts - 1),
                       "accumulating_cells": n_acc, "cells_with_derivatives": n_der, "taylor_faces": n_tay,
                       "launches_per_macro_step_per_gpu": be.launches // steps, "finite": finite,
                       "l2": "inputs exceed L2", "parallelism": f"mesh partition x{world}, LTS"},
            "roofline": {"bound": "hbm", "kernel": "macro step (all launches)", "achieved": achieved, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_kind": peak_kind, "per": "GPU"},
            "gpu_launches": be.launches, "clocks": sampler.result(),
        }
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=CELLS)
    ap.add_argument("--real", type=int, default=REAL, choices=[4, 8])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: do not overlap the halo exchange")
    ap.add_argument("--lts", type=int, default=1, help="N > 1: number of rate-2 time clusters (1 = global time stepping)")
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent proxy replicas instead of the partitioned mesh")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not args.replicas:
        return run_partitioned(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
