#!/usr/bin/env python3
"""Benchmark of the ADER-DG element-update hot path (SeisSol proxy kernels) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]

One "step" = one pass of the hot path over the whole synthetic batch.  Prints ONE JSON line (rank 0): metric
element-updates/s (cells x steps / time, the reference's own proxy metric, src/Proxy/Runner.cpp:122-144) plus non-zero
GFLOP/s, `roofline` (dominant kernel, CUDA-event timed, against MEASURED_PEAKS.json), `e2e` (host buffers in/out every
step), `cpu_baseline` (the reference's generated CPU kernels on the host cores) and `clocks`.

Configurations (BASELINE.json `configs`):
  elastic_f32 (default, N = 1)  configs[1]: proxy `all`, elastic ORDER=6 single precision, 1M elements, random fake topology
  elastic_f64                   configs[1] in double precision
  o4                            configs[0]: ORDER=4 single precision, 100k elements
  visco                         configs[2]: viscoelastic2 (3 mechanisms) ORDER=5, 1M elements, `all`
  dr_lsw | dr_rsfvw(_fast)      configs[3]: 200k fault faces (50k cells): all_dr = local + godunov_dr + friction law + neigh_dr
  lts (default, N > 1)          configs[4]: ONE cube mesh partitioned across the GPUs (x slabs, ~2.5M tetrahedra per GPU =
                                ~20M at N = 8), clustered LTS rate 2 (3 clusters), per-cluster NVLink halo exchange;
                                prints `parity_rel_l2` of a small partitioned run against the single-GPU run
  gts (`--lts 1`)               the same mesh with global time stepping
`--impl reference` times the reference's own CPU implementation of the same configuration (oracle/_ref, all host
threads; rank 0 alone under torchrun).
"""
import argparse
import copy
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    "elastic_f32": dict(kind="elastic", order=6, real=4, cells=1_000_000, ref="elastic_o6_f32", baseline="configs[1]"),
    "elastic_f64": dict(kind="elastic", order=6, real=8, cells=1_000_000, ref="elastic_o6_f64", baseline="configs[1], double"),
    "o4": dict(kind="elastic", order=4, real=4, cells=100_000, ref="elastic_o4_f32", baseline="configs[0]"),
    "visco": dict(kind="visco", order=5, real=4, cells=1_000_000, ref="visco_o5_f32", baseline="configs[2]"),
    "dr_lsw": dict(kind="dr", law="lsw", order=6, real=4, cells=50_000, ref="elastic_o6_f32", baseline="configs[3], linear slip weakening"),
    "dr_rsfvw": dict(kind="dr", law="rsfvw", order=6, real=4, cells=50_000, ref="elastic_o6_f32",
                     baseline="configs[3], rate-and-state fast velocity weakening"),
    # the same with ssb_friction_params_t::rs_float_resolution_exit = 1 (Newton iterations stop at float resolution)
    "dr_rsfvw_fast": dict(kind="dr", law="rsfvw", order=6, real=4, cells=50_000, ref="elastic_o6_f32", float_exit=True,
                          baseline="configs[3], rate-and-state fast velocity weakening, float-resolution Newton exit"),
}
LTS_CELLS_PER_GPU = 2_500_000
# vector-FMA peak MEASURED on this pool (tools/ubench/fma_peak.cu -> profiles/r2_ubench_fma_peak.txt): FFMA 73.4 TFLOP/s
# (nominal 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.4), DFMA 34.3 TFLOP/s
FMA_PEAK_TFLOPS = {4: 73.4, 8: 34.3}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


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

    def finish(self):
        self.stop_flag.set()
        self.join()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(statistics.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons)}


# ======================================================================================================================
# Workloads of the single-GPU configurations.  A workload knows its kernels ("segments", timed one by one with CUDA
# events), the tensors that cross PCIe in the end-to-end measurement, and its algorithmic bytes / non-zero flops.
class Workload:
    name = ""
    unit_name = "elements"

    def segments(self):          # [(kernel name, launch function(stream), algorithmic bytes, NZ flops, launches)]
        raise NotImplementedError

    def io_tensors(self):        # state that enters and leaves the device every step in the e2e measurement
        raise NotImplementedError

    def clone_slot(self):        # a second instance working on its own copy of io_tensors (e2e pipelining)
        raise NotImplementedError

    def step(self, stream=None):
        for _, fn, _, _, _ in self.segments():
            fn(stream)

    def finite(self):
        import torch
        return all(bool(torch.isfinite(t).all()) for t in self.io_tensors())


class ElasticWorkload(Workload):
    def __init__(self, cfg, device, seed=42, data=None):
        from seissol_b200 import api, proxy
        self.proxy, self.cfg = proxy, cfg
        self.order, self.real, self.cells = cfg["order"], cfg["real"], cfg["cells"]
        self.ctx = api.Context(self.order, self.real)
        if data is None:
            data = proxy.ProxyData.synthetic(self.ctx, self.cells, device, seed=seed)
            data.sort_by_face_relation()
        self.data = data
        self.k = proxy.ProxyKernels(self.ctx, data)
        prec = "single" if self.real == 4 else "double"
        self.name = (f"proxy elastic ORDER={self.order} {prec} precision kernel=all, {self.cells} elements with the proxy's random "
                     f"fake topology (BASELINE.json {cfg['baseline']})")
        local_fp = ("k_local5" if self.real == 4 else "k_local4") + " (volume + local flux, writes face projections)"
        self.kernel_names = {"ader": "k_local<ader>", "localwoader": local_fp if self.k.faceproj else "k_local<localwoader>",
                             "neigh": "k_neigh_fp (face projections)" if self.k.faceproj else "k_neigh"}

    def segments(self):
        p, o, r, n = self.proxy, self.order, self.real, self.cells
        return [(self.kernel_names[s], getattr(self.k, s), p.algorithmic_bytes_per_element(o, r, s) * n,
                 p.nz_flops_per_element(o, s) * n, 1) for s in ("ader", "localwoader", "neigh")]

    def step_bytes(self):
        """`all` of SURVEY.md §8(d): 9 B NQ + S + 8 NQ^2 reals (I counted once), not the sum of the three kernels."""
        return self.proxy.algorithmic_bytes_per_element(self.order, self.real, "all") * self.cells

    def io_tensors(self):
        return [self.data.dofs]

    def clone_slot(self):
        import torch
        d = copy.copy(self.data)
        d.dofs = torch.empty_like(self.data.dofs)
        return ElasticWorkload(self.cfg, self.data.device, data=d)

    def resident_bytes(self):
        return self.data.bytes_resident()


class ViscoWorkload(Workload):
    def __init__(self, cfg, device, seed=42, data=None):
        from seissol_b200 import api, proxy
        self.proxy, self.cfg = proxy, cfg
        self.real, self.cells = cfg["real"], cfg["cells"]
        self.ctx = api.Context(5, self.real, equation=1, mechanisms=3)
        if data is None:
            data = proxy.ViscoProxyData.synthetic(self.ctx, self.cells, device, seed=seed)
            data.sort_by_face_relation()
        self.data = data
        self.k = proxy.ViscoProxyKernels(self.ctx, data)
        self.name = (f"proxy viscoelastic2 (3 mechanisms) ORDER=5 {'single' if self.real == 4 else 'double'} precision kernel=all, "
                     f"{self.cells} elements (BASELINE.json {cfg['baseline']})")

    def segments(self):
        p, r, n = self.proxy, self.real, self.cells
        return [(f"k_visco_{s}", getattr(self.k, s), p.visco_algorithmic_bytes_per_element(r, s) * n,
                 p.visco_nz_flops_per_element(s) * n, 1) for s in ("ader", "localwoader", "neigh")]

    def step_bytes(self):
        return self.proxy.visco_algorithmic_bytes_per_element(self.real, "all") * self.cells

    def io_tensors(self):
        return [self.data.dofs, self.data.dofs_ane]

    def clone_slot(self):
        import torch
        d = copy.copy(self.data)
        d.dofs = torch.empty_like(self.data.dofs)
        d.dofs_ane = torch.empty_like(self.data.dofs_ane)
        return ViscoWorkload(self.cfg, self.data.device, data=d)

    def resident_bytes(self):
        return self.data.bytes_resident()


class DrWorkload(Workload):
    """all_dr of the proxy (local + godunov_dr + neigh_dr, src/Proxy/KernelHost.h:29-56) with the friction law between the
    space-time interpolation and the nodal flux, the order of TimeCluster::computeDynamicRupture*
    (src/Solver/TimeStepping/TimeCluster.cpp:282-344)."""
    unit_name = "cells (4 fault faces each)"

    def __init__(self, cfg, device, seed=42, data=None, shared=None):
        import torch
        from seissol_b200 import api, proxy
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import quick_bench_dr
        self.proxy, self.cfg, self.law = proxy, cfg, cfg["law"]
        self.order, self.real, self.cells = cfg["order"], cfg["real"], cfg["cells"]
        self.ctx = api.Context(self.order, self.real)
        if data is None:
            data = proxy.ProxyData.synthetic(self.ctx, self.cells, device, seed=seed)
        self.data = data
        if shared is None:
            dr = proxy.DrProxyData(self.ctx, data)
            fric = quick_bench_dr.friction_data(self.ctx, dr.faces, self.law, data.device, data.dtype, dr.q_plus, dr.q_minus)
            # the friction law writes the imposed state that the nodal flux reads (CellDRMapping::godunov points into it)
            fric["imposed_plus"], fric["imposed_minus"] = dr.imposed_plus, torch.zeros_like(dr.imposed_plus)
            dr.flux_solver.mul_(1e-6)     # keeps the DOFs O(1) over the run: the imposed tractions are O(1e7) Pa
            shared = (dr, fric)
        self.dr, self.fric = shared
        self.faces = self.dr.faces
        self.k = proxy.DrProxyKernels(self.ctx, data, self.dr)
        self.fparams, _, _ = proxy.friction_params(self.order, 1e-3, full_update_time=0.0,
                                                       rs_float_resolution_exit=bool(cfg.get("float_exit", False)))
        self.torch = torch
        self.name = (f"proxy all_dr + friction ({'linear slip weakening' if self.law == 'lsw' else 'rate-and-state fast velocity weakening'}), "
                     f"elastic ORDER={self.order} {'single' if self.real == 4 else 'double'} precision, {self.faces} fault faces = "
                     f"{self.cells} cells (BASELINE.json {cfg['baseline']})")

    def _local(self, stream=None):
        # like the reference proxy's all_dr the fault reads FAKE derivative families that are never updated
        # (src/Proxy/Allocator.cpp:204-213): `local` = ader + localwoader on the cells, all faces DynamicRupture
        d, dr = self.data, self.dr
        from seissol_b200 import api
        self.ctx.ader(api.strided(d.dofs), api.strided(d.buffers), None, api.strided(d.local_int), self.k.coeffs, d.cells,
                      stream=stream)
        self.ctx.local(api.strided(d.dofs), api.strided(d.buffers), api.strided(d.local_int), dr.dr_face_types, d.cells,
                       stream=stream)

    def _friction(self, stream=None):
        self.ctx.friction(self.law, self.fric, self.fparams, self.faces, stream=stream)

    def segments(self):
        es, O, P, B = self.real, self.order, self.ctx.P, self.ctx.B
        der_reals = sum(9 * self.proxy.nbasis(O - i) for i in range(O))
        f, n = self.faces, self.cells
        by_local = (5 * 9 * B + 144) * es * n                       # ader + volume (no local flux on fault faces)
        by_interp = (2 * der_reals + 81 + 2 * O * 9 * P) * es * f
        by_fric = (2 * O * 9 * P + 30 * P + 2 * 9 * P) * es * f
        by_nodal = (2 * 9 * B + 4 * (9 * P + 81)) * es * n
        fl = self.proxy.FLOPS[O]
        nz_local = (fl["derivative"][0] + fl["volume"][0]) * n     # every face is a fault face: no local flux
        nz_interp = 692226.0 * f if O == 6 else 0.0
        nz_nodal = 56387.0 * f if O == 6 else 0.0
        return [("k_local<ader> + k_local5 (volume)", self._local, by_local, nz_local, 2),
                ("k_dr_interp (godunov_dr)", self.k.godunov_dr, by_interp, nz_interp, 1),
                (f"k_friction<{self.law}>", self._friction, by_fric, 0.0, 1),
                ("k_neigh_dr (nodal flux)", self.k.neigh_dr, by_nodal, nz_nodal, 1)]

    def io_tensors(self):
        return [self.data.dofs]

    def clone_slot(self):
        d = copy.copy(self.data)
        d.dofs = self.torch.empty_like(self.data.dofs)
        return DrWorkload(self.cfg, self.data.device, data=d, shared=(self.dr, self.fric))

    def resident_bytes(self):
        return self.data.bytes_resident() + sum(t.numel() * t.element_size() for t in self.fric.values()
                                                if hasattr(t, "numel"))


def make_workload(cfg, device, seed=42):
    return {"elastic": ElasticWorkload, "visco": ViscoWorkload, "dr": DrWorkload}[cfg["kind"]](cfg, device, seed)


# ======================================================================================================================
# CPU reference: the reference's generated kernels (oracle/_ref) on the host cores
def _ref_threads(R):
    """All host cores, whatever OMP_NUM_THREADS says (torchrun exports 1)."""
    n = os.cpu_count() or 1
    try:
        return R.set_threads(n)
    except Exception:
        return R.threads()


def cpu_reference(cfg, budget_s, full_cells=None, steps=2, warmup=1):
    """Times the reference CPU implementation of the configuration.  Returns dict(value, cores, sample, seconds, cells, steps,
    kind) or None.  `full_cells`: run exactly this many cells (the reference arm); otherwise a sample sized for ~budget_s."""
    from oracle import aderdg as O
    from oracle import ref_lib
    if not ref_lib.available(cfg["ref"]):
        return None
    kind, order = cfg["kind"], cfg["order"]
    dt = np.float32 if cfg["real"] == 4 else np.float64
    if kind == "elastic":
        R = ref_lib.RefLib(cfg["ref"])
        cores = _ref_threads(R)
        coeffs = O.time_integrate_coeffs(order, 0.0, 1.0, dt)
        cal = O.proxy_data(order, 256 * cores, dt, seed=3)
        R.proxy("all", cal, coeffs, 1)
        t, _, _ = R.proxy("all", cal, coeffs, 1)
        rate = cal["Q"].shape[0] / max(t, 1e-9)
        cells = full_cells or int(max(1024, min(cfg["cells"], rate * budget_s / (steps + warmup))))
        d = O.proxy_data(order, cells, dt, seed=4)
        if warmup:
            R.proxy("all", d, coeffs, warmup)                # warm-up pass (Runner.cpp:93-95)
        t, _, _ = R.proxy("all", d, coeffs, steps)
        return dict(value=cells * steps / t, cores=cores, seconds=t, cells=cells, steps=steps, kind="reference",
                    sample=f"{cells} elements x {steps} passes of proxy `all` (reference generated kernels, yateto generic C++ "
                           "back-end: LIBXSMM/PSpaMM unavailable offline; OpenMP over all host threads)")
    if kind == "visco":
        from oracle import visco as V
        R = ref_lib.ViscoRefLib(cfg["ref"])
        cores = _ref_threads(R)
        coeffs = O.time_integrate_coeffs(5, 0.0, 1.0, dt)
        cal = V.proxy_data(5, 256 * cores, dt, seed=3)
        R.proxy_all(cal, coeffs, 1)
        t, _, _ = R.proxy_all(cal, coeffs, 1)
        rate = cal["Q"].shape[0] / max(t, 1e-9)
        cells = full_cells or int(max(1024, min(cfg["cells"], rate * budget_s / (steps + warmup))))
        d = V.proxy_data(5, cells, dt, seed=4)
        if warmup:
            R.proxy_all(d, coeffs, warmup)
        t, _, _ = R.proxy_all(d, coeffs, steps)
        return dict(value=cells * steps / t, cores=cores, seconds=t, cells=cells, steps=steps, kind="reference",
                    sample=f"{cells} elements x {steps} passes of viscoelastic2 `all` (reference generated kernels, generic back-end, "
                           "OpenMP over all host threads)")
    if kind == "dr":
        # reference generated kernels for ader / local / space-time interpolation / nodal flux (OpenMP), the friction law
        # from the numpy port (the reference's friction headers do not compile stand-alone): kind "port"
        from oracle import friction as F
        R = ref_lib.RefLib(cfg["ref"])
        cores = _ref_threads(R)
        tab = O.load_tables(order, dt)
        # the numpy port iterates face by face: rate-and-state in single precision never meets the reference's 1e-8
        # tolerance and runs all 10 x 60 Newton iterations per time point -> a very small sample
        cells = full_cells or (1500 if cfg["law"] == "lsw" else 40)
        faces = 4 * cells
        rng = np.random.default_rng(5)
        d = O.proxy_data(order, cells, dt, seed=4)
        coeffs = O.time_integrate_coeffs(order, 0.0, 1e-3, dt)
        t0 = time.perf_counter()
        I, ders = R.ader(d["Q"], d["star"], coeffs, True)
        ft = np.full((cells, 4), 3, dtype=np.uint8)
        R.local(d["Q"], I, d["star"], d["Aplus"], ft)
        t_local = time.perf_counter() - t0
        plus, minus = rng.integers(0, cells, faces), rng.integers(0, cells, faces)
        info = np.stack([rng.integers(0, 4, faces), rng.integers(0, 4, faces), rng.integers(0, 3, faces)], 1).astype(np.uint8)
        tinv = rng.uniform(-1, 1, (faces, 9, 9)).astype(dt)
        pts, _ = O.gauss_legendre_shifted(order, 0.0, 1e-3)
        cpt = np.stack([O.time_point_coeffs(order, tau, dt) for tau in pts]).astype(dt)
        t_interp = R.dr_interpolate_timed([x[plus] for x in ders], [x[minus] for x in ders], tinv, info, cpt)
        st, imp, params = F.fault_state(cfg["law"], faces, order, dt)
        qip, qim = F.fault_qinterpolated(faces, order, dt)
        law = (F.LinearSlipWeakening if cfg["law"] == "lsw" else F.RateAndStateFastVelocityWeakening)(params, tab["resample"])
        dts = O.friction_delta_t(pts, 1e-3)
        wts = np.polynomial.legendre.leggauss(order)[1] / 2.0
        t0 = time.perf_counter()
        F.evaluate(law, st, imp, qip, qim, 0.0, dts, wts)
        t_fric = time.perf_counter() - t0
        t = t_local + t_interp + t_fric
        return dict(value=cells / t, cores=cores, seconds=t, cells=cells, steps=1, kind="port",
                    sample=f"{cells} cells / {faces} fault faces, one pass: ader+local {t_local:.2f} s and space-time interpolation "
                           f"{t_interp:.2f} s on the reference's generated kernels (OpenMP, {cores} threads), friction law {t_fric:.2f} s "
                           "on the numpy port (1 thread); nodal flux not included")
    return None


# ======================================================================================================================
def run_single(args, cfg_name):
    import torch
    import torch.distributed as dist
    from seissol_b200 import proxy

    cfg = dict(CONFIGS[cfg_name])
    if args.cells:
        cfg["cells"] = args.cells
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
    w = make_workload(cfg, device, seed=42 + rank)
    segs = w.segments()
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        w.step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(segs) + 1)] for _ in range(steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin.record(stream)
    for s in range(steps):
        ev[s][0].record(stream)
        for i, (_, fn, _, _, _) in enumerate(segs):
            fn(None)
            ev[s][i + 1].record(stream)
    t_end.record(stream)
    barrier()
    clocks = sampler.finish()
    total_ms = t_begin.elapsed_time(t_end)
    seg_ms = [statistics.mean(ev[s][i].elapsed_time(ev[s][i + 1]) for s in range(steps)) for i in range(len(segs))]
    finite = w.finite()
    if world > 1:
        t = torch.tensor([total_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # ---- end to end: the state of a step enters from pinned host memory and leaves to pinned host memory EVERY step,
    #      through the public API.  Three device slots on three streams: the copy-out of step s, the kernels of step s+1
    #      and the copy-in of step s+2 overlap (PCIe is full duplex).
    e2e_steps = max(3, min(steps, 9))
    n_slots = 3
    slots = [w] + [w.clone_slot() for _ in range(n_slots - 1)]
    h_in = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in w.io_tensors()]
    for h, t in zip(h_in, w.io_tensors()):
        h.copy_(t)
    h_out = [[torch.empty_like(h, pin_memory=True) for h in h_in] for _ in range(n_slots)]
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    ev_in = [torch.cuda.Event() for _ in range(n_slots)]
    ev_done = [torch.cuda.Event() for _ in range(n_slots)]
    ev_out = [torch.cuda.Event() for _ in range(n_slots)]

    def e2e_run(n_steps):
        for st in range(n_steps):
            i = st % n_slots
            wi = slots[i]
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_out[i])          # the previous result of this slot has left the device
                for t, h in zip(wi.io_tensors(), h_in):
                    t.copy_(h, non_blocking=True)
                ev_in[i].record(s_in)
            stream.wait_event(ev_in[i])
            wi.step(stream)
            ev_done[i].record(stream)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[i])
                for t, h in zip(wi.io_tensors(), h_out[i]):
                    h.copy_(t, non_blocking=True)
                ev_out[i].record(s_out)
        stream.wait_stream(s_out)
        stream.wait_stream(s_in)

    for e in ev_out:
        e.record(stream)
    e2e_run(n_slots)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
    io_bytes = sum(t.numel() * t.element_size() for t in w.io_tensors())

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        ms_per_step = total_ms / steps
        cells = cfg["cells"]
        value = world * cells / (ms_per_step * 1e-3)
        nz_total = sum(s[3] for s in segs)
        by_total = w.step_bytes() if hasattr(w, "step_bytes") else sum(s[2] for s in segs)
        dom = max(range(len(segs)), key=lambda i: seg_ms[i])
        dname, _, dbytes, dflops, _ = segs[dom]
        achieved = dbytes / (seg_ms[dom] * 1e-3) / 1e9
        fma_peak = FMA_PEAK_TFLOPS[cfg["real"]]
        line = {
            "metric": "element-updates/s", "value": value, "unit": "element-updates/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if cfg["real"] == 4 else "f64", "data": "synthetic",
            "gflops_nz": value * nz_total / cells * 1e-9,
            "config": {"workload": w.name, "name": cfg_name, "cells_per_gpu": cells,
                       "l2": "inputs exceed L2 (%.1f GiB resident per GPU)" % (w.resident_bytes() / 2 ** 30),
                       "parallelism": f"replicas x{world}" if world > 1 else "single GPU", "finite": finite},
            "roofline": {"bound": "hbm", "kernel": dname, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": None,
                         "traffic_note": "not measured in this run (ncu captures of the same kernels: profiles/r2_ncu_*.txt)",
                         "algorithmic_bytes": dbytes, "peak_kind": peak_kind, "ms_per_launch": seg_ms[dom],
                         "binding_ceiling": "hbm" if achieved / peaks["hbm_gbs"] >= dflops / (seg_ms[dom] * 1e-3) / 1e12 / fma_peak
                         else "fp32/fp64 vector FMA (the kernel's arithmetic intensity is above the ridge)",
                         "fma": {"achieved_tflops": dflops / (seg_ms[dom] * 1e-3) / 1e12, "peak_tflops": fma_peak,
                                 "frac": dflops / (seg_ms[dom] * 1e-3) / 1e12 / fma_peak,
                                 "note": "non-zero flops of the dominant kernel vs the vector-FMA peak (profiles/r2_ubench_fma_peak.txt)"},
                         "step": {"kernels_ms": {segs[i][0]: seg_ms[i] for i in range(len(segs))},
                                  "algorithmic_GBs": by_total / (ms_per_step * 1e-3) / 1e9,
                                  "frac_of_hbm_peak": by_total / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                  "nz_tflops": nz_total / (ms_per_step * 1e-3) * 1e-12,
                                  "frac_of_fma_peak": nz_total / (ms_per_step * 1e-3) * 1e-12 / fma_peak}},
            "e2e": {"value": world * cells * e2e_steps / (e2e_ms * 1e-3), "unit": "element-updates/s",
                    "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes, "steps": e2e_steps,
                    "what": "per step: state (DOFs) pinned host -> device, the step through the C ABI, state device -> pinned host; "
                            "three device slots, copy-in / kernels / copy-out of consecutive steps overlap on three streams"},
            "gpu_launches": sum(s[4] for s in segs) * steps,
            "clocks": clocks,
        }
        if cfg["kind"] == "dr":
            line["config"]["fault_faces"] = w.faces
            line["fault_face_updates_per_s"] = world * w.faces / (ms_per_step * 1e-3)
        if world == 1 and not args.no_cpu:
            cpu = cpu_reference(cfg, 12.0)
            if cpu is not None:
                line["cpu_baseline"] = {"value": cpu["value"], "unit": "element-updates/s", "cores": cpu["cores"],
                                        "kind": cpu["kind"], "sample": cpu["sample"], "seconds": cpu["seconds"]}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "element-updates/s", "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref not built"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# ======================================================================================================================
def cube_dims(cells_per_gpu, world):
    """Slab of nxl x ny x nz cubes (5 tetrahedra each) per GPU."""
    ny, nz = 70, 72
    nxl = max(2, int(round(cells_per_gpu / (5.0 * ny * nz))))
    nxl += nxl % 2 if (nxl * world) % 2 else 0
    return nxl * world, ny, nz


def lts_parity(torch, dist, ctx, rank, world, device, dtype, clusters, order):
    """Multi-GPU correctness evidence inside the benchmark run: a small cube mesh advanced by two macro steps on the
    partition (this run's ranks, NCCL halo exchange) and on ONE GPU (every rank repeats the single-rank run of the whole
    mesh); relative L2 difference of the DOFs of the own cells, maximum over the ranks."""
    from seissol_b200 import lts, partition, proxy
    from seissol_b200.distributed import element_values_torch
    nx, ny, nz = 2 * world, 6, 3 * clusters

    def run(r, w, d):
        layout = partition.build_layout(nx, ny, nz, r, w, clusters)
        plan = lts.build_plan(layout, clusters)
        gid = torch.from_numpy(plan.gid).to(device)
        dofs = torch.zeros((plan.n, 9, ctx.ld), dtype=dtype, device=device)
        dofs[:, :, :ctx.B] = element_values_torch(torch, gid, 9 * ctx.B, 0, 1.0, dtype).reshape(plan.n, 9, ctx.B)
        dofs = dofs.reshape(plan.n, 9 * ctx.ld).contiguous()
        li = element_values_torch(torch, gid, 396, 1, 2e-3, dtype).contiguous()
        ni = element_values_torch(torch, gid, 324, 2, 2e-3, dtype).contiguous()
        be = lts.CudaLtsBackend(torch, ctx, plan, dofs, li, ni, dist=d)
        solver = lts.LtsSolver(plan, be, order, dt0=proxy.TIMESTEP / 2 ** (clusters - 1))
        for _ in range(2):
            solver.macro_step()
        torch.cuda.synchronize()
        return plan.gid, dofs

    gid_p, q_p = run(rank, world, dist)
    dist.barrier()
    gid_s, q_s = run(0, 1, None)
    pos = np.empty(len(gid_s), dtype=np.int64)
    pos[gid_s] = np.arange(len(gid_s))
    ref = q_s[torch.from_numpy(pos[gid_p]).to(device)]
    err = float((q_p - ref).double().norm() / ref.double().norm())
    t = torch.tensor([err], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), nx * ny * nz * 5


def gts_parity(torch, dist, ctx, rank, world, device, dtype, order):
    from seissol_b200 import distributed as D
    from seissol_b200 import proxy
    nx, ny, nz = 2 * world, 6, 6

    def run(r, w, d):
        layout, data = D.build_partitioned(torch, d, nx, ny, nz, r, w, ctx.B, ctx.ld, dtype, device, scale=2e-3)
        data.sort_by_face_relation()
        backend = D.CudaBackend(ctx, data, proxy.integrate_coeffs(order, 0.0, proxy.TIMESTEP))
        exchange = D.HaloExchange(torch, d, layout, backend.halo_tensor())
        solver = D.PartitionedSolver(torch, layout, data, backend, exchange, overlap=w > 1)
        for _ in range(3):
            solver.step()
        torch.cuda.synchronize()
        return layout.gid[:layout.n_own], data.dofs

    gid_p, q_p = run(rank, world, dist)
    dist.barrier()
    gid_s, q_s = run(0, 1, dist)
    pos = np.empty(len(gid_s), dtype=np.int64)
    pos[gid_s] = np.arange(len(gid_s))
    ref = q_s[torch.from_numpy(pos[gid_p]).to(device)]
    err = float((q_p - ref).double().norm() / ref.double().norm())
    t = torch.tensor([err], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), nx * ny * nz * 5


def e2e_staged(torch, stream, device, dofs, step_fn, e2e_steps, barrier):
    """End-to-end loop for a solver that owns its DOF array: host -> staging buffer (copy stream) -> DOFs (device copy on
    the compute stream) -> step -> staging buffer -> host (copy stream); two staging slots per direction keep PCIe busy
    while the kernels run."""
    n_slots = 2
    h_in = torch.empty(dofs.shape, dtype=dofs.dtype, pin_memory=True)
    h_in.copy_(dofs)
    h_out = [torch.empty_like(h_in, pin_memory=True) for _ in range(n_slots)]
    d_in = [torch.empty_like(dofs) for _ in range(n_slots)]
    d_out = [torch.empty_like(dofs) for _ in range(n_slots)]
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    ev_in = [torch.cuda.Event() for _ in range(n_slots)]
    ev_used = [torch.cuda.Event() for _ in range(n_slots)]
    ev_done = [torch.cuda.Event() for _ in range(n_slots)]
    ev_out = [torch.cuda.Event() for _ in range(n_slots)]

    def run(n):
        for st in range(n):
            i = st % n_slots
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_used[i])
                d_in[i].copy_(h_in, non_blocking=True)
                ev_in[i].record(s_in)
            stream.wait_event(ev_in[i])
            dofs.copy_(d_in[i], non_blocking=True)
            ev_used[i].record(stream)
            step_fn()
            stream.wait_event(ev_out[i])
            d_out[i].copy_(dofs, non_blocking=True)
            ev_done[i].record(stream)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[i])
                h_out[i].copy_(d_out[i], non_blocking=True)
                ev_out[i].record(s_out)
        stream.wait_stream(s_out)
        stream.wait_stream(s_in)

    for i in range(n_slots):
        ev_used[i].record(stream)
        ev_out[i].record(stream)
    run(n_slots)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    s_in.wait_stream(stream)
    run(e2e_steps)
    e1.record(stream)
    barrier()
    return e0.elapsed_time(e1), dofs.numel() * dofs.element_size()


def run_partitioned(args):
    """N > 1: ONE cube mesh partitioned across the GPUs (BASELINE.json configs[4]).  Default: clustered LTS rate 2 with 3
    time clusters, ~2.5M tetrahedra per GPU (20M at N = 8), per-cluster halo exchange over NCCL/NVLink overlapped with the
    other clusters' kernels (seissol_b200/lts.py).  `--lts 1`: global time stepping (seissol_b200/distributed.py)."""
    import torch
    import torch.distributed as dist
    from seissol_b200 import api, proxy
    from seissol_b200 import distributed as D

    order, real = 6, 4
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    ctx = api.Context(order, real)
    dtype = torch.float32
    nx, ny, nz = cube_dims(args.cells or LTS_CELLS_PER_GPU, world)
    stream = torch.cuda.current_stream()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    peaks, peak_kind = measured_peaks()
    all_bytes = proxy.algorithmic_bytes_per_element(order, real, "all")
    nzf = proxy.nz_flops_per_element(order, "all")
    if args.lts > 1:
        from seissol_b200 import lts, partition
        parity, parity_cells = lts_parity(torch, dist, ctx, rank, world, device, dtype, args.lts, order)
        layout = partition.build_layout(nx, ny, nz, rank, world, args.lts)
        plan = lts.build_plan(layout, args.lts)
        gid = torch.from_numpy(plan.gid).to(device)
        dofs = torch.zeros((plan.n, 9, ctx.ld), dtype=dtype, device=device)
        dofs[:, :, :ctx.B] = D.element_values_torch(torch, gid, 9 * ctx.B, 0, 1.0, dtype).reshape(plan.n, 9, ctx.B)
        dofs = dofs.reshape(plan.n, 9 * ctx.ld).contiguous()
        li = D.element_values_torch(torch, gid, 396, 1, 2e-3, dtype).contiguous()
        ni = D.element_values_torch(torch, gid, 324, 2, 2e-3, dtype).contiguous()
        be = lts.CudaLtsBackend(torch, ctx, plan, dofs, li, ni, dist=dist)
        solver = lts.LtsSolver(plan, be, order, dt0=proxy.TIMESTEP / 2 ** (args.lts - 1))
        step_fn = solver.macro_step
        cnt = torch.tensor([plan.updates_per_macro_step(), plan.n, plan.n_acc, plan.n_der, plan.n_scratch, plan.n_ghost],
                           device=device, dtype=torch.int64)
    else:
        parity, parity_cells = gts_parity(torch, dist, ctx, rank, world, device, dtype, order)
        layout, data = D.build_partitioned(torch, dist, nx, ny, nz, rank, world, ctx.B, ctx.ld, dtype, device, scale=2e-3)
        data.sort_by_face_relation()
        backend = D.CudaBackend(ctx, data, proxy.integrate_coeffs(order, 0.0, proxy.TIMESTEP))
        exchange = D.HaloExchange(torch, dist, layout, backend.halo_tensor())
        solver = D.PartitionedSolver(torch, layout, data, backend, exchange, overlap=not args.no_overlap)
        step_fn = solver.step
        dofs = data.dofs
        cnt = torch.tensor([layout.n_own, layout.n_own, 0, 0, 0, layout.n_ghost], device=device, dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    upd, cells_total, n_acc, n_der, n_tay, n_ghost = (int(v) for v in cnt.tolist())

    for _ in range(warmup):
        step_fn()
    barrier()
    if args.lts > 1:
        be.launches = 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step_fn()
    e1.record(stream)
    barrier()
    clocks = sampler.finish()
    total_ms = e0.elapsed_time(e1)
    launches = be.launches if args.lts > 1 else backend.launches_per_step * steps
    finite = bool(torch.isfinite(dofs).all())

    e2e_steps = max(2, min(steps, 4))
    e2e_ms, io_bytes = e2e_staged(torch, stream, device, dofs, step_fn, e2e_steps, barrier)

    red = torch.tensor([total_ms, e2e_ms], device=device, dtype=torch.float64)
    dist.all_reduce(red, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = (float(v) for v in red.tolist())
    io = torch.tensor([io_bytes], device=device, dtype=torch.int64)
    dist.all_reduce(io, op=dist.ReduceOp.SUM)

    if rank == 0:
        ms = total_ms / steps
        value = upd / (ms * 1e-3)
        achieved = all_bytes * upd / world / (ms * 1e-3) / 1e9
        mode = (f"clustered LTS rate 2 with {args.lts} time clusters (bands in z), per-cluster NCCL halo exchange on a "
                "communication stream" if args.lts > 1 else "global time stepping, halo exchange overlapped with the interior update")
        line = {
            "metric": "element-updates/s", "value": value, "unit": "element-updates/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "gflops_nz": value * nzf * 1e-9,
            "parity_rel_l2": parity,
            "config": {"workload": f"cube mesh {nx}x{ny}x{nz} cubes x 5 tetrahedra = {cells_total} elements, elastic ORDER={order} "
                                   f"single precision, element update `all`, {mode}, x-slab partition over {world} GPUs "
                                   "(BASELINE.json configs[4])",
                       "name": "lts" if args.lts > 1 else "gts", "cells_per_gpu": cells_total // world,
                       "element_updates_per_step": upd, "ghost_cells": n_ghost,
                       "gts_updates_for_the_same_simulated_time": cells_total * 2 ** (args.lts - 1),
                       "accumulating_cells": n_acc, "cells_with_derivatives": n_der, "taylor_faces": n_tay,
                       "launches_per_step_per_gpu": launches // steps,
                       "parity": f"{parity_cells}-element cube mesh, partitioned over the {world} GPUs of this run vs one GPU, "
                                 f"after {2 if args.lts > 1 else 3} steps: relative L2 of the DOFs, max over ranks",
                       "l2": "inputs exceed L2", "parallelism": f"mesh partition x{world}", "finite": finite},
            "roofline": {"bound": "hbm", "kernel": "whole step (all launches of the step)", "achieved": achieved,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": None,
                         "peak_kind": peak_kind, "per": "GPU",
                         "note": "algorithmic bytes of `all` x element updates of the step / step time, per GPU"},
            "e2e": {"value": upd * e2e_steps / (e2e_ms * 1e-3), "unit": "element-updates/s",
                    "h2d_bytes_per_step": int(io.item()), "d2h_bytes_per_step": int(io.item()), "steps": e2e_steps,
                    "what": "per step and GPU: DOFs pinned host -> device staging, partitioned step through the C ABI (incl. halo "
                            "exchange), DOFs -> device staging -> pinned host; two staging slots per direction, copies overlap the kernels"},
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0


# ======================================================================================================================
def run_reference(args, cfg_name):
    """`--impl reference`: rank 0 alone times the CPU reference on the configuration of our arm; other ranks exit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from oracle import ref_lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    partitioned = world > 1 and not args.replicas
    if partitioned:
        return run_reference_mesh(args, world, steps, warmup)
    cfg = dict(CONFIGS[cfg_name])
    if args.cells:
        cfg["cells"] = args.cells
    if not ref_lib.available(cfg["ref"]):
        print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref/libssref_{cfg['ref']}.so not built"}))
        return 0
    # the full configuration when the whole run fits ~4 minutes, else a bounded sample of it per step
    probe = cpu_reference(cfg, 3.0, steps=1, warmup=0)
    rate = probe["value"]
    full = cfg["cells"]
    fits = full * (steps + warmup) / rate <= 240.0
    cells = full if fits else int(max(1024, rate * 240.0 / (steps + warmup)))
    if cfg["kind"] == "dr":
        cells = min(full, 2500 if cfg["law"] == "lsw" else 60)
    r = cpu_reference(cfg, 0.0, full_cells=cells, steps=steps, warmup=warmup)
    from seissol_b200 import proxy
    nz = {"elastic": lambda: proxy.nz_flops_per_element(cfg["order"], "all"),
          "visco": lambda: proxy.visco_nz_flops_per_element("all"), "dr": lambda: 0.0}[cfg["kind"]]()
    line = {
        "impl": "reference", "metric": "element-updates/s", "value": r["value"], "unit": "element-updates/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": warmup, "ms_per_step": 1e3 * r["seconds"] / r["steps"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if cfg["real"] == 4 else "f64", "data": "synthetic", "gflops_nz": r["value"] * nz * 1e-9,
        "config": {"workload": make_name(cfg, cfg_name), "name": cfg_name, "cells_per_gpu": cfg["cells"], "sample_cells": r["cells"],
                   "same_config": r["cells"] == cfg["cells"]},
        "cpu_baseline": {"value": r["value"], "unit": "element-updates/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "element-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def make_name(cfg, cfg_name):
    if cfg["kind"] == "elastic":
        return (f"proxy elastic ORDER={cfg['order']} {'single' if cfg['real'] == 4 else 'double'} precision kernel=all, "
                f"{cfg['cells']} elements (BASELINE.json {cfg['baseline']})")
    if cfg["kind"] == "visco":
        return f"proxy viscoelastic2 (3 mechanisms) ORDER=5 kernel=all, {cfg['cells']} elements (BASELINE.json {cfg['baseline']})"
    return f"proxy all_dr + friction {cfg['law']}, {4 * cfg['cells']} fault faces (BASELINE.json {cfg['baseline']})"


def run_reference_mesh(args, world, steps, warmup):
    """Reference arm of the partitioned configurations: the reference's generated kernels (`all` = ader + local + neigh
    per element) on a cube mesh with the same connectivity conventions as our arm's mesh -- a bounded sample (a smaller
    periodic box) of the ~2.5M-elements-per-GPU workload.  The CPU kernels have no LTS scheduler here: the metric is
    element updates per second, which is what an LTS step is made of."""
    from oracle import aderdg as O
    from oracle import ref_lib
    from seissol_b200 import mesh as M
    from seissol_b200 import proxy
    cfg = "elastic_o6_f32"
    if not ref_lib.available(cfg):
        print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref/libssref_{cfg}.so not built"}))
        return 0
    R = ref_lib.RefLib(cfg)
    cores = _ref_threads(R)
    dt = np.float32
    coeffs = O.time_integrate_coeffs(6, 0.0, 1.0, dt)
    cal = O.proxy_data(6, 256 * cores, dt, seed=3)
    R.proxy("all", cal, coeffs, 1)
    t, _, _ = R.proxy("all", cal, coeffs, 1)
    rate = cal["Q"].shape[0] / max(t, 1e-9)
    target = int(max(5000, min((args.cells or LTS_CELLS_PER_GPU) * world, rate * 240.0 / (steps + warmup))))
    n = max(4, int(round((target / 5.0) ** (1.0 / 3.0))))
    n += n % 2
    n = max(n, 4)
    cm = M.CubeMesh(n, n, n, periodic_yz=True)
    cells = cm.n
    d = O.proxy_data(6, cells, dt, seed=4, scale=2e-3)
    d["neighbors"] = np.ascontiguousarray(np.maximum(cm.neighbors, 0), dtype=np.int64)
    d["face_rel"] = np.ascontiguousarray(cm.face_rel, dtype=np.uint8)
    d["face_types"] = np.ascontiguousarray(np.where(cm.neighbors < 0, 5, 0), dtype=np.uint8)
    if warmup:
        R.proxy("all", d, coeffs, warmup)
    t, _, _ = R.proxy("all", d, coeffs, steps)
    value = cells * steps / t
    sample = (f"{n}x{n}x{n} cubes x 5 tetrahedra = {cells} elements (periodic box, same connectivity conventions), `all` with global "
              f"time stepping, {steps} passes")
    line = {
        "impl": "reference", "metric": "element-updates/s", "value": value, "unit": "element-updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "gflops_nz": value * proxy.nz_flops_per_element(6, "all") * 1e-9,
        "config": {"workload": "cube mesh, elastic ORDER=6 single precision, element update `all` (BASELINE.json configs[4])",
                   "name": "lts" if args.lts > 1 else "gts", "sample_cells": cells, "same_config": False},
        "cpu_baseline": {"value": value, "unit": "element-updates/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "element-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS) + ["lts", "gts"],
                    help="default: elastic_f32 on one GPU, lts on several")
    ap.add_argument("--cells", type=int, default=0, help="elements (per GPU); default: the configuration's size")
    ap.add_argument("--real", type=int, default=4, choices=[4, 8], help="shortcut: --real 8 = --config elastic_f64")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-overlap", action="store_true", help="gts: do not overlap the halo exchange")
    ap.add_argument("--lts", type=int, default=3, help="N > 1: number of rate-2 time clusters (1 = global time stepping)")
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent replicas of a single-GPU configuration")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = args.config
    if cfg == "gts":
        args.lts = 1
    if cfg in (None, "lts", "gts"):
        single = "elastic_f64" if args.real == 8 else "elastic_f32"
    else:
        single = cfg
    partitioned = world > 1 and not args.replicas and cfg in (None, "lts", "gts")
    if args.impl == "reference":
        if world > 1 and not partitioned:
            args.replicas = True
        return run_reference(args, single)
    if partitioned:
        return run_partitioned(args)
    return run_single(args, single)


if __name__ == "__main__":
    sys.exit(main())
