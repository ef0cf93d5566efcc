"""Host-side mirror of the reference proxy (src/Proxy, app/Proxy) on top of the C ABI.

Same kernel names (`all, local, neigh, ader, localwoader, ...`; src/Proxy/Common.cpp:33-40),
same composition (`all = local ∘ neigh`, `local = ader ∘ localwoader` fused;
src/Proxy/KernelHost.h:29-56), same timing protocol (one untimed warm-up pass, then
`timesteps` timed passes; src/Proxy/Runner.cpp:93-111) and the same output fields
(src/Proxy/Common.cpp:91-151).  Data lives in torch CUDA tensors in the device layout
described in DESIGN.md §2; all arithmetic happens in libseissol_b200.so.
"""
import math
import os
import time

import numpy as np

from . import api

KERNEL_NAMES = ["all", "local", "neigh", "ader", "localwoader", "neigh_dr", "godunov_dr", "all_dr"]
TIMESTEP = 1.0  # src/Proxy/Constants.h

# Static non-zero / hardware flop counts per element, read from the reference's generated
# kernels (`kernel::X::NonZeroFlops/HardwareFlops`, Monitoring/Metric.h:54-59) with
# oracle/ref_lib.py::RefLib.flops(); neighboringFlux is the mean over the 48 relations
# (uniform (h,j) as drawn by the proxy, src/Proxy/Allocator.cpp:67-87).
FLOPS = {
    4: {"derivative": (6642, 34146), "volume": (4806, 19584), "localFlux": (18972, None),
        "neighboringFlux": (4 * 5073.0, None)},
    5: {"derivative": (20943, 98001), "volume": (13986, 56448), "localFlux": (40320, 102528),
        "neighboringFlux": (4 * 10671.0, 4 * 29952)},
    6: {"derivative": (56286, 248292), "volume": (34839, 127872), "localFlux": (79200, 246528),
        "neighboringFlux": (4 * 20805.0, 4 * 73728)},
}


def nbasis(order):
    return order * (order + 1) * (order + 2) // 6


def integrate_coeffs(order, t0, t1):
    """MonomialBasis::integrate (src/Numerical/TimeBasis.h:82-96)."""
    return np.array([(t1 ** (d + 1) - t0 ** (d + 1)) / math.factorial(d + 1) for d in range(order)])


def nz_flops_per_element(order, kernel):
    f = FLOPS[order]
    ader = f["derivative"][0]
    lwo = f["volume"][0] + f["localFlux"][0]
    neigh = f["neighboringFlux"][0]
    return {"ader": ader, "localwoader": lwo, "local": ader + lwo, "neigh": neigh,
            "all": ader + lwo + neigh}[kernel]


def algorithmic_bytes_per_element(order, real_size, kernel):
    """unpadded reals x sizeof(real); constants and indices excluded."""
    bq = nbasis(order) * 9
    s, a = 72, 81
    reals = {"ader": 2 * bq + s, "localwoader": 3 * bq + s + 4 * a, "local": 3 * bq + s + 4 * a,
             "neigh": 6 * bq + 4 * a, "all": 9 * bq + s + 8 * a}[kernel]
    return reals * real_size


class ProxyData:
    """Fake single-layer storage, as ProxyData/fakeData (src/Proxy/Allocator.cpp:49-134)
    but with bounded values.  Device layout: dofs/buffers [n, 9, ld]
    column-major per element; local_int [n, 396] = star[3][24] | nApNm1[4][81];
    neigh_int [n, 324] = nAmNm1[4][81]."""

    def __init__(self, ctx, cells, device="cuda:0"):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.cells = cells
        self.device = torch.device(device)
        self.dtype = torch.float32 if ctx.real_size == 4 else torch.float64
        self.np_dtype = np.float32 if ctx.real_size == 4 else np.float64
        self.perm = None

    # -- construction ---------------------------------------------------------------
    @classmethod
    def from_host(cls, ctx, d, device="cuda:0"):
        """d: dict of numpy arrays in the oracle's logical layout (oracle/aderdg.py::proxy_data)."""
        self = cls(ctx, d["Q"].shape[0], device)
        t = self.torch
        self.dofs = self.to_device_layout(d["Q"])
        self.buffers = (self.to_device_layout(d["I"]) if "I" in d
                        else t.zeros_like(self.dofs))
        li = np.concatenate([d["star"].reshape(self.cells, 72), d["Aplus"].reshape(self.cells, 324)], axis=1)
        self.local_int = t.from_numpy(np.ascontiguousarray(li, dtype=self.np_dtype)).to(self.device)
        self.neigh_int = t.from_numpy(np.ascontiguousarray(d["Aminus"].reshape(self.cells, 324),
                                                           dtype=self.np_dtype)).to(self.device)
        self.face_rel = t.from_numpy(np.ascontiguousarray(d["face_rel"], dtype=np.uint8)).to(self.device)
        self.face_types = t.from_numpy(np.ascontiguousarray(d["face_types"], dtype=np.uint8)).to(self.device)
        self.neighbors = t.from_numpy(np.ascontiguousarray(d["neighbors"], dtype=np.int64)).to(self.device)
        return self

    @classmethod
    def synthetic(cls, ctx, cells, device="cuda:0", seed=42, scale=1e-2):
        """Bounded synthetic data generated on the device (large problem sizes)."""
        self = cls(ctx, cells, device)
        t = self.torch
        g = t.Generator(device=self.device)
        g.manual_seed(seed)
        ld, B = ctx.ld, ctx.B
        q = t.rand((cells, 9, ld), generator=g, device=self.device, dtype=self.dtype) * 2 - 1
        if ld > B:
            q[:, :, B:] = 0
        self.dofs = q.reshape(cells, 9 * ld).contiguous()
        self.buffers = t.zeros_like(self.dofs)
        self.local_int = ((t.rand((cells, 396), generator=g, device=self.device, dtype=self.dtype) * 2 - 1)
                          * scale).contiguous()
        self.neigh_int = ((t.rand((cells, 324), generator=g, device=self.device, dtype=self.dtype) * 2 - 1)
                          * scale).contiguous()
        rel = t.empty((cells, 4, 2), dtype=t.uint8, device=self.device)
        rel[:, :, 0] = t.randint(0, 4, (cells, 4), generator=g, device=self.device).to(t.uint8)
        rel[:, :, 1] = t.randint(0, 3, (cells, 4), generator=g, device=self.device).to(t.uint8)
        self.face_rel = rel.contiguous()
        self.face_types = t.zeros((cells, 4), dtype=t.uint8, device=self.device)
        self.neighbors = t.randint(0, cells, (cells, 4), generator=g, device=self.device, dtype=t.int64)
        return self

    # -- layouts ----------------------------------------------------------------------
    def to_device_layout(self, x):
        n, b, q = x.shape
        out = np.zeros((n, q, self.ctx.ld), dtype=self.np_dtype)
        out[:, :, :b] = x.transpose(0, 2, 1)
        return self.torch.from_numpy(out.reshape(n, q * self.ctx.ld)).to(self.device)

    def to_logical(self, t):
        x = t.detach().cpu().numpy().reshape(t.shape[0], 9, self.ctx.ld)
        return np.ascontiguousarray(x[:, :, :self.ctx.B].transpose(0, 2, 1))

    def sort_by_face_relation(self, coarse=False):
        """Processing order that makes (neighbor's face id j, orientation h) of all four faces
        uniform inside a warp and inside the element pairs that share packed FFMA2
        instructions (12^4 classes), removing divergence in the rT_j / fP_h selection of k_neigh.
        This is our analogue of the reference's recorder grouping by ConditionalKey
        (src/Initializer/BatchRecorders/NeighIntegrationRecorder.cpp:142-144)."""
        t = self.torch
        j = self.face_rel[:, :, 0].to(t.int64)
        h = self.face_rel[:, :, 1].to(t.int64)
        # j (selects the large rT operators) varies slowest: all CTAs resident at one time execute
        # the same instruction stream, which the instruction caches reward;
        # h (small fP operators) is sorted inside a j class.
        key = (j[:, 0] + 4 * j[:, 1] + 16 * j[:, 2] + 64 * j[:, 3]) * 81
        if not coarse:
            key = key + h[:, 0] + 3 * h[:, 1] + 9 * h[:, 2] + 27 * h[:, 3]
        self.perm = t.argsort(key, stable=True).to(t.int32).contiguous()
        return self.perm

    def bytes_resident(self):
        return sum(x.numel() * x.element_size() for x in
                   (self.dofs, self.buffers, self.local_int, self.neigh_int, self.face_rel,
                    self.face_types, self.neighbors) + ((self.face_proj,) if getattr(self, "face_proj", None) is not None else ()))


class ProxyKernels:
    """Kernel-name dispatch (getProxyKernelDevice, src/Proxy/KernelDevice.cpp:126-150)."""

    def __init__(self, ctx, data, faceproj=None):
        """faceproj: route `localwoader`/`local` + `neigh` through the face-projection path (ssb_local_faceproj +
        ssb_neigh_faceproj: the local kernel stores F_f = fMrT_f I per face, the neighbor kernel gathers those blocks
        instead of the neighbors' I).  Default: on (environment SSB_FACEPROJ=0 switches back to ssb_local + ssb_neigh)."""
        self.ctx, self.d = ctx, data
        self.coeffs = integrate_coeffs(ctx.order, 0.0, TIMESTEP)
        if faceproj is None:
            # pays where the rT_j products dominate the neighbor kernel: order 6 (all 3.21 -> 3.82 ms at order 4, equal
            # at order 5, 9.1 -> 8.2 ms at order 6; profiles/r2_quick_bench_orders.txt)
            faceproj = os.environ.get("SSB_FACEPROJ", "1" if ctx.order >= 6 else "0") != "0"
        self.faceproj = bool(faceproj)
        if self.faceproj and getattr(data, "face_proj", None) is None:
            data.face_proj = data.torch.zeros((data.cells, ctx.face_proj_reals()), dtype=data.dtype, device=data.device)

    def ader(self, stream=None):
        d = self.d
        self.ctx.ader(api.strided(d.dofs), api.strided(d.buffers), None, api.strided(d.local_int),
                      self.coeffs, d.cells, stream=stream)

    def localwoader(self, stream=None):
        d = self.d
        if self.faceproj:
            self.ctx.local_faceproj(api.strided(d.dofs), api.strided(d.buffers), api.strided(d.local_int),
                                    api.strided(d.face_proj), d.face_types, d.cells, stream=stream)
            return
        self.ctx.local(api.strided(d.dofs), api.strided(d.buffers), api.strided(d.local_int),
                       d.face_types, d.cells, stream=stream)

    def local(self, stream=None, fused=False):
        """ader ∘ localwoader.  Two launches by default: the fused kernel (`ssb_ader_local`, I never
        leaves the SM) executes 12k straight-line instructions per thread and is instruction-fetch
        bound; the two smaller kernels run 35 % faster although I makes a round trip through L2/HBM
        (profiles/)."""
        d = self.d
        if fused and self.faceproj:
            self.ctx.ader_local_faceproj(api.strided(d.dofs), api.strided(d.buffers), None, api.strided(d.local_int),
                                         api.strided(d.face_proj), d.face_types, self.coeffs, d.cells, stream=stream)
        elif fused:
            self.ctx.ader_local(api.strided(d.dofs), api.strided(d.buffers), None, api.strided(d.local_int),
                                d.face_types, self.coeffs, d.cells, stream=stream)
        else:
            self.ader(stream)
            self.localwoader(stream)

    def neigh(self, stream=None):
        d = self.d
        if self.faceproj:
            self.ctx.neigh_faceproj(api.strided(d.dofs), api.strided(d.face_proj), api.strided(d.neigh_int),
                                    d.face_types, d.face_rel, d.cells, nb_idx=d.neighbors, perm=d.perm, stream=stream)
            return
        self.ctx.neigh(api.strided(d.dofs), api.strided(d.buffers), api.strided(d.neigh_int),
                       d.face_types, d.face_rel, d.cells, nb_idx=d.neighbors, perm=d.perm, stream=stream)

    def all(self, stream=None):
        self.local(stream)
        self.neigh(stream)

    def run(self, name, stream=None):
        for part in name.split(","):
            getattr(self, part)(stream)

    def launches(self, name):
        return sum({"all": 3, "local": 2, "neigh": 1, "ader": 1, "localwoader": 1}[p] for p in name.split(","))


def run_proxy(order, real_size, cells, timesteps, kernel, device="cuda:0", data=None, sort=True):
    """runProxy (src/Proxy/Runner.cpp:57-151): returns the reference's output fields plus
    element-updates/s."""
    import torch
    ctx = api.Context(order, real_size)
    d = data if data is not None else ProxyData.synthetic(ctx, cells, device)
    if sort and d.perm is None:
        d.sort_by_face_relation()
    k = ProxyKernels(ctx, d)
    k.run(kernel)  # warm-up pass (Runner.cpp:93-95)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(timesteps):
        k.run(kernel)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    nz = sum(nz_flops_per_element(order, p) for p in kernel.split(",")) * cells * timesteps
    by = sum(algorithmic_bytes_per_element(order, real_size, p) for p in kernel.split(",")) * cells * timesteps
    return {
        "name": kernel, "time": dt, "cycles": 0.0,
        "gflop-nz": nz * 1e-9, "gib": by / 2 ** 30,
        "gflops-nz": nz * 1e-9 / dt, "gibs": by / 2 ** 30 / dt,
        "elem-updates-per-s": cells * timesteps / dt,
    }


# ------------------------------------------------------------------------------------------------
# Dynamic rupture (proxy kernels godunov_dr, neigh_dr, all_dr; src/Proxy/Allocator.cpp:193-294)
def gauss_legendre_points(order, dt):
    """Gauss-Legendre points/weights on [0, dt] (src/Numerical/Quadrature.h:76-90)."""
    x, w = np.polynomial.legendre.leggauss(order)
    return 0.5 * dt * (x + 1.0), 0.5 * dt * w


def point_coeffs(order, tau):
    """MonomialBasis::point (src/Numerical/TimeBasis.h:71-80): tau^d / d!."""
    return np.array([tau ** d / math.factorial(d) for d in range(order)])


def ders_to_device(ctx, ders, np_dtype):
    """list over d of [n, B, 9] -> [n, dersize] in the context's derivative-family layout."""
    n = ders[0].shape[0]
    out = np.zeros((n, ctx.dersize), dtype=np_dtype)
    for d, dq in enumerate(ders):
        rows = nbasis(ctx.order - d)
        blk = np.zeros((n, 9, ctx.der_ld[d]), dtype=np_dtype)
        blk[:, :, :rows] = dq[:, :rows, :].transpose(0, 2, 1)
        out[:, ctx.der_offset[d]:ctx.der_offset[d] + 9 * ctx.der_ld[d]] = blk.reshape(n, -1)
    return out


def points_to_device(ctx, x):
    """[..., P, 9] -> [..., 9, ppad] (point index fastest, zero padding)."""
    out = np.zeros(x.shape[:-2] + (9, ctx.ppad), dtype=x.dtype)
    out[..., :x.shape[-2]] = np.swapaxes(x, -1, -2)
    return out


def points_from_device(ctx, x, P):
    """[..., 9, ppad] -> [..., P, 9]."""
    return np.ascontiguousarray(np.swapaxes(x[..., :P], -1, -2))


def friction_params(order, dt, full_update_time=0.0, **kw):
    """ssb_friction_params_t for one time step: Gauss-Legendre sub-steps (FrictionSolver.cpp:20-37) and
    time weights normalised by dt as TimeCluster::computeDynamicRupture does."""
    pts, wts = gauss_legendre_points(order, dt)
    delta = np.empty(order)
    delta[0] = pts[0]
    delta[1:] = np.diff(pts)
    delta[-1] += pts[0]
    p = api.SsbFrictionParams()
    for i in range(order):
        p.delta_t[i] = delta[i]
        p.time_weights[i] = wts[i] / dt
    p.full_update_time = full_update_time
    p.eta_p_damp = kw.get("eta_p_damp", 1.0)
    p.healing_threshold = kw.get("healing_threshold", -1.0)
    p.t0 = kw.get("t0", 0.0)
    p.rs_sr0 = kw.get("rs_sr0", 1e-6)
    p.rs_slip_rate_tolerance = kw.get("rs_slip_rate_tolerance", 1e-8)
    p.rs_state_tolerance = kw.get("rs_state_tolerance", 1e-8)
    p.rs_max_number_slip_rate_updates = kw.get("rs_max_number_slip_rate_updates", 60)
    p.rs_number_state_variable_updates = kw.get("rs_number_state_variable_updates", 10)
    nuc_t0, nuc_s0 = kw.get("nucleation_t0", []), kw.get("nucleation_s0", [])
    p.energies_from_across_fault_velocities = 1 if kw.get("energies_from_across_fault_velocities", False) else 0
    p.rs_float_resolution_exit = 1 if kw.get("rs_float_resolution_exit", False) else 0
    p.nucleation_count = len(nuc_t0)
    for i in range(len(nuc_t0)):
        p.nucleation_t0[i], p.nucleation_s0[i] = float(nuc_t0[i]), float(nuc_s0[i])
    return p, delta, wts / dt


class DrProxyData:
    """Fault-side fake data of the proxy: 4 fault faces per cell, each picking random plus/minus cells'
    derivative families; every cell face maps to a random fault face (src/Proxy/Allocator.cpp:230-294)."""

    def __init__(self, ctx, data, seed=7, scale=1e-2):
        import torch
        t = torch
        self.ctx, self.data = ctx, data
        dev, dt = data.device, data.dtype
        cells = data.cells
        self.faces = 4 * cells
        g = t.Generator(device=dev)
        g.manual_seed(seed)
        self.derivs = t.rand((cells, ctx.dersize), generator=g, device=dev, dtype=dt)
        self.plus_cell = t.randint(0, cells, (self.faces,), generator=g, device=dev, dtype=t.int64)
        self.minus_cell = t.randint(0, cells, (self.faces,), generator=g, device=dev, dtype=t.int64)
        info = t.empty((self.faces, 3), dtype=t.uint8, device=dev)
        info[:, 0] = t.randint(0, 4, (self.faces,), generator=g, device=dev).to(t.uint8)
        info[:, 1] = t.randint(0, 4, (self.faces,), generator=g, device=dev).to(t.uint8)
        info[:, 2] = t.randint(0, 3, (self.faces,), generator=g, device=dev).to(t.uint8)
        self.face_info = info.contiguous()
        self.tinv_t = (t.rand((self.faces, 81), generator=g, device=dev, dtype=dt) * 2 - 1).contiguous()
        self.q_plus = t.zeros((self.faces, ctx.order, 9, ctx.ppad), device=dev, dtype=dt)
        self.q_minus = t.zeros_like(self.q_plus)
        self.imposed_plus = ((t.rand((self.faces, 9, ctx.ppad), generator=g, device=dev, dtype=dt) * 2 - 1) * scale)
        self.flux_solver = ((t.rand((self.faces, 81), generator=g, device=dev, dtype=dt) * 2 - 1)).contiguous()
        # CellDRMapping per (cell, face)
        sr = t.empty((cells, 4, 2), dtype=t.uint8, device=dev)
        sr[:, :, 0] = t.randint(0, 4, (cells, 4), generator=g, device=dev).to(t.uint8)
        sr[:, :, 1] = t.randint(0, 3, (cells, 4), generator=g, device=dev).to(t.uint8)
        self.side_rel = sr.contiguous()
        self.dr_face = t.randint(0, self.faces, (cells, 4), generator=g, device=dev, dtype=t.int64)
        es = self.imposed_plus.element_size()
        self.godunov_ptrs = (self.imposed_plus.data_ptr() + self.dr_face * (9 * ctx.ppad * es)).contiguous()
        self.flux_solver_ptrs = (self.flux_solver.data_ptr() + self.dr_face * (81 * es)).contiguous()
        self.dr_face_types = t.full((cells, 4), 3, dtype=t.uint8, device=dev)
        # fault faces sorted by their (plus, minus) operator variants: uniform instruction stream per CTA
        key = info[:, 0].to(t.int64) + 4 * (info[:, 1].to(t.int64) + 4 * info[:, 2].to(t.int64))
        self.perm = t.argsort(key, stable=True).to(t.int32).contiguous()
        pts, _ = gauss_legendre_points(ctx.order, TIMESTEP)
        self.coeffs = np.stack([point_coeffs(ctx.order, tau) for tau in pts])

    def der_array(self, cell_index):
        stride = self.derivs.stride(0) * self.derivs.element_size()
        self._keep = getattr(self, "_keep", [])
        ptrs = (self.derivs.data_ptr() + cell_index * stride).contiguous()
        self._keep.append(ptrs)
        return api.pointer_table(ptrs)


class DrProxyKernels(ProxyKernels):
    """godunov_dr, neigh_dr, all_dr (src/Proxy/KernelHost.h:29-56)."""

    def __init__(self, ctx, data, dr):
        super().__init__(ctx, data, faceproj=False)   # all faces are DynamicRupture: no regular neighbor flux
        self.dr = dr
        self._plus = dr.der_array(dr.plus_cell)
        self._minus = dr.der_array(dr.minus_cell)

    def godunov_dr(self, stream=None):
        r = self.dr
        self.ctx.dr_interp(self._plus, self._minus, r.tinv_t, r.face_info, r.q_plus, r.q_minus, r.coeffs, r.faces,
                           perm=r.perm, stream=stream)

    def neigh_dr(self, stream=None):
        r, d = self.dr, self.d
        self.ctx.neigh_dr(api.strided(d.dofs), r.godunov_ptrs, r.flux_solver_ptrs, r.side_rel, r.dr_face_types,
                          d.cells, stream=stream)

    def all_dr(self, stream=None):
        self.local(stream)
        self.godunov_dr(stream)
        self.neigh_dr(stream)


# ------------------------------------------------------------------------------------------------
# viscoelastic2 (anelastic attenuation, 3 mechanisms, order 5; BASELINE.json configs[2])
VISCO_FLOPS = {"derivative": 120489, "volumeExt": 15924, "localFluxExt": 12708 + 12780 + 22068 + 20124, "local": 14175,
               "neighborFluxExt": 4 * 17511.0, "neighbor": 1575}   # kernel::X::NonZeroFlops of the generated kernels


def visco_nz_flops_per_element(kernel):
    f = VISCO_FLOPS
    ader = f["derivative"]
    lwo = f["volumeExt"] + f["localFluxExt"] + f["local"]
    neigh = f["neighborFluxExt"] + f["neighbor"]
    return {"ader": ader, "localwoader": lwo, "local": ader + lwo, "neigh": neigh, "all": ader + lwo + neigh}[kernel]


def visco_algorithmic_bytes_per_element(real_size, kernel, mech=3):
    """SURVEY.md §8(d), unpadded reals: Q 315, Qane 630, star 99, E/w/W 168, A 4 x 135."""
    q, qa, s, ew, a = 315, 210 * mech, 99, 54 * mech + 2 * mech, 135
    reals = {"ader": 2 * q + 2 * qa + s + ew, "localwoader": 3 * q + 3 * qa + s + ew + 4 * a,
             "local": 3 * q + 3 * qa + s + ew + 4 * a, "neigh": 4 * q + 2 * (q + qa) + 4 * a + mech}
    reals["all"] = reals["ader"] + reals["localwoader"] + reals["neigh"]
    return reals[kernel] * real_size


class ViscoProxyData:
    """Fake single-layer storage of the viscoelastic2 proxy.  Device layout: dofs/buffers [n, 9*ld]; dofs_ane/buffers_ane
    [n, 18*ld] (column p + 6 m); local_int [n, 807] = star[3][33] | nApNm1[4][135] | E[162] | w[3] | W[3];
    neigh_int [n, 543] = nAmNm1[4][135] | w[3]."""

    MECH = 3

    def __init__(self, ctx, cells, device="cuda:0"):
        import torch
        self.torch, self.ctx, self.cells = torch, ctx, cells
        self.device = torch.device(device)
        self.dtype = torch.float32 if ctx.real_size == 4 else torch.float64
        self.np_dtype = np.float32 if ctx.real_size == 4 else np.float64
        self.perm = None

    @classmethod
    def from_host(cls, ctx, d, device="cuda:0"):
        """d: dict in the logical layout of oracle/visco.py::proxy_data."""
        self = cls(ctx, d["Q"].shape[0], device)
        t, n = self.torch, self.cells
        self.dofs = self.q_to_device(d["Q"])
        self.dofs_ane = self.ane_to_device(d["Qane"])
        self.buffers = self.q_to_device(d["I"]) if "I" in d else t.zeros_like(self.dofs)
        self.buffers_ane = self.ane_to_device(d["Iane"]) if "Iane" in d else t.zeros_like(self.dofs_ane)
        li = np.concatenate([d["star"].reshape(n, 99), d["Aplus"].reshape(n, 540), d["E"].reshape(n, 162),
                             d["w"].reshape(n, 3), d["W"].reshape(n, 3)], axis=1)
        ni = np.concatenate([d["Aminus"].reshape(n, 540), d["w"].reshape(n, 3)], axis=1)
        self.local_int = t.from_numpy(np.ascontiguousarray(li, dtype=self.np_dtype)).to(self.device)
        self.neigh_int = t.from_numpy(np.ascontiguousarray(ni, dtype=self.np_dtype)).to(self.device)
        self.face_rel = t.from_numpy(np.ascontiguousarray(d["face_rel"], dtype=np.uint8)).to(self.device)
        self.face_types = t.from_numpy(np.ascontiguousarray(d["face_types"], dtype=np.uint8)).to(self.device)
        self.neighbors = t.from_numpy(np.ascontiguousarray(d["neighbors"], dtype=np.int64)).to(self.device)
        return self

    @classmethod
    def synthetic(cls, ctx, cells, device="cuda:0", seed=42, scale=1e-2):
        self = cls(ctx, cells, device)
        t = self.torch
        g = t.Generator(device=self.device)
        g.manual_seed(seed)
        ld, B = ctx.ld, ctx.B

        def modal(cols):
            x = t.rand((cells, cols, ld), generator=g, device=self.device, dtype=self.dtype) * 2 - 1
            if ld > B:
                x[:, :, B:] = 0
            return x.reshape(cells, cols * ld).contiguous()

        self.dofs, self.dofs_ane = modal(9), modal(18)
        self.buffers, self.buffers_ane = t.zeros_like(self.dofs), t.zeros_like(self.dofs_ane)
        self.local_int = ((t.rand((cells, 807), generator=g, device=self.device, dtype=self.dtype) * 2 - 1) * scale).contiguous()
        self.neigh_int = ((t.rand((cells, 543), generator=g, device=self.device, dtype=self.dtype) * 2 - 1) * scale).contiguous()
        rel = t.empty((cells, 4, 2), dtype=t.uint8, device=self.device)
        rel[:, :, 0] = t.randint(0, 4, (cells, 4), generator=g, device=self.device).to(t.uint8)
        rel[:, :, 1] = t.randint(0, 3, (cells, 4), generator=g, device=self.device).to(t.uint8)
        self.face_rel = rel.contiguous()
        self.face_types = t.zeros((cells, 4), dtype=t.uint8, device=self.device)
        self.neighbors = t.randint(0, cells, (cells, 4), generator=g, device=self.device, dtype=t.int64)
        return self

    def q_to_device(self, x):
        n, b, q = x.shape
        out = np.zeros((n, q, self.ctx.ld), dtype=self.np_dtype)
        out[:, :, :b] = x.transpose(0, 2, 1)
        return self.torch.from_numpy(out.reshape(n, q * self.ctx.ld)).to(self.device)

    def ane_to_device(self, x):
        n, b, p, m = x.shape
        out = np.zeros((n, m, p, self.ctx.ld), dtype=self.np_dtype)
        out[:, :, :, :b] = x.transpose(0, 3, 2, 1)
        return self.torch.from_numpy(out.reshape(n, m * p * self.ctx.ld)).to(self.device)

    def q_to_logical(self, t):
        x = t.detach().cpu().numpy().reshape(t.shape[0], 9, self.ctx.ld)
        return np.ascontiguousarray(x[:, :, :self.ctx.B].transpose(0, 2, 1))

    def ane_to_logical(self, t):
        x = t.detach().cpu().numpy().reshape(t.shape[0], self.MECH, 6, self.ctx.ld)
        return np.ascontiguousarray(x[:, :, :, :self.ctx.B].transpose(0, 3, 2, 1))

    def sort_by_face_relation(self):
        t = self.torch
        j = self.face_rel[:, :, 0].to(t.int64)
        h = self.face_rel[:, :, 1].to(t.int64)
        key = (j[:, 0] + 4 * j[:, 1] + 16 * j[:, 2] + 64 * j[:, 3]) * 81 + h[:, 0] + 3 * h[:, 1] + 9 * h[:, 2] + 27 * h[:, 3]
        self.perm = t.argsort(key, stable=True).to(t.int32).contiguous()
        return self.perm

    def bytes_resident(self):
        return sum(x.numel() * x.element_size() for x in
                   (self.dofs, self.dofs_ane, self.buffers, self.buffers_ane, self.local_int, self.neigh_int,
                    self.face_rel, self.face_types, self.neighbors))


class ViscoProxyKernels:
    """Proxy kernel names on the viscoelastic2 context (src/Proxy/KernelDevice.cpp with the LinearCKAnelastic kernels)."""

    def __init__(self, ctx, data):
        self.ctx, self.d = ctx, data
        self.coeffs = integrate_coeffs(ctx.order, 0.0, TIMESTEP)

    def ader(self, stream=None):
        d = self.d
        self.ctx.visco_ader(api.strided(d.dofs), api.strided(d.dofs_ane), api.strided(d.buffers), api.strided(d.buffers_ane),
                            api.strided(d.local_int), self.coeffs, d.cells, stream=stream)

    def localwoader(self, stream=None):
        d = self.d
        self.ctx.visco_local(api.strided(d.dofs), api.strided(d.dofs_ane), api.strided(d.buffers), api.strided(d.buffers_ane),
                             api.strided(d.local_int), d.face_types, d.cells, stream=stream)

    def local(self, stream=None):
        self.ader(stream)
        self.localwoader(stream)

    def neigh(self, stream=None):
        d = self.d
        self.ctx.visco_neigh(api.strided(d.dofs), api.strided(d.dofs_ane), api.strided(d.buffers), api.strided(d.neigh_int),
                             d.face_types, d.face_rel, d.cells, nb_idx=d.neighbors, perm=d.perm, stream=stream)

    def all(self, stream=None):
        self.local(stream)
        self.neigh(stream)

    def run(self, name, stream=None):
        for part in name.split(","):
            getattr(self, part)(stream)
