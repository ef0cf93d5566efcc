"""GPU parity tests (run with -m gpu): the sm_100a kernels, called through the C ABI, against
the CPU oracle on identical seeded inputs.  Tolerances are BASELINE.json's: relative L2 of
the DOFs <= 1e-5 (single) / 1e-11 (double) after N time steps; single kernels are held
to a tighter 100*eps-class bound."""
import os

import numpy as np
import pytest

from conftest import rel_l2
from oracle import aderdg as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CONFIGS = [(4, np.float32), (5, np.float32), (6, np.float32), (4, np.float64), (6, np.float64)]


def step_tol(dt):
    return 2e-6 if dt == np.float32 else 5e-15


def run_tol(dt):
    return 1e-5 if dt == np.float32 else 1e-11


@pytest.fixture(scope="module")
def gpu(built_library):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    from seissol_b200 import api, proxy
    return torch, api, proxy


def face_proj_host(tab, I):
    """F_f = fMrT_f . I for the four faces, logical layout [n, 4, B2, 9] (what ssb_local_faceproj stores)."""
    return np.stack([np.matmul(tab["fMrT"][f], I) for f in range(4)], 1)


def face_proj_to_device(torch, ctx, F, device):
    n, _, b2, _ = F.shape
    b2p = (b2 + 3) // 4 * 4
    out = np.zeros((n, 4, 9, b2p), dtype=F.dtype)
    out[:, :, :, :b2] = F.transpose(0, 1, 3, 2)
    return torch.from_numpy(out.reshape(n, -1)).to(device)


def face_proj_from_device(ctx, t):
    b2 = ctx.order * (ctx.order + 1) // 2
    b2p = (b2 + 3) // 4 * 4
    x = t.cpu().numpy().reshape(t.shape[0], 4, 9, b2p)
    return np.ascontiguousarray(x[:, :, :, :b2].transpose(0, 1, 3, 2)), x[:, :, :, b2:]


def make(order, dt, n, api, proxy, seed=5, reference_layout=False):
    ctx = api.Context(order, np.dtype(dt).itemsize, reference_layout=reference_layout)
    host = O.proxy_data(order, n, dt, seed=seed)
    data = proxy.ProxyData.from_host(ctx, host)
    return ctx, host, data, O.load_tables(order, dt), O.time_integrate_coeffs(order, 0.0, 1.0, dt)


@pytest.mark.parametrize("order,dt", CONFIGS)
def test_ader_with_derivatives(gpu, order, dt):
    torch, api, proxy = gpu
    n = 333  # ragged: not a multiple of the elements-per-CTA tile
    ctx, host, data, tab, c = make(order, dt, n, api, proxy)
    ders_dev = torch.zeros((n, ctx.dersize), dtype=data.dtype, device=data.device)
    ctx.ader(api.strided(data.dofs), api.strided(data.buffers), api.strided(ders_dev),
             api.strided(data.local_int), c, n)
    torch.cuda.synchronize()
    star = O.star_dense(host["star"], tab["star_pattern"])
    I, ders = O.ader(tab, host["Q"], star, c, True)
    assert rel_l2(data.to_logical(data.buffers), I) < step_tol(dt)
    assert rel_l2(data.to_logical(data.dofs), host["Q"]) == 0.0  # ader must not touch Q
    dd = ders_dev.cpu().numpy()
    for d in range(order):
        rows = O.nbasis(order - d)
        blk = dd[:, ctx.der_offset[d]:ctx.der_offset[d] + 9 * ctx.der_ld[d]].reshape(n, 9, ctx.der_ld[d])
        assert rel_l2(blk[:, :, :rows].transpose(0, 2, 1), ders[d][:, :rows]) < step_tol(dt)
    # Taylor sum of the stored family over a sub-interval (LTS neighbor path)
    c2 = O.time_integrate_coeffs(order, 0.25, 0.75, dt)
    out = torch.zeros_like(data.buffers)
    ctx.taylor(api.strided(ders_dev), api.strided(out), c2, n)
    torch.cuda.synchronize()
    assert rel_l2(data.to_logical(out), O.taylor(ders, c2)) < step_tol(dt)


@pytest.mark.parametrize("order,dt", CONFIGS)
@pytest.mark.parametrize("faceproj", [False, True])
def test_localwoader_and_fused_local(gpu, order, dt, faceproj):
    torch, api, proxy = gpu
    n = 200
    ctx, host, data, tab, c = make(order, dt, n, api, proxy, seed=9)
    host["face_types"][5, 2] = 3   # DynamicRupture faces get no local flux
    host["face_types"][17, 0] = 3
    data = proxy.ProxyData.from_host(ctx, host)
    star = O.star_dense(host["star"], tab["star_pattern"])
    I = O.ader(tab, host["Q"], star, c)
    want = O.local(tab, host["Q"], I, star, O.flux_dense(host["Aplus"]), host["face_types"])
    # ader + local (two launches; `fused` = one launch)
    for fused in (False, True):
        data = proxy.ProxyData.from_host(ctx, host)
        k = proxy.ProxyKernels(ctx, data, faceproj=faceproj)
        k.local(fused=fused)
        torch.cuda.synchronize()
        assert rel_l2(data.to_logical(data.buffers), I) < step_tol(dt)
        assert rel_l2(data.to_logical(data.dofs), want) < step_tol(dt)
        assert rel_l2(data.to_logical(data.dofs) - host["Q"], want - host["Q"]) < 50 * step_tol(dt)
        if faceproj:
            # the by-product of the local flux: F_f = fMrT_f . I for ALL faces (DR faces included), padding rows zero
            got, pad = face_proj_from_device(ctx, data.face_proj)
            assert rel_l2(got, face_proj_host(tab, I)) < step_tol(dt)
            assert np.all(pad == 0)
    # stand-alone localwoader on the oracle's I
    host2 = dict(host)
    host2["I"] = I
    data2 = proxy.ProxyData.from_host(ctx, host2)
    proxy.ProxyKernels(ctx, data2, faceproj=faceproj).localwoader()
    torch.cuda.synchronize()
    assert rel_l2(data2.to_logical(data2.dofs), want) < step_tol(dt)


@pytest.mark.parametrize("order,dt", CONFIGS)
@pytest.mark.parametrize("sort", [False, True])
@pytest.mark.parametrize("faceproj", [False, True])
def test_neighbor(gpu, order, dt, sort, faceproj):
    """faceproj: ssb_neigh_faceproj on the projections F_f = fMrT_f . I_nb (computed here with the oracle's tables)
    must give what the reference's rDivM_i fP_h rT_j I_nb AminusT_i gives."""
    torch, api, proxy = gpu
    n = 301
    ctx, host, data, tab, c = make(order, dt, n, api, proxy, seed=13)
    rng = np.random.default_rng(0)
    host["I"] = rng.uniform(-1, 1, host["Q"].shape).astype(dt)
    host["face_types"][3, 1] = 1   # free surface: no neighbor contribution
    host["face_types"][8, 3] = 5   # outflow
    data = proxy.ProxyData.from_host(ctx, host)
    if sort:
        data.sort_by_face_relation()
    if faceproj:
        data.face_proj = face_proj_to_device(torch, ctx, face_proj_host(tab, host["I"]), data.device)
    proxy.ProxyKernels(ctx, data, faceproj=faceproj).neigh()
    torch.cuda.synchronize()
    want = O.neighbor(tab, host["Q"], host["I"], host["neighbors"], O.flux_dense(host["Aminus"]),
                      host["face_rel"], host["face_types"])
    assert rel_l2(data.to_logical(data.dofs), want) < step_tol(dt)
    assert rel_l2(data.to_logical(data.dofs) - host["Q"], want - host["Q"]) < 50 * step_tol(dt)


@pytest.mark.parametrize("order,dt", [(6, np.float32), (6, np.float64)])
@pytest.mark.parametrize("faceproj", [False, True])
def test_neighbor_boundary_faces_ignore_stale_shared_memory(gpu, order, dt, faceproj):
    """Faces without a regular neighbor load no tile: whatever the previous kernel left in shared memory (here: NaN
    patterns pushed through torch scan/sort kernels) must not leak into the update (regression: NaN * 0)."""
    torch, api, proxy = gpu
    n = 4000
    ctx, host, data, tab, c = make(order, dt, n, api, proxy, seed=21)
    rng = np.random.default_rng(1)
    host["I"] = rng.uniform(-1, 1, host["Q"].shape).astype(dt)
    host["face_types"][rng.uniform(size=(n, 4)) < 0.3] = 5        # many boundary (outflow) faces, every CTA has some
    host["face_types"][::2, 0] = 5
    data = proxy.ProxyData.from_host(ctx, host)
    data.sort_by_face_relation()
    for _ in range(3):
        poison = torch.full((1 << 22,), float("nan"), device=data.device, dtype=torch.float32)
        poison.cumsum(0)
        torch.sort(poison)
        torch.sort(torch.full((1 << 22,), -1, device=data.device, dtype=torch.int32))    # 0xffffffff = NaN pattern
    if faceproj:
        data.face_proj = face_proj_to_device(torch, ctx, face_proj_host(tab, host["I"]), data.device)
    proxy.ProxyKernels(ctx, data, faceproj=faceproj).neigh()
    torch.cuda.synchronize()
    got = data.to_logical(data.dofs)
    assert np.isfinite(got).all()
    want = O.neighbor(tab, host["Q"], host["I"], host["neighbors"], O.flux_dense(host["Aminus"]),
                      host["face_rel"], host["face_types"])
    assert rel_l2(got, want) < step_tol(dt)


@pytest.mark.parametrize("order,dt,steps", [(6, np.float32, 10), (6, np.float64, 10), (4, np.float32, 10),
                                            (5, np.float32, 10), (6, np.float32, 50), (6, np.float64, 50)])
@pytest.mark.parametrize("faceproj", [False, True])
def test_all_n_steps(gpu, order, dt, steps, faceproj):
    """BASELINE.json parity bar: DOFs after N = 10 and N = 50 steps of `all` within 1e-5 (f32) / 1e-11 (f64), on
    >= 10 k elements for N = 10, through both neighbor paths."""
    torch, api, proxy = gpu
    n = 10007 if steps == 10 else 600
    ctx, host, data, tab, c = make(order, dt, n, api, proxy, seed=21)
    data.sort_by_face_relation()
    k = proxy.ProxyKernels(ctx, data, faceproj=faceproj)
    for _ in range(steps):
        k.all()
        host = O.proxy_step(tab, host, c, "all")
    torch.cuda.synchronize()
    err = rel_l2(data.to_logical(data.dofs), host["Q"])
    assert np.isfinite(err) and err < run_tol(dt), err
    assert rel_l2(data.to_logical(data.buffers), host["I"]) < run_tol(dt)


@pytest.mark.parametrize("cfg,order,dt", [("elastic_o4_f32", 4, np.float32), ("elastic_o6_f32", 6, np.float32),
                                          ("elastic_o6_f64", 6, np.float64)])
def test_against_reference_golden_vectors(gpu, cfg, order, dt):
    """Outputs of the REFERENCE's generated kernels (tests/golden) on the same inputs."""
    torch, api, proxy = gpu
    g = dict(np.load(os.path.join(GOLDEN, f"{cfg}.npz")))
    ctx = api.Context(order, np.dtype(dt).itemsize)
    n = g["Q"].shape[0]
    data = proxy.ProxyData.from_host(ctx, {k: g[k] for k in ("Q", "star", "Aplus", "Aminus", "face_rel",
                                                             "neighbors", "face_types")})
    ctx.ader_local(api.strided(data.dofs), api.strided(data.buffers), None, api.strided(data.local_int),
                   data.face_types, g["coeffs"], n)
    torch.cuda.synchronize()
    assert rel_l2(data.to_logical(data.buffers), g["I"]) < step_tol(dt)
    assert rel_l2(data.to_logical(data.dofs), g["Q_local"]) < step_tol(dt)
    ftn = torch.from_numpy(g["face_types_neigh"]).to(data.device)
    ctx.neigh(api.strided(data.dofs), api.strided(data.buffers), api.strided(data.neigh_int), ftn,
              data.face_rel, n, nb_idx=data.neighbors)
    torch.cuda.synchronize()
    assert rel_l2(data.to_logical(data.dofs), g["Q_all"]) < step_tol(dt)


def test_reference_layout_and_pointer_tables(gpu):
    """Drop-in mode: the reference's 64-byte padded layout, addressed through `real**` tables
    (ConditionalPointersToRealsTable) instead of base+stride."""
    torch, api, proxy = gpu
    order, dt, n = 6, np.float32, 150
    ctx, host, data, tab, c = make(order, dt, n, api, proxy, seed=4, reference_layout=True)
    assert ctx.ld == 64 and ctx.der_ld[:3] == [64, 48, 32]

    keep = []  # pointer tables are device tensors: keep them alive until the launches finish

    def table(t):
        stride = t.stride(0) * t.element_size()
        keep.append(t.data_ptr() + stride * torch.arange(n, device=t.device, dtype=torch.int64))
        return api.pointer_table(keep[-1])

    nb_ptrs = (data.buffers.data_ptr() + data.buffers.stride(0) * data.buffers.element_size()
               * data.neighbors).contiguous()
    ctx.ader_local(table(data.dofs), table(data.buffers), None, table(data.local_int), data.face_types, c, n)
    ctx.neigh(table(data.dofs), api.NONE, table(data.neigh_int), data.face_types, data.face_rel, n,
              nb_ptrs=nb_ptrs)
    torch.cuda.synchronize()
    want = O.proxy_step(tab, host, c, "all")
    assert rel_l2(data.to_logical(data.dofs), want["Q"]) < step_tol(dt)
    # padding rows stay zero
    pad = data.dofs.reshape(n, 9, ctx.ld)[:, :, ctx.B:]
    assert float(pad.abs().max()) == 0.0


@pytest.mark.parametrize("n", [0, 1, 13, 14, 15])
def test_edge_sizes(gpu, n):
    torch, api, proxy = gpu
    order, dt = 4, np.float32
    ctx = api.Context(order, 4)
    if n == 0:
        z = torch.zeros((0, ctx.qsize), device="cuda", dtype=torch.float32)
        ctx.ader(api.strided(z), api.strided(z), None, api.strided(z), np.ones(order), 0)
        return
    ctx, host, data, tab, c = make(order, dt, n, api, proxy, seed=n)
    proxy.ProxyKernels(ctx, data).all()
    torch.cuda.synchronize()
    want = O.proxy_step(tab, host, c, "all")
    assert rel_l2(data.to_logical(data.dofs), want["Q"]) < step_tol(dt)


def test_linearity_at_full_size(gpu):
    """Size-independent property at BASELINE.json's full size (1M elements, O6 f32):
    all(a*Q) == a*all(Q) up to rounding, and the result is finite."""
    torch, api, proxy = gpu
    ctx = api.Context(6, 4)
    n = 1_000_000
    d1 = proxy.ProxyData.synthetic(ctx, n, seed=1)
    d1.sort_by_face_relation()
    d2 = proxy.ProxyData.synthetic(ctx, n, seed=1)
    d2.perm = d1.perm
    d2.dofs *= 2.0
    proxy.ProxyKernels(ctx, d1).all()
    proxy.ProxyKernels(ctx, d2).all()
    torch.cuda.synchronize()
    diff = (d2.dofs - 2.0 * d1.dofs).double().norm() / d2.dofs.double().norm()
    assert torch.isfinite(diff) and float(diff) < 1e-6


def test_entry_points_are_cuda_graph_capture_safe(gpu):
    """The reference enqueues its device kernels inside CUDA-graph capture (src/Parallel/Runtime/Stream.h:163-190): the
    entry points must not synchronise or allocate.  Capture `all` (ader, local, neigh) once, replay it, compare with
    eager launches."""
    torch, api, proxy = gpu
    order, dt, n = 6, np.float32, 1500
    ctx, host, data, tab, c = make(order, dt, n, api, proxy, seed=31)
    data.sort_by_face_relation()
    k = proxy.ProxyKernels(ctx, data)
    start = data.dofs.clone()
    k.all()                                   # warm-up outside the capture (one-time kernel attribute set-up)
    torch.cuda.synchronize()
    eager = data.dofs.clone()
    data.dofs.copy_(start)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            k.all(stream=side)
    torch.cuda.current_stream().wait_stream(side)
    data.dofs.copy_(start)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(data.dofs, eager)
    graph.replay()                            # second replay = second time step
    torch.cuda.synchronize()
    host = O.proxy_step(tab, O.proxy_step(tab, host, c, "all"), c, "all")
    assert rel_l2(data.to_logical(data.dofs), host["Q"]) < run_tol(dt)


def test_invalid_arguments_are_rejected(gpu):
    """Error behaviour at the boundary: bad arguments return SSB_EINVAL (-1) before anything is launched; the DOFs stay
    untouched."""
    import ctypes
    torch, api, proxy = gpu
    ctx = api.Context(6, 4)
    lib = ctx.lib
    d = proxy.ProxyData.synthetic(ctx, 64)
    before = d.dofs.clone()
    coeffs = (ctypes.c_double * 6)(*[1.0] * 6)
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ok, none = api.strided(d.dofs), api.NONE
    assert lib.ssb_ader(ctx.handle, none, api.strided(d.buffers), none, api.strided(d.local_int), coeffs, 64, None, s) == -1
    assert lib.ssb_ader(ctx.handle, ok, api.strided(d.buffers), none, api.strided(d.local_int), None, 64, None, s) == -1
    assert lib.ssb_ader(ctx.handle, ok, api.strided(d.buffers), none, api.strided(d.local_int), coeffs, -5, None, s) == -1
    assert lib.ssb_local(ctx.handle, ok, none, api.strided(d.local_int), None, 64, None, s) == -1
    assert lib.ssb_neigh(ctx.handle, ok, None, None, api.strided(d.buffers), api.strided(d.neigh_int), None,
                         ctypes.c_void_p(d.face_rel.data_ptr()), 64, None, s) == -1          # neither pointers nor indices
    assert lib.ssb_neigh(ctx.handle, ok, None, ctypes.c_void_p(d.neighbors.data_ptr()), api.strided(d.buffers),
                         api.strided(d.neigh_int), None, None, 64, None, s) == -1              # no face relations
    assert lib.ssb_set_layout(ctx.handle, 54, None, None) == -1                                # ld not a multiple of 4
    assert lib.ssb_set_layout(ctx.handle, 52, None, None) == -1                                # ld < padded basis size
    assert lib.ssb_visco_ader(ctx.handle, ok, ok, ok, ok, ok, coeffs, 64, s) == -1             # elastic context
    assert lib.ssb_lts_accumulate(ctx.handle, none, ok, 1, 64, s) == -1
    torch.cuda.synchronize()
    assert torch.equal(before, d.dofs)
    assert b"invalid" in lib.ssb_error_string(-1)


@pytest.mark.parametrize("order,dt", [(4, np.float32), (6, np.float32), (6, np.float64)])
def test_local_nodal_boundary_conditions(gpu, order, dt):
    """ssb_local_nodal_bc (Dirichlet with an easi boundary map, FreeSurfaceGravity) against the oracle, whose nodal
    path is pinned against the reference's generated kernels (tests/test_oracle_dr.py).  Some cells carry two boundary
    faces (atomic adds)."""
    from test_oracle_dr import nodal_bc_case
    torch, api, proxy = gpu
    n, m = 61, 90
    ctx, host, data, tab, c = make(order, dt, n, api, proxy)
    rng = np.random.default_rng(33)
    B2 = order * (order + 1) // 2
    I = rng.uniform(-1, 1, host["Q"].shape).astype(dt)
    bc = nodal_bc_case(rng, n, m, B2, dt)
    bc["kind"][::9] = 0      # a face type without a nodal boundary condition: skipped
    tab64 = O.load_tables(order, np.float64)
    want = O.local_nodal_bc(tab64, host["Q"].astype(np.float64), I.astype(np.float64), O.flux_dense(host["Aminus"]).astype(np.float64),
                            {k: (v.astype(np.float64) if v.dtype.kind == "f" else v) for k, v in bc.items()})
    dev = data.dofs.device
    data = proxy.ProxyData.from_host(ctx, dict(host, I=I))
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    ctx.local_nodal_bc(api.strided(data.dofs), api.strided(data.buffers), api.strided(data.neigh_int), d(bc["cell"]),
                       d(bc["face"]), d(bc["kind"]), d(bc["tinv"]), d(bc["easiMap"]), d(bc["easiConst"]), d(bc["displacement"]),
                       d(bc["rhoG"].astype(dt)), m)
    torch.cuda.synchronize()
    got = data.to_logical(data.dofs).astype(np.float64)
    t = 5e-6 if dt == np.float32 else 1e-13
    assert rel_l2(got, want) < t
    assert rel_l2(got - host["Q"], want - host["Q"]) < 10 * t
    assert np.unique(bc["cell"]).size < m      # at least one cell with two boundary faces


@pytest.mark.parametrize("order,dt", [(4, np.float32), (6, np.float32), (6, np.float64)])
def test_analytical_boundary_faces(gpu, order, dt):
    """ssb_local_nodal_flux (FaceType::Analytical: host callback integrated over the time step by
    seissol_b200.boundary.analytical_inodal, nodal flux on the device) against the oracle, which is pinned against the
    reference's generated localFluxNodal (tests/test_oracle_dr.py)."""
    from test_oracle_dr import nodal_bc_case, plane_wave_solution
    from seissol_b200 import boundary
    torch, api, proxy = gpu
    n, m = 61, 90
    ctx, host, data, tab, c = make(order, dt, n, api, proxy)
    rng = np.random.default_rng(34)
    B2 = order * (order + 1) // 2
    bc = nodal_bc_case(rng, n, m, B2, dt)
    nodes = rng.uniform(-1, 1, (m, B2, 3))
    inodal = boundary.analytical_inodal(plane_wave_solution, nodes, 0.1, 0.02, order, dtype=dt)     # [m, 9, B2]
    tab64 = O.load_tables(order, np.float64)
    want = O.local_nodal_flux(tab64, host["Q"].astype(np.float64), O.flux_dense(host["Aminus"]).astype(np.float64),
                              bc["cell"], bc["face"], inodal.astype(np.float64).transpose(0, 2, 1))
    dev = data.dofs.device
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    ctx.local_nodal_flux(api.strided(data.dofs), api.strided(data.neigh_int), d(bc["cell"]), d(bc["face"]), d(inodal), m)
    torch.cuda.synchronize()
    got = data.to_logical(data.dofs).astype(np.float64)
    t = 5e-6 if dt == np.float32 else 1e-13
    assert rel_l2(got, want) < t
    assert rel_l2(got - host["Q"], want - host["Q"]) < 10 * t
    lib = api.load_library()
    assert lib.ssb_local_nodal_flux(ctx.handle, api.strided(data.dofs), api.strided(data.neigh_int), None, None, None, 3, None) == -1


@pytest.mark.parametrize("order,dt", [(4, np.float32), (6, np.float32), (6, np.float64)])
def test_volume_energies(gpu, order, dt):
    """ssb_volume_energies (modal integrals, one streaming pass) against the oracle's restatement of
    EnergyOutput::computeVolumeEnergies (evaluation at the (O+1)^3 quadrature points); elastic and acoustic elements,
    ragged count, accumulation over two calls."""
    torch, api, proxy = gpu
    n = 1237
    ctx, host, data, tab, c = make(order, dt, n, api, proxy)
    rng = np.random.default_rng(12)
    vol = rng.uniform(0.5, 2.0, n)
    mat = np.stack([rng.uniform(1.0, 3.0, n), rng.uniform(0.5, 2.0, n), rng.uniform(1.0, 3.0, n)], -1)
    mat[::7, 1] = 0.0                                   # acoustic elements
    want = O.volume_energies(order, host["Q"].astype(np.float64), vol, mat)
    dev = data.dofs.device
    out = torch.zeros(8, dtype=torch.float64, device=dev)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    ctx.volume_energies(api.strided(data.dofs), d(vol), d(mat), n, out)
    got = out.cpu().numpy()
    t = 2e-6 if dt == np.float32 else 1e-12
    assert np.all(np.abs(got[:7] - want[:7]) <= t * np.abs(want[:7]).max()), (got, want)
    assert want[3] > 0 and want[0] > 0
    ctx.volume_energies(api.strided(data.dofs), d(vol), d(mat), n, out)          # accumulates
    assert np.allclose(out.cpu().numpy()[:7], 2 * got[:7], rtol=1e-12)
    lib = api.load_library()
    assert lib.ssb_volume_energies(ctx.handle, api.strided(data.dofs), None, None, 5, None, None) == -1
