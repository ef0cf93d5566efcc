"""Physics check of the whole scheme: elastic plane waves in a fully periodic cube mesh converge to the exact solution
at the order of the discretisation.  Uses the restated cell-local matrix set-up (seissol_b200/material.py, SURVEY.md §8f
item 1: star matrices, Godunov flux solvers from geometry + material), the modal basis (seissol_b200/basis.py) and
  * CPU: the numpy oracle as the time stepper,
  * GPU (-m gpu): the CUDA kernels through the C ABI, which must reproduce the oracle run and its error level."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import aderdg as O
from seissol_b200 import basis as Bs
from seissol_b200 import material as Mt
from seissol_b200 import mesh as M

LAM, MU, RHO = 2.0, 1.0, 1.0


def setup(order, n, kind, dtype=np.float64):
    tab = O.load_tables(order, dtype)
    m = M.CubeMesh(n, n, n, periodic_yz=True)
    h = 1.0 / n
    coords = m.coords * h
    N = m.n
    star, ap, am = Mt.cell_local_matrices(coords, m.neighbors, np.full(N, LAM), np.full(N, MU), np.full(N, RHO),
                                          tab["star_pattern"])
    f, _ = Mt.plane_wave(LAM, MU, RHO, (1, 1, 1), 1, 1.0, kind)
    Q = Bs.project(order, coords, lambda x: f(x, 0.0))
    cp = np.sqrt((LAM + 2 * MU) / RHO)
    d = {"Q": Q.astype(dtype), "star": star.astype(dtype), "Aplus": ap.astype(dtype), "Aminus": am.astype(dtype),
         "face_rel": m.face_rel, "neighbors": m.neighbors, "face_types": np.zeros((N, 4), dtype=np.uint8)}
    return tab, coords, f, d, 0.25 * h / (cp * (2 * order - 1))


def run_oracle(order, n, kind, t_end=0.05):
    tab, coords, f, d, dt = setup(order, n, kind)
    steps = int(np.ceil(t_end / dt))
    coeffs = O.time_integrate_coeffs(order, 0.0, t_end / steps)
    for _ in range(steps):
        d = O.proxy_step(tab, d, coeffs, "all")
    err = np.linalg.norm(Bs.l2_error(order, coords, d["Q"], lambda x: f(x, t_end)))
    ref = np.linalg.norm(Bs.l2_error(order, coords, 0 * d["Q"], lambda x: f(x, t_end)))
    proj = np.linalg.norm(Bs.l2_error(order, coords, Bs.project(order, coords, lambda x: f(x, t_end)), lambda x: f(x, t_end)))
    return err / ref, proj / ref, d["Q"], steps


def test_basis_matches_operator_tables():
    """The Dubiner basis restated here is the one the operator tables were generated for:
    kDivM = M^-1 K, kDivMT = M^-1 K^T with K[k, l] = int dphi_k/dxi phi_l (mass matrix diagonal)."""
    order = 4
    tab = O.load_tables(order, np.float64)
    pts, w = Bs.tet_quadrature(order + 2)
    phi = Bs.dubiner(order, pts[:, 0], pts[:, 1], pts[:, 2])
    mass = np.einsum("m,mk,ml->kl", w, phi, phi)
    assert abs(w.sum() - 1.0 / 6.0) < 1e-14 and np.abs(mass - np.diag(np.diag(mass))).max() < 1e-14
    h = 1e-6
    for d in range(3):
        sh = [h * (d == i) for i in range(3)]
        dphi = (Bs.dubiner(order, pts[:, 0] + sh[0], pts[:, 1] + sh[1], pts[:, 2] + sh[2])
                - Bs.dubiner(order, pts[:, 0] - sh[0], pts[:, 1] - sh[1], pts[:, 2] - sh[2])) / (2 * h)
        K = np.einsum("m,mk,ml->kl", w, dphi, phi)
        assert np.abs(np.diag(1 / np.diag(mass)) @ K - tab["kDivM"][d]).max() < 1e-7
        assert np.abs(np.diag(1 / np.diag(mass)) @ K.T + tab["kDivMT"][d]).max() < 1e-7     # load_tables negates kDivMT


def test_flux_solver_properties():
    """Godunov state: QgodLocal + QgodNeighbor = I; rotation: T Tinv = I; homogeneous material: AplusT + AminusT is the
    rotated normal flux matrix scaled by the face, independent of the tangent choice."""
    ql, qn = Mt.godunov_state_T(LAM, MU, RHO, 1.3 * LAM, 0.7 * MU, 1.1 * RHO)
    assert np.abs(ql + qn - np.eye(9)).max() < 1e-14
    n = np.array([1.0, 2.0, 2.0]) / 3
    s = np.array([2.0, 1.0, -2.0]) / 3
    T, Ti = Mt.rotation_matrices(n, s, np.cross(n, s))
    assert np.abs(T @ Ti - np.eye(9)).max() < 1e-14
    # sum of both flux solvers = normal flux: sum_d n_d A_d^T (consistency of the numerical flux)
    at = Mt.coefficient_matrix_T(LAM, MU, RHO, 0)
    ql, qn = Mt.godunov_state_T(LAM, MU, RHO, LAM, MU, RHO)
    total = Ti.T @ ((ql + qn) @ at) @ T.T
    want = sum(n[d] * Mt.coefficient_matrix_T(LAM, MU, RHO, d) for d in range(3))
    assert np.abs(total - want).max() < 1e-13


@pytest.mark.parametrize("order,kind", [(3, "P"), (3, "S"), (4, "P")])
def test_plane_wave_convergence(order, kind):
    e4, p4, _, _ = run_oracle(order, 4, kind)
    e8, p8, _, _ = run_oracle(order, 8, kind)
    rate = np.log2(e4 / e8)
    assert rate > order - 0.5, (e4, e8, rate)
    assert e8 < 2.0 * p8            # the discretisation error stays at the level of the projection error


@pytest.mark.gpu
@pytest.mark.parametrize("order,dt,bar", [(4, np.float64, 1e-11), (4, np.float32, 1e-5), (6, np.float32, 1e-5)])
def test_plane_wave_cuda(built_library, order, dt, bar):
    import torch
    from seissol_b200 import api, proxy
    n, kind, t_end = 4, "P", 0.05
    tab, coords, f, d, step = setup(order, n, kind, dt)
    steps = int(np.ceil(t_end / step))
    coeffs = O.time_integrate_coeffs(order, 0.0, t_end / steps, dt)
    ctx = api.Context(order, np.dtype(dt).itemsize)
    data = proxy.ProxyData.from_host(ctx, d)
    data.sort_by_face_relation()
    k = proxy.ProxyKernels(ctx, data)
    k.coeffs = coeffs
    want = dict(d)
    for _ in range(steps):
        k.all()
        want = O.proxy_step(tab, want, coeffs, "all")
    torch.cuda.synchronize()
    got = data.to_logical(data.dofs)
    assert rel_l2(got, want["Q"]) < bar
    err = np.linalg.norm(Bs.l2_error(order, coords, got.astype(np.float64), lambda x: f(x, t_end)))
    ref = np.linalg.norm(Bs.l2_error(order, coords, 0 * got.astype(np.float64), lambda x: f(x, t_end)))
    assert err / ref < (2e-2 if order == 4 else 2e-3)
    # energy output on the device (ssb_volume_energies): equals the oracle's quadrature formulation, and has the physics of
    # a travelling plane wave -- kinetic = strain energy (equipartition), total energy conserved up to the (tiny) upwind
    # dissipation, momentum of a full period = 0
    N = got.shape[0]
    vol = np.abs(np.linalg.det(coords[:, 1:] - coords[:, :1])) / 6.0
    mat = np.tile(np.array([LAM, MU, RHO]), (N, 1))
    dev = data.dofs.device
    out = torch.zeros(8, dtype=torch.float64, device=dev)
    ctx.volume_energies(api.strided(data.dofs), torch.from_numpy(vol).to(dev),
                        torch.from_numpy(mat).to(dev), N, out)
    e = out.cpu().numpy()
    want_e = O.volume_energies(order, got.astype(np.float64), vol, mat)
    assert np.all(np.abs(e[:7] - want_e[:7]) <= 1e-6 * np.abs(want_e[:2]).max())
    e0 = O.volume_energies(order, d["Q"].astype(np.float64), vol, mat)
    assert abs(e[0] - e[1]) < 2e-3 * (e[0] + e[1])
    assert 0.995 < (e[0] + e[1]) / (e0[0] + e0[1]) < 1.0 + 1e-5
    assert np.all(np.abs(e[4:7]) < 1e-6 * np.sqrt(e[0]))


def test_plane_wave_lts():
    """Clustered LTS (3 clusters, rate 2) on the periodic mesh propagates the plane wave as accurately as global time
    stepping with the smallest step: physics-level check of the LTS schedule and its buffer/derivative plumbing."""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(__file__))
    from lts_support import NumpyLtsBackend
    from seissol_b200 import lts as L
    from seissol_b200.partition import RankLayout
    order, n, t_end = 4, 8, 0.05
    tab = O.load_tables(order, np.float64)
    m = M.CubeMesh(n, n, n, periodic_yz=True)
    coords = m.coords / n
    N = m.n
    band = np.array([0, 0, 1, 1, 2, 2, 1, 1])                 # clusters as bands in z, periodic, neighbors differ by <= 1
    cluster = band[m.cube_z]
    layout = RankLayout(rank=0, world=1, n_own=N, n_ghost=0, n_copy=0, gid=np.arange(N), cluster=cluster,
                        neighbors=m.neighbors, face_rel=m.face_rel, face_types=np.zeros((N, 4), dtype=np.uint8))
    plan = L.build_plan(layout, 3)
    assert all(r.end > r.start for r in plan.ranges) and plan.n_acc > 0
    pc = coords[plan.gid]                                      # geometry in the plan's cell order
    star, ap, am = Mt.cell_local_matrices(pc, plan.neighbors, np.full(N, LAM), np.full(N, MU), np.full(N, RHO),
                                          tab["star_pattern"])
    f, _ = Mt.plane_wave(LAM, MU, RHO, (1, 1, 1), 1, 1.0, "P")
    Q0 = Bs.project(order, pc, lambda x: f(x, 0.0))
    cp = np.sqrt((LAM + 2 * MU) / RHO)
    # the mesh is uniform, so the LARGEST cluster step (4 dt0) has to respect the CFL bound of the cells
    dt0 = 0.25 * (1.0 / n) / (cp * (2 * order - 1)) / 4
    macro = int(np.ceil(t_end / (4 * dt0)))
    dt0 = t_end / (4 * macro)
    be = NumpyLtsBackend(tab, order, plan, Q0, O.star_dense(star, tab["star_pattern"]), O.flux_dense(ap), O.flux_dense(am))
    solver = L.LtsSolver(plan, be, order, dt0=dt0)
    for _ in range(macro):
        solver.macro_step()
    ref = np.linalg.norm(Bs.l2_error(order, pc, 0 * Q0, lambda x: f(x, t_end)))
    err_lts = np.linalg.norm(Bs.l2_error(order, pc, be.Q, lambda x: f(x, t_end))) / ref
    # global time stepping with dt0 on the same mesh
    d = {"Q": Q0, "star": star, "Aplus": ap, "Aminus": am, "face_rel": plan.face_rel, "neighbors": plan.neighbors,
         "face_types": plan.face_types}
    coeffs = O.time_integrate_coeffs(order, 0.0, dt0)
    for _ in range(4 * macro):
        d = O.proxy_step(tab, d, coeffs, "all")
    err_gts = np.linalg.norm(Bs.l2_error(order, pc, d["Q"], lambda x: f(x, t_end))) / ref
    assert err_gts < 2e-3
    assert err_lts < 1.5 * err_gts + 1e-4, (err_lts, err_gts)


def _shear_pulse_energy(bc_type, order=4, n=6, t_end=0.5):
    """SH pulse (sigma_xz, u) travelling in +z inside a box: periodic in x, free surface on the y faces, the z faces carry
    the boundary condition under test.  Returns (E(t_end) / E(0))."""
    tab = O.load_tables(order, np.float64)
    m = M.CubeMesh(n, n, n)
    coords = m.coords / n
    N = m.n
    # y faces: traction free (exactly satisfied by this wave, so they do not disturb it; an absorbing boundary would
    # treat the tangential velocity as an outgoing shear wave and damp it)
    ft = np.where(m.neighbors < 0, 1, 0).astype(np.uint8)
    zc = coords[:, M.FACE2NODES, 2]                                             # [N, 4 faces, 3 vertices]
    on_z = (m.neighbors < 0) & (np.all(zc < 1e-12, axis=2) | np.all(zc > 1 - 1e-12, axis=2))
    ft[on_z] = bc_type                                                          # ... except the z faces
    nb = np.where(m.neighbors < 0, np.arange(N)[:, None], m.neighbors)
    star, ap, am = Mt.cell_local_matrices(coords, m.neighbors, np.full(N, LAM), np.full(N, MU), np.full(N, RHO),
                                          tab["star_pattern"], ft)
    cs = np.sqrt(MU / RHO)

    def pulse(x):
        g = np.exp(-((x[..., 2] - 0.6) / 0.12) ** 2)
        out = np.zeros(x.shape[:-1] + (9,))
        out[..., 5] = MU * g          # sigma_xz
        out[..., 6] = -cs * g         # u: right-going in z
        return out

    def energy(Q):
        pts, w = Bs.tet_quadrature(order + 2)
        phi = Bs.dubiner(order, pts[:, 0], pts[:, 1], pts[:, 2])
        q = np.einsum("mk,nkq->nmq", phi, Q)
        jac = np.abs(np.linalg.det(coords[:, 1:] - coords[:, :1]))
        dens = 0.5 * RHO * (q[..., 6] ** 2 + q[..., 7] ** 2 + q[..., 8] ** 2) + (q[..., 3] ** 2 + q[..., 4] ** 2 + q[..., 5] ** 2) / (2 * MU)
        return float(np.einsum("n,m,nm->", jac, w, dens))

    d = {"Q": Bs.project(order, coords, pulse), "star": star, "Aplus": ap, "Aminus": am, "face_rel": m.face_rel,
         "neighbors": nb, "face_types": ft}
    e0 = energy(d["Q"])
    cp = np.sqrt((LAM + 2 * MU) / RHO)
    dt = 0.25 * (1.0 / n) / (cp * (2 * order - 1))
    steps = int(np.ceil(t_end / dt))
    coeffs = O.time_integrate_coeffs(order, 0.0, t_end / steps)
    for _ in range(steps):
        d = O.proxy_step(tab, d, coeffs, "all")
    return energy(d["Q"]) / e0


def test_free_surface_reflects_and_outflow_absorbs():
    """Boundary face types only change the local flux solver (material.py): a free surface (FaceType 1) sends the pulse
    back (energy stays), an absorbing boundary (FaceType 5) lets it leave (energy drops)."""
    kept = _shear_pulse_energy(1)
    left = _shear_pulse_energy(5)
    assert 0.85 < kept < 1.0 + 1e-6, kept      # upwind fluxes dissipate a little, never create energy
    assert left < 0.15, left


POINTS = np.array([[0.31, 0.42, 0.53], [0.77, 0.15, 0.6], [0.05, 0.93, 0.21]])


def test_receivers_follow_the_plane_wave():
    """Receivers (seissol_b200/receiver.py) sample the predictor's Taylor expansion inside every step: the time series at
    three points follow the exact plane wave."""
    from seissol_b200 import receiver as Rc
    order, n, t_end = 4, 8, 0.04
    tab, coords, f, d, dt = setup(order, n, "P")
    steps = int(np.ceil(t_end / dt))
    dt = t_end / steps
    coeffs = O.time_integrate_coeffs(order, 0.0, dt)
    rc = Rc.ReceiverCluster(order, coords, POINTS, sampling_interval=dt / 3)
    star = O.star_dense(d["star"], tab["star_pattern"])
    for s in range(steps):
        _, ders = O.ader(tab, d["Q"][rc.elements], star[rc.elements], coeffs, True)
        rc.step(ders, s * dt, dt)
        d = O.proxy_step(tab, d, coeffs, "all")
    times, series = rc.series()
    assert len(times) == 3 * steps and abs(times[1] - dt / 3) < 1e-15
    exact = np.stack([f(POINTS, t) for t in times], 1)
    assert rel_l2(series, exact) < 5e-3


@pytest.mark.gpu
def test_receivers_cuda(built_library):
    """Same receivers with the CUDA kernels: the derivative families of the receiver elements come from `ssb_ader` on a
    pointer-table batch of three elements."""
    import torch
    from seissol_b200 import api, proxy
    from seissol_b200 import receiver as Rc
    order, n, t_end, dtp = 4, 8, 0.04, np.float64
    tab, coords, f, d, dt = setup(order, n, "P", dtp)
    steps = int(np.ceil(t_end / dt))
    dt = t_end / steps
    coeffs = O.time_integrate_coeffs(order, 0.0, dt, dtp)
    ctx = api.Context(order, 8)
    data = proxy.ProxyData.from_host(ctx, d)
    data.sort_by_face_relation()
    k = proxy.ProxyKernels(ctx, data)
    k.coeffs = coeffs
    rc = Rc.ReceiverCluster(order, coords, POINTS, sampling_interval=dt / 3)
    for s in range(steps):
        rc.step(Rc.receiver_derivatives_cuda(torch, api, ctx, data, rc.elements, coeffs), s * dt, dt)
        k.all()
    torch.cuda.synchronize()
    times, series = rc.series()
    exact = np.stack([f(POINTS, t) for t in times], 1)
    assert rel_l2(series, exact) < 5e-3
