"""GPU parity tests of the dynamic-rupture kernels (-m gpu), through the C ABI, against the CPU oracle
(oracle/aderdg.py::dr_interpolate / neighbor(dr=...), oracle/friction.py)."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import aderdg as O
from oracle import friction as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(built_library):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    from seissol_b200 import api, proxy
    return torch, api, proxy


def tol(dt):
    return 5e-6 if dt == np.float32 else 1e-13


@pytest.mark.parametrize("order,dt", [(4, np.float32), (5, np.float32), (6, np.float32), (4, np.float64), (6, np.float64)])
@pytest.mark.parametrize("reference_layout", [False, True])
def test_godunov_dr_interpolation(gpu, order, dt, reference_layout):
    torch, api, proxy = gpu
    ctx = api.Context(order, np.dtype(dt).itemsize, reference_layout=reference_layout)
    tab = O.load_tables(order, dt)
    rng = np.random.default_rng(3)
    cells, faces = 40, 77   # ragged vs faces-per-CTA
    B, P = O.nbasis(order), (order + 1) ** 2
    ders = []
    for d in range(order):
        dq = np.zeros((cells, B, 9), dtype=dt)
        dq[:, :O.nbasis(order - d)] = rng.uniform(0, 1, (cells, O.nbasis(order - d), 9))
        ders.append(dq)
    plus = rng.integers(0, cells, faces)
    minus = rng.integers(0, cells, faces)
    info = np.stack([rng.integers(0, 4, faces), rng.integers(0, 4, faces), rng.integers(0, 4, faces)], 1).astype(np.uint8)
    tinv = rng.uniform(-1, 1, (faces, 9, 9)).astype(dt)
    pts, _ = proxy.gauss_legendre_points(order, 0.7)
    coeffs = np.stack([proxy.point_coeffs(order, tau) for tau in pts])
    want_p, want_m = O.dr_interpolate(tab, [d[plus] for d in ders], [d[minus] for d in ders], tinv, info,
                                      coeffs.astype(dt))
    dev = torch.device("cuda")
    derivs = torch.from_numpy(proxy.ders_to_device(ctx, ders, dt)).to(dev)
    stride = derivs.stride(0) * derivs.element_size()
    pp = (derivs.data_ptr() + torch.from_numpy(plus).to(dev) * stride).contiguous()
    pm = (derivs.data_ptr() + torch.from_numpy(minus).to(dev) * stride).contiguous()
    tdt = torch.float32 if dt == np.float32 else torch.float64
    qp = torch.full((faces, order, 9, ctx.ppad), float("nan"), device=dev, dtype=tdt)
    qm = torch.full_like(qp, float("nan"))
    tinv_mem = torch.from_numpy(np.ascontiguousarray(tinv.transpose(0, 2, 1)).reshape(faces, 81)).to(dev)
    ctx.dr_interp(api.pointer_table(pp), api.pointer_table(pm), tinv_mem, torch.from_numpy(info).to(dev), qp, qm,
                  coeffs, faces)
    torch.cuda.synchronize()
    got_p = proxy.points_from_device(ctx, qp.cpu().numpy(), P)
    got_m = proxy.points_from_device(ctx, qm.cpu().numpy(), P)
    assert rel_l2(got_p, want_p) < tol(dt)
    assert rel_l2(got_m, want_m) < tol(dt)


def _friction_device_data(torch, ctx, proxy, st, imp, qip, qim, resample, law):
    dev = torch.device("cuda")
    P = qip.shape[2]

    def pts(x):  # [nf, P] -> [nf, ppad]
        out = np.zeros((x.shape[0], ctx.ppad), dtype=x.dtype)
        out[:, :P] = x
        return torch.from_numpy(out).to(dev)

    def flags(x):
        out = np.zeros((x.shape[0], ctx.ppad), dtype=np.uint8)
        out[:, :P] = x.astype(np.uint8)
        return torch.from_numpy(out).to(dev)

    d = {"q_plus": torch.from_numpy(proxy.points_to_device(ctx, qip)).to(dev),
         "q_minus": torch.from_numpy(proxy.points_to_device(ctx, qim)).to(dev)}
    nf = qip.shape[0]
    d["imposed_plus"] = torch.zeros((nf, 9, ctx.ppad), device=dev, dtype=d["q_plus"].dtype)
    d["imposed_minus"] = torch.zeros_like(d["imposed_plus"])
    d["imp_and_eta"] = torch.from_numpy(np.stack([imp[k] for k in F.IMP_FIELDS], 1).astype(qip.dtype)).to(dev)
    stress = np.zeros((nf, 6, ctx.ppad), dtype=qip.dtype)
    stress[:, :, :P] = st["initialStress"]
    d["initial_stress"] = torch.from_numpy(stress).to(dev)
    names = {"mu": "mu", "slip1": "slip1", "slip2": "slip2", "slip_rate1": "slipRate1", "slip_rate2": "slipRate2",
             "slip_rate_magnitude": "slipRateMagnitude", "accumulated_slip_magnitude": "accumulatedSlipMagnitude",
             "traction1": "traction1", "traction2": "traction2", "rupture_time": "ruptureTime",
             "peak_slip_rate": "peakSlipRate", "dyn_stress_time": "dynStressTime"}
    if law == "lsw":
        names.update({"d_c": "dC", "mu_s": "muS", "mu_d": "muD", "cohesion": "cohesion",
                      "forced_rupture_time": "forcedRuptureTime"})
    else:
        names.update({"rs_a": "a", "rs_sl0": "sl0", "rs_f0": "f0", "rs_mu_w": "muW", "rs_b": "b", "rs_sr_w": "srW",
                      "state_variable": "stateVariable"})
        d["convergence_inner"] = flags(st["convergenceInner"])
        d["convergence_outer"] = flags(st["convergenceOuter"])
    for k, v in names.items():
        d[k] = pts(st[v])
    d["rupture_time_pending"] = flags(st["ruptureTimePending"])
    d["dyn_stress_time_pending"] = flags(st["dynStressTimePending"])
    d["resample_t"] = torch.from_numpy(np.ascontiguousarray(resample.T.astype(qip.dtype))).to(dev)
    return d, names


@pytest.mark.parametrize("law", ["lsw", "rsfvw"])
@pytest.mark.parametrize("order,dt", [(6, np.float64), (4, np.float64), (6, np.float32)])
def test_friction_laws(gpu, law, order, dt):
    torch, api, proxy = gpu
    ctx = api.Context(order, np.dtype(dt).itemsize)
    tab = O.load_tables(order, dt)
    nf = 37
    P = (order + 1) ** 2
    st, imp, params = F.fault_state(law, nf, order, dt)
    qip, qim = F.fault_qinterpolated(nf, order, dt)
    d, names = _friction_device_data(torch, ctx, proxy, st, imp, qip, qim, tab["resample"], law)
    step = 1e-3
    fp, dts, wts = proxy.friction_params(order, step, full_update_time=0.25, rs_sr0=params.rsSr0,
                                         healing_threshold=params.healingThreshold, t0=params.t0)
    ctx.friction(law, d, fp, nf)
    torch.cuda.synchronize()
    lawobj = (F.LinearSlipWeakening if law == "lsw" else F.RateAndStateFastVelocityWeakening)(params, tab["resample"])
    plus, minus = F.evaluate(lawobj, st, imp, qip, qim, 0.25, dts, wts)
    # single precision rate-and-state: the Newton iteration count depends on rounding (the reference's
    # default tolerance 1e-8 is below float resolution), so only agreement to solver accuracy can be asked
    t = (1e-11 if law == "lsw" else 1e-8) if dt == np.float64 else (2e-5 if law == "lsw" else 2e-2)
    got_p = proxy.points_from_device(ctx, d["imposed_plus"].cpu().numpy(), P)
    got_m = proxy.points_from_device(ctx, d["imposed_minus"].cpu().numpy(), P)
    assert rel_l2(got_p, plus) < t
    assert rel_l2(got_m, minus) < t
    for k, v in names.items():
        if k in ("d_c", "mu_s", "mu_d", "cohesion", "forced_rupture_time", "rs_a", "rs_sl0", "rs_f0", "rs_mu_w",
                 "rs_b", "rs_sr_w"):
            continue
        got = d[k].cpu().numpy()[:, :P]
        want = st[v]
        scale = max(np.abs(want).max(), 1e-30)
        assert np.abs(got - want).max() / scale < (100 * t), (k, np.abs(got - want).max(), scale)
    if dt == np.float64:
        assert np.array_equal(d["rupture_time_pending"].cpu().numpy()[:, :P].astype(bool), st["ruptureTimePending"])
        assert np.array_equal(d["dyn_stress_time_pending"].cpu().numpy()[:, :P].astype(bool), st["dynStressTimePending"])
        if law == "rsfvw":
            assert np.array_equal(d["convergence_inner"].cpu().numpy()[:, :P].astype(bool), st["convergenceInner"])
            assert np.array_equal(d["convergence_outer"].cpu().numpy()[:, :P].astype(bool), st["convergenceOuter"])


@pytest.mark.parametrize("order,dt", [(4, np.float32), (6, np.float32), (6, np.float64)])
def test_neigh_dr_nodal_flux(gpu, order, dt):
    torch, api, proxy = gpu
    ctx = api.Context(order, np.dtype(dt).itemsize)
    tab = O.load_tables(order, dt)
    n, m = 53, 20
    P = (order + 1) ** 2
    host = O.proxy_data(order, n, dt, seed=8)
    rng = np.random.default_rng(5)
    ft = np.zeros((n, 4), dtype=np.uint8)
    ft[rng.uniform(size=(n, 4)) < 0.4] = 3
    dr = {"godunov": rng.uniform(-1, 1, (m, P, 9)).astype(dt), "fluxSolver": (rng.uniform(-1, 1, (m, 9, 9)) * 1e-2).astype(dt),
          "index": rng.integers(0, m, (n, 4)), "side": rng.integers(0, 4, (n, 4)).astype(np.uint8),
          "rel": rng.integers(0, 4, (n, 4)).astype(np.uint8)}
    # only the DR faces contribute in this test: regular neighbor part removed by zero Aminus
    want = O.neighbor(tab, host["Q"], np.zeros_like(host["Q"]), host["neighbors"], O.flux_dense(np.zeros_like(host["Aminus"])),
                      host["face_rel"], ft, dr=dr)
    data = proxy.ProxyData.from_host(ctx, host)
    dev = data.device
    god = torch.from_numpy(proxy.points_to_device(ctx, dr["godunov"])).to(dev)
    fs = torch.from_numpy(np.ascontiguousarray(dr["fluxSolver"].transpose(0, 2, 1)).reshape(m, 81)).to(dev)
    idx = torch.from_numpy(dr["index"]).to(dev)
    es = god.element_size()
    gp = (god.data_ptr() + idx * (9 * ctx.ppad * es)).contiguous()
    fp = (fs.data_ptr() + idx * (81 * es)).contiguous()
    sr = torch.from_numpy(np.stack([dr["side"], dr["rel"]], -1).astype(np.uint8)).to(dev).contiguous()
    ctx.neigh_dr(api.strided(data.dofs), gp, fp, sr, torch.from_numpy(ft).to(dev), n)
    torch.cuda.synchronize()
    got = data.to_logical(data.dofs)
    assert rel_l2(got, want) < tol(dt)
    assert rel_l2(got - host["Q"], want - host["Q"]) < 50 * tol(dt)


@pytest.mark.parametrize("order", [4, 6])
def test_fault_time_step_chain(gpu, order):
    """One time step across a fault, all kernels chained on the device exactly as TimeCluster does
    (ader with derivatives -> local (DR faces skipped) -> godunov_dr interpolation of both sides -> LSW friction ->
    neighbor flux of the regular faces + nodal flux of the imposed state), against the same chain of oracle functions.
    Checks that the layouts the kernels hand to each other agree (QInterpolated, ImposedState, CellDRMapping pointers)."""
    torch, api, proxy = gpu
    dt, step = np.float64, 1e-3
    ctx = api.Context(order, 8)
    tab = O.load_tables(order, dt)
    rng = np.random.default_rng(17)
    n, m = 60, 11
    P = (order + 1) ** 2
    host = O.proxy_data(order, n, dt, seed=23)
    # fault faces: distinct (cell, side) pairs for the plus and the minus side
    slots = rng.permutation(4 * n)[:2 * m]
    plus_cell, plus_side = slots[:m] // 4, slots[:m] % 4
    minus_cell, minus_side = slots[m:] // 4, slots[m:] % 4
    rel = rng.integers(1, 4, m)           # DRFaceInformation::faceRelation of the minus side (orientation + 1)
    ft = np.zeros((n, 4), dtype=np.uint8)
    ft[plus_cell, plus_side] = 3
    ft[minus_cell, minus_side] = 3
    host["face_types"] = ft
    info = np.stack([plus_side, minus_side, rel], 1).astype(np.uint8)
    tinv = rng.uniform(-1, 1, (m, 9, 9)).astype(dt)
    flux_solver = (rng.uniform(-1, 1, (2 * m, 9, 9)) * 1e-9).astype(dt)     # rows: plus faces, then minus faces
    st, imp, params = F.fault_state("lsw", m, order, dt)
    c_int = O.time_integrate_coeffs(order, 0.0, step, dt)
    pts, _ = proxy.gauss_legendre_points(order, step)
    c_pts = np.stack([proxy.point_coeffs(order, tau) for tau in pts]).astype(dt)
    fp, dts, wts = proxy.friction_params(order, step, full_update_time=0.0, healing_threshold=params.healingThreshold, t0=params.t0)

    # ---- oracle chain
    star = O.star_dense(host["star"], tab["star_pattern"])
    I, ders = O.ader(tab, host["Q"], star, c_int, True)
    q_loc = O.local(tab, host["Q"], I, star, O.flux_dense(host["Aplus"]), ft)
    qip, qim = O.dr_interpolate(tab, [d[plus_cell] for d in ders], [d[minus_cell] for d in ders], tinv, info, c_pts)
    st_o = {k: v.copy() for k, v in st.items()}
    law = F.LinearSlipWeakening(params, tab["resample"])
    imp_plus, imp_minus = F.evaluate(law, st_o, imp, qip, qim, 0.0, dts, wts)
    dr_index = np.zeros((n, 4), dtype=np.int64)
    dr_side = np.zeros((n, 4), dtype=np.uint8)
    dr_rel = np.zeros((n, 4), dtype=np.uint8)
    dr_index[plus_cell, plus_side] = np.arange(m)
    dr_side[plus_cell, plus_side] = plus_side
    dr_index[minus_cell, minus_side] = m + np.arange(m)
    dr_side[minus_cell, minus_side] = minus_side
    dr_rel[minus_cell, minus_side] = rel
    want = O.neighbor(tab, q_loc, I, host["neighbors"], O.flux_dense(host["Aminus"]), host["face_rel"], ft,
                      dr={"godunov": np.concatenate([imp_plus, imp_minus]), "fluxSolver": flux_solver, "index": dr_index,
                          "side": dr_side, "rel": dr_rel})

    # ---- device chain
    dev = torch.device("cuda")
    data = proxy.ProxyData.from_host(ctx, host)
    derivs = torch.zeros((n, ctx.dersize), device=dev, dtype=torch.float64)
    ctx.ader(api.strided(data.dofs), api.strided(data.buffers), api.strided(derivs), api.strided(data.local_int), c_int, n)
    ctx.local(api.strided(data.dofs), api.strided(data.buffers), api.strided(data.local_int), data.face_types, n)
    stride = derivs.stride(0) * derivs.element_size()
    pp = (derivs.data_ptr() + torch.from_numpy(plus_cell).to(dev) * stride).contiguous()
    pm = (derivs.data_ptr() + torch.from_numpy(minus_cell).to(dev) * stride).contiguous()
    qp = torch.zeros((m, order, 9, ctx.ppad), device=dev, dtype=torch.float64)
    qm = torch.zeros_like(qp)
    tinv_mem = torch.from_numpy(np.ascontiguousarray(tinv.transpose(0, 2, 1)).reshape(m, 81)).to(dev)
    ctx.dr_interp(api.pointer_table(pp), api.pointer_table(pm), tinv_mem, torch.from_numpy(info).to(dev), qp, qm, c_pts, m)
    d, names = _friction_device_data(torch, ctx, proxy, st, imp, np.zeros((m, order, P, 9), dtype=dt),
                                     np.zeros((m, order, P, 9), dtype=dt), tab["resample"], "lsw")
    d["q_plus"], d["q_minus"] = qp, qm
    ctx.friction("lsw", d, fp, m)
    imposed = torch.cat([d["imposed_plus"], d["imposed_minus"]]).contiguous()                 # [2m, 9, ppad]
    fs_mem = torch.from_numpy(np.ascontiguousarray(flux_solver.transpose(0, 2, 1)).reshape(2 * m, 81)).to(dev)
    idx = torch.from_numpy(dr_index).to(dev)
    gp = (imposed.data_ptr() + idx * (9 * ctx.ppad * 8)).contiguous()
    fsp = (fs_mem.data_ptr() + idx * (81 * 8)).contiguous()
    sr = torch.from_numpy(np.stack([dr_side, dr_rel], -1).astype(np.uint8)).to(dev).contiguous()
    ctx.neigh(api.strided(data.dofs), api.strided(data.buffers), api.strided(data.neigh_int), data.face_types, data.face_rel, n,
              nb_idx=data.neighbors)
    ctx.neigh_dr(api.strided(data.dofs), gp, fsp, sr, data.face_types, n)
    torch.cuda.synchronize()
    assert rel_l2(proxy.points_from_device(ctx, qp.cpu().numpy(), P), qip) < 5e-15
    assert rel_l2(proxy.points_from_device(ctx, d["imposed_plus"].cpu().numpy(), P), imp_plus) < 1e-11
    assert rel_l2(proxy.points_from_device(ctx, d["imposed_minus"].cpu().numpy(), P), imp_minus) < 1e-11
    got = data.to_logical(data.dofs)
    assert rel_l2(got, want) < 1e-11
    assert rel_l2(got - q_loc, want - q_loc) < 1e-9            # the neighbor + nodal-flux update itself
    assert np.abs(d["slip_rate_magnitude"].cpu().numpy()[:, :P] - st_o["slipRateMagnitude"]).max() < 1e-9
