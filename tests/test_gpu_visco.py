"""GPU parity tests (-m gpu) of the viscoelastic2 kernels, through the C ABI, against the numpy oracle
(oracle/visco.py, pinned against the reference's generated kernels) and the committed golden vectors.
Tolerances: BASELINE.json's bar for DOFs after N steps is 1e-5 (single) / 1e-11 (double) relative L2; single kernels
are held to 2e-6 / 5e-15."""
import os

import numpy as np
import pytest

from conftest import rel_l2
from oracle import aderdg as O
from oracle import visco as V

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def tol(dt):
    return 2e-6 if dt == np.float32 else 5e-15


@pytest.fixture(scope="module")
def gpu(built_library):
    import torch
    from seissol_b200 import api, proxy
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch, api, proxy


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 13, 200])
def test_visco_kernels_match_oracle(gpu, dt, n):
    torch, api, proxy = gpu
    ctx = api.Context(5, np.dtype(dt).itemsize, equation=1, mechanisms=3)
    tab = O.load_tables(5, dt)
    d = V.proxy_data(5, n, dt, seed=100 + n)
    d["face_types"][::5, 1] = 3
    d["face_types"][::7, 2] = 5
    coeffs = O.time_integrate_coeffs(5, 0.0, 1.0, dt)
    star, E = V.star_dense(d["star"]), V.e_dense(d["E"], 3)
    data = proxy.ViscoProxyData.from_host(ctx, d)
    k = proxy.ViscoProxyKernels(ctx, data)
    k.ader()
    torch.cuda.synchronize()
    I, Ia = V.ader(tab, d["Q"], d["Qane"], star, E, d["w"], d["W"], coeffs)
    assert rel_l2(data.q_to_logical(data.buffers), I) < tol(dt)
    assert rel_l2(data.ane_to_logical(data.buffers_ane), Ia) < tol(dt)
    k.localwoader()
    torch.cuda.synchronize()
    ql, qal = V.local(tab, d["Q"], d["Qane"], I, Ia, star, V.flux_dense(d["Aplus"]), E, d["w"], d["W"], d["face_types"])
    assert rel_l2(data.q_to_logical(data.dofs), ql) < tol(dt)
    assert rel_l2(data.ane_to_logical(data.dofs_ane), qal) < tol(dt)
    assert rel_l2(data.q_to_logical(data.dofs) - d["Q"], ql - d["Q"]) < 50 * tol(dt)
    data.sort_by_face_relation()
    k.neigh()
    torch.cuda.synchronize()
    qn, qan = V.neighbor(tab, ql, qal, I, d["neighbors"], V.flux_dense(d["Aminus"]), d["w"], d["face_rel"], d["face_types"])
    assert rel_l2(data.q_to_logical(data.dofs), qn) < tol(dt)
    assert rel_l2(data.ane_to_logical(data.dofs_ane), qan) < tol(dt)
    assert rel_l2(data.q_to_logical(data.dofs) - ql, qn - ql) < 50 * tol(dt)


@pytest.mark.parametrize("cfg,dt", [("visco_o5_f32", np.float32), ("visco_o5_f64", np.float64)])
def test_visco_matches_reference_golden(gpu, cfg, dt):
    torch, api, proxy = gpu
    g = dict(np.load(os.path.join(GOLDEN, f"{cfg}.npz")))
    ctx = api.Context(5, np.dtype(dt).itemsize, equation=1, mechanisms=3)
    data = proxy.ViscoProxyData.from_host(ctx, g)
    k = proxy.ViscoProxyKernels(ctx, data)
    k.coeffs = g["coeffs"]
    k.ader()
    k.localwoader()
    torch.cuda.synchronize()
    assert rel_l2(data.q_to_logical(data.buffers), g["I"]) < tol(dt)
    assert rel_l2(data.ane_to_logical(data.buffers_ane), g["Iane"]) < tol(dt)
    assert rel_l2(data.q_to_logical(data.dofs), g["Q_local"]) < tol(dt)
    assert rel_l2(data.ane_to_logical(data.dofs_ane), g["Qane_local"]) < tol(dt)
    k.neigh()
    torch.cuda.synchronize()
    assert rel_l2(data.q_to_logical(data.dofs), g["Q_all"]) < tol(dt)
    assert rel_l2(data.ane_to_logical(data.dofs_ane), g["Qane_all"]) < tol(dt)


@pytest.mark.parametrize("dt,bar", [(np.float32, 1e-5), (np.float64, 1e-11)])
def test_visco_all_ten_steps(gpu, dt, bar):
    """BASELINE.json north_star: DOFs after N time steps within 1e-5 (single) / 1e-11 (double) relative L2."""
    torch, api, proxy = gpu
    ctx = api.Context(5, np.dtype(dt).itemsize, equation=1, mechanisms=3)
    tab = O.load_tables(5, dt)
    d = V.proxy_data(5, 300, dt, seed=9)
    coeffs = O.time_integrate_coeffs(5, 0.0, 1.0, dt)
    data = proxy.ViscoProxyData.from_host(ctx, d)
    data.sort_by_face_relation()
    k = proxy.ViscoProxyKernels(ctx, data)
    for _ in range(10):
        k.all()
        d = V.proxy_step(tab, d, coeffs, "all")
    torch.cuda.synchronize()
    assert rel_l2(data.q_to_logical(data.dofs), d["Q"]) < bar
    assert rel_l2(data.ane_to_logical(data.dofs_ane), d["Qane"]) < bar


def test_visco_context_rejects_unsupported():
    from seissol_b200 import api
    with pytest.raises(api.SsbError):
        api.Context(6, 4, equation=1, mechanisms=3)
    with pytest.raises(api.SsbError):
        api.Context(5, 4, equation=1, mechanisms=2)


def test_visco_empty_batch_and_reference_layout(gpu):
    """n = 0 is a no-op for every entry point; the reference's 64-byte padded layout (ld = 48 in single precision at
    order 5, yateto/arch.py:197-198) gives the same numbers as the compact one."""
    torch, api, proxy = gpu
    dt = np.float32
    tab = O.load_tables(5, dt)
    d = V.proxy_data(5, 40, dt, seed=4)
    coeffs = O.time_integrate_coeffs(5, 0.0, 1.0, dt)
    ctx = api.Context(5, 4, equation=1, mechanisms=3)
    data = proxy.ViscoProxyData.from_host(ctx, d)
    before = data.dofs.clone()
    ctx.visco_ader(api.strided(data.dofs), api.strided(data.dofs_ane), api.strided(data.buffers), api.strided(data.buffers_ane),
                   api.strided(data.local_int), coeffs, 0)
    ctx.visco_local(api.strided(data.dofs), api.strided(data.dofs_ane), api.strided(data.buffers), api.strided(data.buffers_ane),
                    api.strided(data.local_int), data.face_types, 0)
    ctx.visco_neigh(api.strided(data.dofs), api.strided(data.dofs_ane), api.strided(data.buffers), api.strided(data.neigh_int),
                    data.face_types, data.face_rel, 0, nb_idx=data.neighbors)
    torch.cuda.synchronize()
    assert torch.equal(before, data.dofs)
    ref = api.Context(5, 4, equation=1, mechanisms=3, reference_layout=True)
    assert ref.ld == 48 and ctx.ld == 36
    data_r = proxy.ViscoProxyData.from_host(ref, d)
    for dd, cc in ((data, ctx), (data_r, ref)):
        k = proxy.ViscoProxyKernels(cc, dd)
        k.all()
    torch.cuda.synchronize()
    star, E = V.star_dense(d["star"]), V.e_dense(d["E"], 3)
    want = V.proxy_step(tab, dict(d), coeffs, "all")
    for dd in (data, data_r):
        assert rel_l2(dd.q_to_logical(dd.dofs), want["Q"]) < tol(dt)
        assert rel_l2(dd.ane_to_logical(dd.dofs_ane), want["Qane"]) < tol(dt)
    del star, E


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n,m", [(5, 7), (211, 300)])
def test_visco_neigh_dr_matches_oracle(gpu, dt, n, m):
    """ssb_visco_neigh + ssb_visco_neigh_dr = Neighbor::computeNeighborsIntegral of the viscoelastic2 build with
    DynamicRupture faces (LinearCKAnelastic/Neighbor.cpp:71-130): regular faces and the nodal flux of the imposed
    fault state against the oracle (whose DR branch is pinned against the reference's generated kernels)."""
    from test_oracle_visco import visco_dr_case
    torch, api, proxy = gpu
    ctx = api.Context(5, np.dtype(dt).itemsize, equation=1, mechanisms=3)
    tab = O.load_tables(5, dt)
    rng = np.random.default_rng(31)
    P = 36
    assert ctx.P == P and ctx.ppad >= P
    d = V.proxy_data(5, n, dt, seed=200 + n)
    d["face_types"], dr = visco_dr_case(rng, n, m, P, dt)
    I = rng.uniform(-1, 1, d["Q"].shape).astype(dt)
    d["I"] = I
    data = proxy.ViscoProxyData.from_host(ctx, d)
    dev = data.dofs.device
    es = np.dtype(dt).itemsize
    god = np.zeros((m, 9, ctx.ppad), dtype=dt)
    god[:, :, :P] = dr["godunov"].transpose(0, 2, 1)
    god_d = torch.from_numpy(god).to(dev)
    fs_d = torch.from_numpy(np.ascontiguousarray(dr["fluxSolver"].transpose(0, 2, 1)).reshape(m, 135)).to(dev)
    is_dr = d["face_types"] == 3
    gp = np.where(is_dr, god_d.data_ptr() + dr["index"] * 9 * ctx.ppad * es, 0).astype(np.int64)
    fp = np.where(is_dr, fs_d.data_ptr() + dr["index"] * 135 * es, 0).astype(np.int64)
    sr = np.stack([dr["side"], dr["rel"]], -1).astype(np.uint8)
    k = proxy.ViscoProxyKernels(ctx, data)
    k.neigh()
    ctx.visco_neigh_dr(api.strided(data.dofs), api.strided(data.dofs_ane), api.strided(data.neigh_int),
                       torch.from_numpy(gp).to(dev), torch.from_numpy(fp).to(dev), torch.from_numpy(sr).to(dev),
                       data.face_types, n)
    torch.cuda.synchronize()
    qn, qan = V.neighbor(tab, d["Q"], d["Qane"], I, d["neighbors"], V.flux_dense(d["Aminus"]), d["w"], d["face_rel"],
                         d["face_types"], dr=dr)
    q0, qa0 = V.neighbor(tab, d["Q"], d["Qane"], I, d["neighbors"], V.flux_dense(d["Aminus"]), d["w"], d["face_rel"],
                         d["face_types"])
    assert np.linalg.norm(qn - q0) > 1e-3 * np.linalg.norm(q0)      # the fault faces do contribute
    assert rel_l2(data.q_to_logical(data.dofs), qn) < tol(dt)
    assert rel_l2(data.ane_to_logical(data.dofs_ane), qan) < tol(dt)
    assert rel_l2(data.q_to_logical(data.dofs) - q0, qn - q0) < 50 * tol(dt)
