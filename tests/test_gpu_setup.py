"""GPU tests (-m gpu) of the device set-up kernels (csrc/kernels_setup.cuh): cell-local matrices (star, AplusT, AminusT)
and fault-face matrices (impedances, TinvT, flux solvers) against the host restatement seissol_b200/material.py, whose
matrices the plane-wave convergence tests (tests/test_planewave.py) validate end to end."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import aderdg as O
from seissol_b200 import material as MAT
from seissol_b200 import mesh as M

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(built_library):
    import torch
    from seissol_b200 import api
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch, api


def _mesh_and_material(seed=2):
    mesh = M.CubeMesh(4, 2, 4)
    rng = np.random.default_rng(seed)
    n = mesh.n
    lam = rng.uniform(2e10, 4e10, n)
    mu = rng.uniform(2e10, 4e10, n)
    rho = rng.uniform(2000.0, 3000.0, n)
    coords = mesh.coords * np.array([1.0, 1.3, 0.8]) + 0.05 * rng.uniform(-1, 1, (1, 1, 3))   # anisotropic cells
    return mesh, coords, lam, mu, rho


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_cell_local_matrices_on_device(gpu, dt):
    torch, api = gpu
    mesh, coords, lam, mu, rho = _mesh_and_material()
    n = mesh.n
    tab = O.load_tables(4, np.float64)
    face_types = np.zeros((n, 4), dtype=np.uint8)
    face_types[::7, 1] = 1      # free surface
    face_types[::11, 2] = 3     # dynamic rupture: no local flux matrix
    nb = mesh.neighbors.astype(np.int64)
    star, ap, am = MAT.cell_local_matrices(coords, nb, lam, mu, rho, tab["star_pattern"], face_types)
    ctx = api.Context(4, np.dtype(dt).itemsize)
    dev = torch.device("cuda")
    tdt = torch.float32 if dt == np.float32 else torch.float64
    li = torch.full((n, 396), float("nan"), device=dev, dtype=tdt)
    ni = torch.full((n, 324), float("nan"), device=dev, dtype=tdt)
    ctx.setup_cell_matrices(torch.from_numpy(coords).to(dev), torch.from_numpy(nb).to(dev),
                            torch.from_numpy(np.stack([lam, mu, rho], 1)).to(dev), torch.from_numpy(face_types).to(dev),
                            api.strided(li), api.strided(ni), n)
    torch.cuda.synchronize()
    li_h, ni_h = li.cpu().numpy().astype(np.float64), ni.cpu().numpy().astype(np.float64)
    t = 1e-13 if dt == np.float64 else 2e-7
    assert rel_l2(li_h[:, :72].reshape(n, 3, 24), star) < t
    assert rel_l2(li_h[:, 72:].reshape(n, 4, 81), ap) < t
    assert rel_l2(ni_h.reshape(n, 4, 81), am) < t
    assert np.all(li_h[::11, 72 + 2 * 81:72 + 3 * 81] == 0)            # the DR face carries no local flux matrix


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_dr_matrices_on_device(gpu, dt):
    torch, api = gpu
    mesh, coords, lam, mu, rho = _mesh_and_material(seed=5)
    rng = np.random.default_rng(9)
    e, f = np.nonzero(mesh.neighbors >= 0)
    pick = rng.choice(len(e), size=57, replace=False)
    plus, pside = e[pick], f[pick]
    minus = mesh.neighbors[plus, pside]
    mside = mesh.face_rel[plus, pside, 0]
    info = np.stack([pside, mside, mesh.face_rel[plus, pside, 1]], 1).astype(np.uint8)
    matp = np.stack([lam[plus], mu[plus], rho[plus]], 1)
    matm = np.stack([lam[minus], mu[minus], rho[minus]], 1)
    tinv_t, imp, fsp, fsm, area2 = MAT.dr_matrices(coords[plus], coords[minus], info, matp, matm)
    m = len(plus)
    ctx = api.Context(4, np.dtype(dt).itemsize)
    dev = torch.device("cuda")
    tdt = torch.float32 if dt == np.float32 else torch.float64
    out = {k: torch.full(shape, float("nan"), device=dev, dtype=tdt)
           for k, shape in (("tinv", (m, 81)), ("imp", (m, 11)), ("fsp", (m, 81)), ("fsm", (m, 81)), ("area", (m,)))}
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    ctx.setup_dr_matrices(d(coords[plus]), d(coords[minus]), d(info), d(matp), d(matm), out["tinv"], out["imp"], out["fsp"],
                          out["fsm"], out["area"], m)
    torch.cuda.synchronize()
    h = {k: v.cpu().numpy().astype(np.float64) for k, v in out.items()}
    t = 1e-13 if dt == np.float64 else 2e-7
    assert rel_l2(h["tinv"].reshape(m, 9, 9).transpose(0, 2, 1), tinv_t) < t      # memory TinvT[a + 9 b]
    assert rel_l2(h["imp"], imp) < t
    assert rel_l2(h["fsp"].reshape(m, 9, 9).transpose(0, 2, 1), fsp) < t          # memory [c + 9 q]
    assert rel_l2(h["fsm"].reshape(m, 9, 9).transpose(0, 2, 1), fsm) < t
    assert rel_l2(h["area"], area2) < t
    # a conforming mesh: both sides see the same face
    assert np.allclose(np.abs(np.linalg.det(tinv_t)), 1.0)
