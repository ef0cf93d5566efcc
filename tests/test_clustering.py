"""Time clusters from the cells' CFL time steps (seissol_b200/clustering.py) and the LTS scheme on them (CPU)."""
import os

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import rel_l2
from lts_support import NumpyLtsBackend, brute_force_macro_step
from oracle import aderdg as O
from seissol_b200 import clustering as C
from seissol_b200 import lts as L
from seissol_b200 import mesh as M
from seissol_b200 import partition as P
from test_lts import _free_port


def test_insphere_and_cell_timestep():
    a = 2.0
    reg = np.array([[[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]]], dtype=np.float64) * (a / (2 * np.sqrt(2)))
    assert np.allclose(C.insphere_radius(reg), a * np.sqrt(6) / 12)            # regular tetrahedron of edge a
    unit = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]], dtype=np.float64)
    assert np.allclose(C.insphere_radius(unit), 1.0 / (3 + np.sqrt(3)))        # r = 3 V / A
    dt = C.cell_timestep(unit, 4000.0, order=6, cfl=0.5)
    assert np.allclose(dt, 0.5 * 2 / (3 + np.sqrt(3)) / (4000.0 * 11))         # GlobalTimestep.cpp:45-46
    assert np.allclose(C.cell_timestep(unit, 4000.0, 6, 0.5, max_timestep=1e-9), 1e-9)


def layered(nx=4, ny=2, nz=12, order=4):
    """A cube mesh whose P-wave speed grows 8-fold with depth: the time steps span three rate-2 clusters."""
    mesh = M.CubeMesh(nx, ny, nz)
    depth = (nz - 1 - mesh.cube_z) / (nz - 1)
    vp = 1000.0 * 8.0 ** depth
    dt = C.cell_timestep(mesh.coords, vp, order)
    return mesh, dt


def test_clustering_properties_and_cost():
    mesh, dt = layered()
    nb, reg = mesh.neighbors, mesh.neighbors >= 0
    raw, base = C.cluster_ids(dt, 2)
    assert raw.min() == 0 and raw.max() >= 2
    assert np.all(dt >= base * 2.0 ** raw * (1 - 1e-12)) and np.all(dt < base * 2.0 ** (raw + 1))
    ids, sweeps = C.enforce_maximum_difference(raw, nb, reg)
    d = np.where(reg, ids[np.maximum(nb, 0)] - ids[:, None], 0)
    assert np.abs(d).max() <= 1 and np.all(ids <= raw)                        # only ever lowered: still stable
    best = C.best_clustering(dt, nb, reg, wiggle_min=0.6, wiggle_step=0.05, **C.CPU_COST)
    assert best["cost"] <= C.clustering_cost(ids, base) * (1 + 1e-12)          # the wiggle factor never makes it worse
    assert best["cost"] < 0.75 * C.gts_cost(dt)                                # and LTS pays off on this mesh
    d = np.where(reg, best["cluster"][np.maximum(nb, 0)] - best["cluster"][:, None], 0)
    assert np.abs(d).max() <= 1
    assert np.all(dt >= best["base_dt"] * 2.0 ** best["cluster"] * (1 - 1e-12))
    # the GPU cost terms (a cluster update costs its launches, small clusters do not fill the device) make a clustering of
    # this tiny mesh pointless: everything merges into one cluster, at the largest admissible base time step
    gpu = C.best_clustering(dt, nb, reg, wiggle_min=0.6, wiggle_step=0.05, **C.GPU_COST)
    assert gpu["n_clusters"] == 1 and gpu["wiggle"] == 1.0
    # ... while a mesh of the bench's size keeps its clusters under the same terms (weights scale the cell counts)
    w = np.full(mesh.n, 20000.0)
    full = [C.clustering_cost(ids, base, weights=w, **C.GPU_COST), C.clustering_cost(np.zeros_like(ids), base, weights=w, **C.GPU_COST)]
    assert full[0] < full[1]


def _velocity_lts_rank(rank, world, port, out, ids, dims, ncl):
    import torch.distributed as dist
    order, dt = 4, np.float64
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = L.build_plan(P.build_layout(*dims, rank, world, ncl, cluster_by_gid=ids), ncl)
    B = O.nbasis(order)
    Q = M.element_values(plan.gid, 9 * B, 0, 1.0, dt).reshape(plan.n, 9, B).transpose(0, 2, 1).copy()
    li = M.element_values(plan.gid, 396, 1, 1e-2, dt)
    ni = M.element_values(plan.gid, 324, 2, 1e-2, dt)
    tab = O.load_tables(order, dt)
    star = O.star_dense(li[:, :72].reshape(-1, 3, 24), tab["star_pattern"])
    Aplus, Aminus = O.flux_dense(li[:, 72:].reshape(-1, 4, 81)), O.flux_dense(ni.reshape(-1, 4, 81))
    be = NumpyLtsBackend(tab, order, plan, Q, star, Aplus, Aminus, dist=dist if world > 1 else None)
    solver = L.LtsSolver(plan, be, order, dt0=0.5)
    for _ in range(2):
        solver.macro_step()
    if world == 1:
        want = Q
        for _ in range(2):
            want = brute_force_macro_step(tab, order, plan, want, star, Aplus, Aminus, dt0=0.5)
        assert rel_l2(be.Q, want) < 1e-13
    np.save(os.path.join(out, f"vq_{world}_{rank}.npy"), be.Q)
    np.save(os.path.join(out, f"vgid_{world}_{rank}.npy"), plan.gid)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_lts_on_velocity_clusters(tmp_path):
    """Clusters from the CFL time steps of a layered medium (not aligned with the cubes: the tetrahedra of one cube
    differ in their insphere radius) drive the LTS scheme: the single-rank run equals the brute-force LTS, and the
    2-rank run equals the single-rank run."""
    dims = (4, 2, 12)
    mesh, dt = layered(*dims)
    best = C.best_clustering(dt, mesh.neighbors, mesh.neighbors >= 0, max_clusters=3, **C.CPU_COST)
    ids = np.empty(mesh.n, dtype=np.int64)
    ids[mesh.gid] = best["cluster"]
    ncl = best["n_clusters"]
    assert ncl == 3
    out = str(tmp_path)
    _velocity_lts_rank(0, 1, 0, out, ids, dims, ncl)
    mp.spawn(_velocity_lts_rank, args=(2, _free_port(), out, ids, dims, ncl), nprocs=2, join=True)
    q1, g1 = np.load(os.path.join(out, "vq_1_0.npy")), np.load(os.path.join(out, "vgid_1_0.npy"))
    ref = np.empty_like(q1)
    ref[g1] = q1
    for r in range(2):
        q, g = np.load(os.path.join(out, f"vq_2_{r}.npy")), np.load(os.path.join(out, f"vgid_2_{r}.npy"))
        assert rel_l2(q, ref[g]) < 1e-13
