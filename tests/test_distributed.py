"""Host logic of the partitioned mesh (no GPU): cube mesh connectivity, slab partition + layer layout,
halo exchange over torch.distributed (gloo, world_size 2 and 4) with the numpy oracle as compute back-end.
The partitioned runs must reproduce the single-rank run of the same mesh."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seissol_b200 import distributed as D
from seissol_b200 import mesh as M
from seissol_b200 import partition as P

ORDER, DT = 4, np.float64


def test_cube_mesh_connectivity():
    m = M.CubeMesh(4, 4, 6)
    nb, rel = m.neighbors, m.face_rel
    has = nb >= 0
    e, f = np.nonzero(has)
    j = rel[e, f, 0]
    assert np.all(nb[nb[e, f], j] == e) and np.all(rel[nb[e, f], j, 0] == f)
    assert (~has).sum() == 2 * 2 * (4 * 6 + 4 * 4)          # y and z domain boundaries, two triangles per cube face
    c = m.vertex_coords()
    x = c[:, :, 0]
    c[:, :, 0] = np.where(x - x.min(1, keepdims=True) > 2, x - 4, x)
    vol = np.linalg.det(c[:, 1:] - c[:, :1]) / 6
    assert np.all(vol > 0) and abs(vol.sum() - 4 * 4 * 6) < 1e-9
    # all 4 x 3 relations occur, and the orientation is symmetric information: both ends agree on the shared vertex
    assert set(np.unique(rel[:, :, 1][has])) == {0, 1, 2} and set(np.unique(rel[:, :, 0][has])) == {0, 1, 2, 3}
    fv = m.vertices[:, M.FACE2NODES]
    mine = np.sort(fv[e, f], axis=1)
    theirs = np.sort(fv[nb[e, f], j], axis=1)
    assert np.array_equal(mine, theirs)


def test_element_values_torch_matches_numpy():
    gid = np.array([0, 1, 5, 123456789, 2 ** 40 + 17], dtype=np.int64)
    for dt, tdt in ((np.float32, torch.float32), (np.float64, torch.float64)):
        a = M.element_values(gid, 37, 2, 0.5, dt)
        b = D.element_values_torch(torch, torch.from_numpy(gid), 37, 2, 0.5, tdt).numpy()
        assert np.array_equal(a, b)
        assert np.all(np.abs(a) <= 0.5) and a.std() > 0.2


@pytest.mark.parametrize("world,dims,ncl", [(2, (4, 4, 4), 1), (2, (8, 4, 6), 2), (4, (8, 2, 4), 1), (4, (16, 4, 8), 3)])
def test_partition_layout(world, dims, ncl):
    lays = [P.build_layout(*dims, r, world, ncl) for r in range(world)]
    assert sum(l.n_own for l in lays) == 5 * dims[0] * dims[1] * dims[2]
    g = P.build_layout(*dims, 0, 1, ncl)
    gmap = {int(k): i for i, k in enumerate(g.gid)}
    for l in lays:
        assert len(l.send) == len(l.recv)
        for s in l.send:
            p = lays[s.peer]
            m = [r for r in p.recv if r.peer == l.rank and r.side == 1 - s.side and r.cluster == s.cluster]
            assert len(m) == 1 and m[0].count == s.count
            assert np.array_equal(l.gid[s.start:s.start + s.count], p.gid[m[0].start:m[0].start + m[0].count])
        assert sum(s.count for s in l.send) == l.n_copy and sum(r.count for r in l.recv) == l.n_ghost
        for i in range(l.n_own):
            gi = gmap[int(l.gid[i])]
            assert [int(g.gid[k]) for k in g.neighbors[gi]] == [int(l.gid[k]) for k in l.neighbors[i]]
            assert np.array_equal(g.face_types[gi], l.face_types[i]) and np.array_equal(g.face_rel[gi], l.face_rel[i])
            assert g.cluster[gi] == l.cluster[i]
        # layers: copy cells first, ordered by cluster; every cluster's ranges tile the own cells
        covered = sorted(r for c in l.cluster_ranges.values() for r in (c["copy"], c["interior"]) if r[1] > r[0])
        assert covered[0][0] == 0 and covered[-1][1] == l.n_own
        assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run_rank(rank, world, port, dims, steps, out):
    from oracle import aderdg as O
    from oracle_backend import OracleBackend
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    B = O.nbasis(ORDER)
    layout, data = D.build_partitioned(torch, dist, *dims, rank, world, B, B + 0, torch.float64, "cpu")
    coeffs = O.time_integrate_coeffs(ORDER, 0.0, 1.0, DT)
    be = OracleBackend(ORDER, data, coeffs, DT)
    x = D.HaloExchange(torch, dist, layout, data.buffers)
    solver = D.PartitionedSolver(torch, layout, data, be, x)
    for _ in range(steps):
        solver.step()
    np.save(os.path.join(out, f"q_{world}_{rank}.npy"), data.logical_dofs())
    np.save(os.path.join(out, f"gid_{world}_{rank}.npy"), layout.gid[:layout.n_own])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_run_matches_single_rank(world, tmp_path):
    dims, steps = (8, 2, 3), 3
    out = str(tmp_path)
    _run_rank(0, 1, 0, dims, steps, out)
    mp.spawn(_run_rank, args=(world, _free_port(), dims, steps, out), nprocs=world, join=True)
    q1 = np.load(os.path.join(out, "q_1_0.npy"))
    g1 = np.load(os.path.join(out, "gid_1_0.npy"))
    ref = np.empty_like(q1)
    ref[g1] = q1
    seen = 0
    for r in range(world):
        q = np.load(os.path.join(out, f"q_{world}_{r}.npy"))
        g = np.load(os.path.join(out, f"gid_{world}_{r}.npy"))
        err = np.linalg.norm(q - ref[g]) / np.linalg.norm(ref[g])
        assert err < 1e-13, (r, err)
        seen += len(g)
    assert seen == len(g1)
    # the update really coupled the ranks: a run without exchange differs
    assert np.linalg.norm(ref - 0) > 0
