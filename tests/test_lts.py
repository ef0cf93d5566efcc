"""Clustered LTS (rate 2): the scheduler + storage plumbing of seissol_b200/lts.py against the brute-force definition of
LTS (tests/lts_support.py).  CPU: numpy back-end; GPU (-m gpu): the CUDA kernels through the C ABI."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_l2
from lts_support import NumpyLtsBackend, brute_force_macro_step
from oracle import aderdg as O
from seissol_b200 import lts as L
from seissol_b200 import mesh as M
from seissol_b200 import partition as P


def make_problem(order, dt, dims, ncl, scale=1e-2, axis="z"):
    layout = P.build_layout(*dims, 0, 1, ncl, cluster_axis=axis)
    plan = L.build_plan(layout, ncl)
    B = O.nbasis(order)
    Q = M.element_values(plan.gid, 9 * B, 0, 1.0, dt).reshape(plan.n, 9, B).transpose(0, 2, 1).copy()
    li = M.element_values(plan.gid, 396, 1, scale, dt)
    ni = M.element_values(plan.gid, 324, 2, scale, dt)
    return plan, Q, li, ni


@pytest.mark.parametrize("ncl,dims", [(1, (2, 2, 3)), (2, (4, 2, 6)), (3, (2, 2, 10))])
def test_plan_invariants(ncl, dims):
    layout = P.build_layout(*dims, 0, 1, ncl)
    plan = L.build_plan(layout, ncl)
    assert sorted(plan.gid) == sorted(layout.gid[:layout.n_own])
    assert plan.ranges[0].start == 0 and plan.ranges[-1].end == plan.n
    reg = plan.face_types == 0
    d = np.where(reg, plan.cluster[plan.neighbors] - plan.cluster[:, None], 0)
    assert np.abs(d).max() <= 1
    for c, r in enumerate(plan.ranges):
        assert np.all(plan.cluster[r.start:r.end] == c)
        acc = (d[r.start:r.end] == 1).any(1)
        assert acc[:r.acc_end - r.start].all() and not acc[r.acc_end - r.start:].any()
        der = acc | (d[r.start:r.end] == -1).any(1)
        assert der[:r.der_end - r.start].all() and not der[r.der_end - r.start:].any()
    if ncl > 1:
        assert plan.n_acc > 0 and plan.n_scratch > 0 and (plan.kind == 1).any() and (plan.kind == 2).any()
    assert plan.updates_per_macro_step() == sum((r.end - r.start) * 2 ** (ncl - 1 - c) for c, r in enumerate(plan.ranges))


@pytest.mark.parametrize("ncl,dims", [(1, (2, 2, 3)), (2, (2, 2, 6)), (3, (2, 2, 10))])
def test_lts_scheduler_matches_brute_force(ncl, dims):
    order, dt = 4, np.float64
    plan, Q, li, ni = make_problem(order, dt, dims, ncl)
    tab = O.load_tables(order, dt)
    star = O.star_dense(li[:, :72].reshape(-1, 3, 24), tab["star_pattern"])
    Aplus, Aminus = O.flux_dense(li[:, 72:].reshape(-1, 4, 81)), O.flux_dense(ni.reshape(-1, 4, 81))
    be = NumpyLtsBackend(tab, order, plan, Q, star, Aplus, Aminus)
    solver = L.LtsSolver(plan, be, order, dt0=0.5)
    want = Q
    for _ in range(2):
        solver.macro_step()
        want = brute_force_macro_step(tab, order, plan, want, star, Aplus, Aminus, dt0=0.5)
    assert rel_l2(be.Q, want) < 1e-13
    assert rel_l2(be.Q - Q, want - Q) < 1e-11


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _lts_rank(rank, world, port, dims, ncl, steps, out, axis="z"):
    order, dt = 4, np.float64
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    layout = P.build_layout(*dims, rank, world, ncl, cluster_axis=axis)
    plan = L.build_plan(layout, ncl)
    if world > 1 and axis == "x":
        # the slab boundaries separate different clusters: derivative families travel, ghost buffers are accumulated
        assert sum(len(v) for v in plan.send_der.values()) + sum(len(v) for v in plan.recv_der.values()) > 0
        assert sum(len(g) for g, _ in plan.ghost_acc.values()) + sum(len(v) for v in plan.recv_der.values()) > 0
    B = O.nbasis(order)
    Q = M.element_values(plan.gid, 9 * B, 0, 1.0, dt).reshape(plan.n, 9, B).transpose(0, 2, 1).copy()
    li = M.element_values(plan.gid, 396, 1, 1e-2, dt)
    ni = M.element_values(plan.gid, 324, 2, 1e-2, dt)
    tab = O.load_tables(order, dt)
    star = O.star_dense(li[:, :72].reshape(-1, 3, 24), tab["star_pattern"])
    be = NumpyLtsBackend(tab, order, plan, Q, star, O.flux_dense(li[:, 72:].reshape(-1, 4, 81)),
                         O.flux_dense(ni.reshape(-1, 4, 81)), dist=dist)
    solver = L.LtsSolver(plan, be, order, dt0=0.5)
    for _ in range(steps):
        solver.macro_step()
    np.save(os.path.join(out, f"q_{world}_{rank}.npy"), be.Q)
    np.save(os.path.join(out, f"gid_{world}_{rank}.npy"), plan.gid)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,axis", [(2, "z"), (4, "z"), (2, "x"), (4, "x")])
def test_partitioned_lts_matches_single_rank(world, axis, tmp_path):
    """LTS on the x-slab partition (per-cluster halo exchange, gloo) reproduces the single-rank LTS run.  axis "z": the
    remote neighbors share the time cluster (buffers travel); axis "x": every slab boundary separates two clusters
    (derivative families travel to the smaller cluster's side, the larger cluster's side accumulates the ghost buffers:
    DirectGhostTimeCluster.cpp:22-105, LtsSetup.cpp:115-189)."""
    dims, ncl, steps = ((8, 2, 10) if axis == "z" else (8, 2, 2)), 3, 2
    out = str(tmp_path)
    _lts_rank(0, 1, 0, dims, ncl, steps, out, axis)
    mp.spawn(_lts_rank, args=(world, _free_port(), dims, ncl, steps, out, axis), nprocs=world, join=True)
    q1, g1 = np.load(os.path.join(out, "q_1_0.npy")), np.load(os.path.join(out, "gid_1_0.npy"))
    ref = np.empty_like(q1)
    ref[g1] = q1
    seen = 0
    for r in range(world):
        q, g = np.load(os.path.join(out, f"q_{world}_{r}.npy")), np.load(os.path.join(out, f"gid_{world}_{r}.npy"))
        assert rel_l2(q, ref[g]) < 1e-13, r
        seen += len(g)
    assert seen == len(g1)


@pytest.mark.gpu
@pytest.mark.parametrize("order,dt,bar", [(4, np.float32, 1e-5), (6, np.float32, 1e-5), (6, np.float64, 1e-11)])
@pytest.mark.parametrize("faceproj", [False, True])
def test_lts_cuda_matches_brute_force(built_library, order, dt, bar, faceproj):
    """faceproj: the LTS neighbor flux on face projections (accumulated projections for smaller-cluster neighbors,
    ssb_taylor_faceproj for neighbors read through their derivatives)."""
    import torch
    from seissol_b200 import api
    ncl, dims = 3, (2, 2, 10)
    plan, Q, li, ni = make_problem(order, dt, dims, ncl)
    tab = O.load_tables(order, dt)
    star = O.star_dense(li[:, :72].reshape(-1, 3, 24), tab["star_pattern"])
    Aplus, Aminus = O.flux_dense(li[:, 72:].reshape(-1, 4, 81)), O.flux_dense(ni.reshape(-1, 4, 81))
    ctx = api.Context(order, np.dtype(dt).itemsize)
    dev = torch.device("cuda:0")
    qd = np.zeros((plan.n, 9, ctx.ld), dtype=dt)
    qd[:, :, :ctx.B] = Q.transpose(0, 2, 1)
    dofs = torch.from_numpy(qd.reshape(plan.n, 9 * ctx.ld)).to(dev)
    be = L.CudaLtsBackend(torch, ctx, plan, dofs, torch.from_numpy(li).to(dev), torch.from_numpy(ni).to(dev),
                          faceproj=faceproj)
    solver = L.LtsSolver(plan, be, order, dt0=0.5)
    want = Q
    for _ in range(2):
        solver.macro_step()
        want = brute_force_macro_step(tab, order, plan, want, star, Aplus, Aminus, dt0=0.5)
    torch.cuda.synchronize()
    got = dofs.cpu().numpy().reshape(plan.n, 9, ctx.ld)[:, :, :ctx.B].transpose(0, 2, 1)
    assert rel_l2(got, want) < bar
    assert be.launches > 0


def _cuda_lts_rank(rank, world, port, out, faceproj=False, axis="z"):
    """One GPU per rank: partitioned LTS through the C ABI with NCCL; rank-local result saved for the comparison."""
    from seissol_b200 import api
    order, dt, ncl, dims = 4, np.float32, 3, ((8, 2, 10) if axis == "z" else (8, 2, 2))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    layout = P.build_layout(*dims, rank, world, ncl, cluster_axis=axis)
    plan = L.build_plan(layout, ncl)
    ctx = api.Context(order, 4)
    Q = M.element_values(plan.gid, 9 * ctx.B, 0, 1.0, dt).reshape(plan.n, 9, ctx.B)
    qd = np.zeros((plan.n, 9, ctx.ld), dtype=dt)
    qd[:, :, :ctx.B] = Q
    dofs = torch.from_numpy(qd.reshape(plan.n, 9 * ctx.ld)).to(dev)
    li = torch.from_numpy(M.element_values(plan.gid, 396, 1, 1e-2, dt)).to(dev)
    ni = torch.from_numpy(M.element_values(plan.gid, 324, 2, 1e-2, dt)).to(dev)
    be = L.CudaLtsBackend(torch, ctx, plan, dofs, li, ni, dist=dist, faceproj=faceproj)
    solver = L.LtsSolver(plan, be, order, dt0=0.5)
    for _ in range(2):
        solver.macro_step()
    torch.cuda.synchronize()
    got = dofs.cpu().numpy().reshape(plan.n, 9, ctx.ld)[:, :, :ctx.B].transpose(0, 2, 1)
    np.save(os.path.join(out, f"cq_{rank}.npy"), got)
    np.save(os.path.join(out, f"cgid_{rank}.npy"), plan.gid)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("axis", ["z", "x"])
@pytest.mark.parametrize("faceproj", [False, True])
def test_partitioned_lts_cuda_two_gpus(built_library, tmp_path, faceproj, axis):
    """2 GPUs, NCCL: the partitioned LTS run through the CUDA kernels matches the brute-force LTS of the whole mesh
    (faceproj: the halo carries face projections; axis "x": the slab boundaries separate different time clusters, so
    derivative families travel and ghost buffers are accumulated on arrival)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    order, dt, ncl, dims = 4, np.float32, 3, ((8, 2, 10) if axis == "z" else (8, 2, 2))
    out = str(tmp_path)
    mp.spawn(_cuda_lts_rank, args=(2, _free_port(), out, faceproj, axis), nprocs=2, join=True)
    plan, Q, li, ni = make_problem(order, dt, dims, ncl, axis=axis)
    tab = O.load_tables(order, dt)
    star = O.star_dense(li[:, :72].reshape(-1, 3, 24), tab["star_pattern"])
    want = Q
    for _ in range(2):
        want = brute_force_macro_step(tab, order, plan, want, star, O.flux_dense(li[:, 72:].reshape(-1, 4, 81)),
                                      O.flux_dense(ni.reshape(-1, 4, 81)), dt0=0.5)
    ref = np.empty_like(want)
    ref[plan.gid] = want
    for r in range(2):
        q, g = np.load(os.path.join(out, f"cq_{r}.npy")), np.load(os.path.join(out, f"cgid_{r}.npy"))
        assert rel_l2(q, ref[g]) < 1e-5
