"""C++ host above the C ABI (host/ssb_host.hpp, host/proxy_main.cpp -> seissol_b200/lib/ssb_proxy).
CPU: the executable builds, mirrors the reference proxy's command line and refuses to run without a device.
GPU (-m gpu): parity of the whole C++ path (fake data -> kernels::Spacetime/Local/Neighbor -> C ABI -> CUDA kernels)
against the numpy oracle on the dumped inputs."""
import os
import subprocess

import numpy as np
import pytest

from conftest import rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "seissol_b200", "lib", "ssb_proxy")


def test_proxy_cli_without_device(built_library):
    import torch
    assert os.path.exists(EXE)
    r = subprocess.run([EXE, "100", "2", "nonsense"], capture_output=True, text=True)
    assert r.returncode != 0 and "Unknown kernel string: nonsense" in r.stderr       # src/Proxy/Common.cpp:66-78
    r = subprocess.run([EXE, "100"], capture_output=True, text=True)
    assert r.returncode != 0 and "usage" in r.stderr
    if not torch.cuda.is_available():
        r = subprocess.run([EXE, "100", "2", "all"], capture_output=True, text=True)
        assert r.returncode != 0 and "no CUDA device" in r.stderr                      # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("order,prec,dt,bar", [(4, "s", np.float32, 1e-5), (6, "s", np.float32, 1e-5), (6, "d", np.float64, 1e-11)])
def test_cpp_proxy_matches_oracle(built_library, tmp_path, order, prec, dt, bar):
    from oracle import aderdg as O
    cells, steps = 333, 3
    prefix = str(tmp_path / "dump")
    r = subprocess.run([EXE, str(cells), str(steps), "all", "--order", str(order), "--precision", prec, "--dump", prefix,
                        "-f", "json"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert '"name":"all"' in r.stdout and "elem-updates-per-s" in r.stdout
    meta = [int(x) for x in open(prefix + "_meta.txt").read().split()]
    assert meta[:4] == [cells, steps, order, np.dtype(dt).itemsize]
    B, ld = meta[4], meta[5]

    def modal(name):
        x = np.fromfile(prefix + name, dtype=dt).reshape(cells, 9, ld)
        return np.ascontiguousarray(x[:, :, :B].transpose(0, 2, 1))

    li = np.fromfile(prefix + "_localInt.bin", dtype=dt).reshape(cells, 396)
    d = {"Q": modal("_Q0.bin"), "star": li[:, :72].reshape(cells, 3, 24), "Aplus": li[:, 72:].reshape(cells, 4, 81),
         "Aminus": np.fromfile(prefix + "_neighInt.bin", dtype=dt).reshape(cells, 4, 81),
         "face_rel": np.fromfile(prefix + "_faceRel.bin", dtype=np.uint8).reshape(cells, 4, 2),
         "neighbors": np.fromfile(prefix + "_neighbors.bin", dtype=np.int64).reshape(cells, 4),
         "face_types": np.zeros((cells, 4), dtype=np.uint8)}
    assert d["face_rel"][:, :, 0].max() <= 3 and d["face_rel"][:, :, 1].max() <= 2 and d["neighbors"].max() < cells
    tab = O.load_tables(order, dt)
    coeffs = O.time_integrate_coeffs(order, 0.0, 1.0, dt)
    for _ in range(steps + 1):      # one warm-up pass + `timesteps` timed passes (src/Proxy/Runner.cpp:93-111)
        d = O.proxy_step(tab, d, coeffs, "all")
    assert rel_l2(modal("_Q.bin"), d["Q"]) < bar


@pytest.mark.gpu
def test_cpp_proxy_dr_kernels_run(built_library):
    """The dynamic-rupture proxy kernels through the C++ host classes (kernels::DynamicRupture): they run, report the
    reference proxy's fields and leave finite numbers."""
    for kernel in ("godunov_dr", "neigh_dr", "all_dr", "local,godunov_dr,neigh_dr"):
        r = subprocess.run([EXE, "2000", "2", kernel, "--order", "6", "-f", "json"], capture_output=True, text=True)
        assert r.returncode == 0, (kernel, r.stderr)
        assert f'"name":"{kernel}"' in r.stdout and '"time":' in r.stdout
    prefix = "/tmp/ssb_proxy_dr"
    r = subprocess.run([EXE, "500", "1", "all_dr", "--order", "4", "--dump", prefix], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    q = np.fromfile(prefix + "_Q.bin", dtype=np.float32)
    assert np.isfinite(q).all() and np.abs(q).max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("prec,dt,bar", [("s", np.float32, 1e-5), ("d", np.float64, 1e-11)])
def test_cpp_proxy_viscoelastic_matches_oracle(built_library, tmp_path, prec, dt, bar):
    """viscoelastic2 through the C++ anelastic kernel classes (kernels::anelastic::{Spacetime,Local,Neighbor})."""
    from oracle import aderdg as O
    from oracle import visco as V
    cells, steps = 211, 2
    prefix = str(tmp_path / "vdump")
    r = subprocess.run([EXE, str(cells), str(steps), "all", "--equations", "viscoelastic2", "--precision", prec, "--dump", prefix,
                        "-f", "json"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert '"equations":"viscoelastic2"' in r.stdout
    meta = [int(x) for x in open(prefix + "_meta.txt").read().split()]
    B, ld = meta[4], meta[5]
    assert meta[2] == 5 and B == 35

    def q_of(name):
        x = np.fromfile(prefix + name, dtype=dt).reshape(cells, 9, ld)
        return np.ascontiguousarray(x[:, :, :B].transpose(0, 2, 1))

    def ane_of(name):
        x = np.fromfile(prefix + name, dtype=dt).reshape(cells, 3, 6, ld)
        return np.ascontiguousarray(x[:, :, :, :B].transpose(0, 3, 2, 1))

    li = np.fromfile(prefix + "_localInt.bin", dtype=dt).reshape(cells, 807)
    ni = np.fromfile(prefix + "_neighInt.bin", dtype=dt).reshape(cells, 543)
    d = {"Q": q_of("_Q0.bin"), "Qane": ane_of("_Qane0.bin"), "star": li[:, :99].reshape(cells, 3, 33),
         "Aplus": li[:, 99:639].reshape(cells, 4, 135), "E": li[:, 639:801], "w": li[:, 801:804], "W": li[:, 804:807],
         "Aminus": ni[:, :540].reshape(cells, 4, 135),
         "face_rel": np.fromfile(prefix + "_faceRel.bin", dtype=np.uint8).reshape(cells, 4, 2),
         "neighbors": np.fromfile(prefix + "_neighbors.bin", dtype=np.int64).reshape(cells, 4),
         "face_types": np.zeros((cells, 4), dtype=np.uint8)}
    # the neighbor kernel reads w from NeighboringIntegrationData: the proxy's fake data draws it independently
    tab = O.load_tables(5, dt)
    coeffs = O.time_integrate_coeffs(5, 0.0, 1.0, dt)
    star, E = V.star_dense(d["star"]), V.e_dense(d["E"], 3)
    for _ in range(steps + 1):
        I, Ia = V.ader(tab, d["Q"], d["Qane"], star, E, d["w"], d["W"], coeffs)
        d["Q"], d["Qane"] = V.local(tab, d["Q"], d["Qane"], I, Ia, star, V.flux_dense(d["Aplus"]), E, d["w"], d["W"], d["face_types"])
        d["Q"], d["Qane"] = V.neighbor(tab, d["Q"], d["Qane"], I, d["neighbors"], V.flux_dense(d["Aminus"]), ni[:, 540:543],
                                       d["face_rel"], d["face_types"])
    assert rel_l2(q_of("_Q.bin"), d["Q"]) < bar
    assert rel_l2(ane_of("_Qane.bin"), d["Qane"]) < bar


LTS_EXE = os.path.join(ROOT, "seissol_b200", "lib", "ssb_lts_proxy")


def _export(tmp_path, dims, world, ncl, axis):
    from seissol_b200 import partition as P
    prefix = str(tmp_path / "layout")
    layouts = [P.build_layout(*dims, r, world, ncl, cluster_axis=axis) for r in range(world)]
    for r, lay in enumerate(layouts):
        P.export_layout(lay, f"{prefix}_{r}.bin")
    return prefix, layouts


@pytest.mark.parametrize("world,axis,dims", [(1, "z", (2, 2, 10)), (2, "z", (8, 2, 10)), (2, "x", (8, 2, 2)), (4, "x", (8, 2, 2))])
def test_cpp_lts_plan_matches_python(built_library, tmp_path, world, axis, dims):
    """host/ssb_lts.hpp::buildPlan (C++) against seissol_b200/lts.py::build_plan on the same rank layouts: re-ordering,
    cluster ranges, face kinds, scratch / derivative / accumulation rows (incl. the rows of ghost cells) and the sizes of
    the derivative messages.  No device needed."""
    from seissol_b200 import lts as L
    ncl = 3
    prefix, layouts = _export(tmp_path, dims, world, ncl, axis)
    for r, lay in enumerate(layouts):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world))
        out = str(tmp_path / "plan")
        res = subprocess.run([LTS_EXE, "--layout", prefix, "--plan-only", out], capture_output=True, text=True, env=env)
        assert res.returncode == 0, res.stderr
        plan = L.build_plan(lay, ncl)
        x = np.fromfile(f"{out}_{r}_plan.bin", dtype=np.int64)
        n, ng = plan.n, plan.n_ghost
        head = 6 + 3 * ncl
        assert list(x[:6]) == [n, ng, ncl, plan.n_der, plan.n_acc, plan.n_scratch]
        for c in range(ncl):
            assert x[6 + 3 * c] == sum(len(rows) for _, _, rows in plan.send_der[c])
            assert x[7 + 3 * c] == sum(cnt for _, _, _, cnt in plan.recv_der[c])
            assert x[8 + 3 * c] == len(plan.ghost_acc[c][0])
        pos = head
        for want in (plan.gid, plan.cluster, plan.neighbors.ravel(), plan.kind.ravel(), plan.scratch_row.ravel(),
                     plan.der_row, plan.acc_row):
            got = x[pos:pos + want.size]
            assert np.array_equal(got, want.astype(np.int64))
            pos += want.size
        ranges = x[pos:pos + 4 * ncl].reshape(ncl, 4)
        for c, rg in enumerate(plan.ranges):
            assert list(ranges[c]) == [rg.start, rg.acc_end, rg.der_end, rg.end]


def _run_cpp_lts(tmp_path, dims, world, ncl, axis, order, prec, faceproj, steps):
    prefix, layouts = _export(tmp_path, dims, world, ncl, axis)
    dump = str(tmp_path / "lts")
    idfile = str(tmp_path / "nccl_id")
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([LTS_EXE, "--layout", prefix, "--order", str(order), "--precision", prec, "--steps", str(steps),
                                       "--dt0", "0.5", "--faceproj", str(int(faceproj)), "--dump", dump, "--idfile", idfile],
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se
        assert '"name":"lts"' in so
    return dump


def _lts_reference(order, dt, dims, ncl, axis, steps):
    """The brute-force definition of LTS on the whole mesh (tests/lts_support.py), indexed by global element id."""
    from lts_support import brute_force_macro_step
    from oracle import aderdg as O
    from seissol_b200 import lts as L
    from seissol_b200 import mesh as M
    from seissol_b200 import partition as P
    plan = L.build_plan(P.build_layout(*dims, 0, 1, ncl, cluster_axis=axis), ncl)
    B = O.nbasis(order)
    Q = M.element_values(plan.gid, 9 * B, 0, 1.0, dt).reshape(plan.n, 9, B).transpose(0, 2, 1).copy()
    li = M.element_values(plan.gid, 396, 1, 1e-2, dt)
    ni = M.element_values(plan.gid, 324, 2, 1e-2, dt)
    tab = O.load_tables(order, dt)
    star = O.star_dense(li[:, :72].reshape(-1, 3, 24), tab["star_pattern"])
    Aplus, Aminus = O.flux_dense(li[:, 72:].reshape(-1, 4, 81)), O.flux_dense(ni.reshape(-1, 4, 81))
    want = Q
    for _ in range(steps):
        want = brute_force_macro_step(tab, order, plan, want, star, Aplus, Aminus, dt0=0.5)
    ref = np.empty_like(want)
    ref[plan.gid] = want
    return ref


@pytest.mark.gpu
@pytest.mark.parametrize("order,prec,dt,bar,faceproj", [(4, "s", np.float32, 1e-5, False), (6, "s", np.float32, 1e-5, True),
                                                         (6, "d", np.float64, 1e-11, True)])
def test_cpp_lts_single_rank_matches_brute_force(built_library, tmp_path, order, prec, dt, bar, faceproj):
    """The C++ LTS host (plan, rate-2 schedule, back-end over the C ABI) on one GPU against the brute-force LTS."""
    dims, ncl, steps = (2, 2, 10), 3, 2
    dump = _run_cpp_lts(tmp_path, dims, 1, ncl, "z", order, prec, faceproj, steps)
    ref = _lts_reference(order, dt, dims, ncl, "z", steps)
    gid = np.fromfile(dump + "_0_gid.bin", dtype=np.int64)
    B = ref.shape[1]
    q = np.fromfile(dump + "_0_Q.bin", dtype=dt).reshape(len(gid), 9, -1)[:, :, :B].transpose(0, 2, 1)
    assert rel_l2(q, ref[gid]) < bar


@pytest.mark.gpu
@pytest.mark.parametrize("axis,faceproj", [("z", True), ("x", False), ("x", True)])
def test_cpp_lts_two_gpus_nccl(built_library, tmp_path, axis, faceproj):
    """Two processes, two GPUs, ncclSend/ncclRecv halo (axis "x": derivative families travel, ghost buffers are
    accumulated on arrival) against the brute-force LTS of the whole mesh."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    order, dt, ncl, steps = 4, np.float32, 3, 2
    dims = (8, 2, 10) if axis == "z" else (8, 2, 2)
    dump = _run_cpp_lts(tmp_path, dims, 2, ncl, axis, order, "s", faceproj, steps)
    ref = _lts_reference(order, dt, dims, ncl, axis, steps)
    B = ref.shape[1]
    for r in range(2):
        gid = np.fromfile(f"{dump}_{r}_gid.bin", dtype=np.int64)
        q = np.fromfile(f"{dump}_{r}_Q.bin", dtype=dt).reshape(len(gid), 9, -1)[:, :, :B].transpose(0, 2, 1)
        assert rel_l2(q, ref[gid]) < 1e-5, r
