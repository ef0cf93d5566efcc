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
