"""CPU ORACLE (numpy) for the viscoelastic2 (anelastic attenuation) element update — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this module.
Plain restatement of the reference's tensor expressions (codegen/kernels/equations/viscoelastic2.py) and of
the wrappers that call them (src/Kernels/LinearCKAnelastic/{Time,Local,Neighbor}.cpp).
Parity status: PINNED — tests/test_oracle_visco.py checks every function against the reference's own generated
CPU kernels (oracle/_ref/libssref_visco_o5_*.so, built by oracle/build_ref.py) and the golden vectors under
tests/golden/ produced by those kernels.

Logical layouts: Q, I [n, B, 9]; Qane, Iane [n, B, 6, M]; Qext [n, B, 15];
star [n, 3, 33] CSC values of the 9 x 15 pattern (codegen/matrices/matrices_viscoelastic.xml),
Aplus / Aminus [n, 4, 135] column-major A[q + 9 p] (9 x 15), E [n, 162] = E[q + 6 (m + M p)] (6 x M x 9),
w [n, M], W [n, M] (diagonal of the M x M matrix; CSC layout of viscoelastic2.py:92-97).
"""
import numpy as np

from . import aderdg as O

NQ, NA, NE = 9, 6, 15
# star pattern (row q, column p), column-major order of matrices_viscoelastic.xml (0-based)
STAR_PATTERN = np.array(
    [(6, 0), (7, 0), (8, 0), (6, 1), (7, 1), (8, 1), (6, 2), (7, 2), (8, 2), (6, 3), (7, 3), (7, 4), (8, 4), (6, 5), (8, 5),
     (0, 6), (3, 6), (5, 6), (1, 7), (3, 7), (4, 7), (2, 8), (4, 8), (5, 8),
     (6, 9), (7, 10), (8, 11), (6, 12), (7, 12), (7, 13), (8, 13), (6, 14), (8, 14)])


def star_dense(vals):
    """[n, 3, 33] -> [n, 3, 9, 15]."""
    n = vals.shape[0]
    out = np.zeros((n, 3, NQ, NE), dtype=vals.dtype)
    out[:, :, STAR_PATTERN[:, 0], STAR_PATTERN[:, 1]] = vals
    return out


def flux_dense(vals):
    """[n, 4, 135] column-major A[q + 9 p] -> [n, 4, 9 (q), 15 (p)]."""
    n = vals.shape[0]
    return vals.reshape(n, 4, NE, NQ).transpose(0, 1, 3, 2)


def e_dense(vals, mech):
    """[n, 162] E[q + 6 (m + M p)] -> [n, 6 (q), M, 9 (p)]."""
    n = vals.shape[0]
    return vals.reshape(n, NQ, mech, NA).transpose(0, 3, 2, 1)


def ader(tab, Q, Qane, star, E, w, W, coeffs):
    """kernel::derivative of viscoelastic2 (viscoelastic2.py:330-423; Time.cpp:50-120), CPU variant:
      I = c0 Q, Iane = c0 Qane; for d = 1..O-1:
        dQext(d) = sum_j kDivMT_j . dQ(d-1) . star_j                        [B x 15]
        dQ(d)    = dQext(d)[:, :9] + sum_{q,m} dQane(d-1)[:, q, m] E[q, m, :]
        dQane(d) = w[m] dQext(d)[:, 9:] + dQane(d-1)[:, :, m] W[m, m]
        I += c_d dQ(d); Iane += c_d dQane(d)."""
    order = int(tab["order"])
    kt = tab["kDivMT"]
    dq, da = Q, Qane
    I = coeffs[0] * Q
    Iane = coeffs[0] * Qane
    for d in range(1, order):
        ext = np.zeros(Q.shape[:2] + (NE,), dtype=Q.dtype)
        for j in range(3):
            ext += np.matmul(kt[j], np.matmul(dq, star[:, j]))
        new_q = ext[:, :, :NQ] + np.einsum("nkqm,nqmp->nkp", da, E)
        new_a = w[:, None, None, :] * ext[:, :, NQ:, None] + da * W[:, None, None, :]
        I = I + coeffs[d] * new_q
        Iane = Iane + coeffs[d] * new_a
        dq, da = new_q, new_a
    return I, Iane


def local(tab, Q, Qane, I, Iane, star, Aplus, E, w, W, face_types=None):
    """Local::computeIntegral (LinearCKAnelastic/Local.cpp:63-110): volumeExt + localFluxExt(f) into Qext, then
    kernel::local (viscoelastic2.py:259-277):
      Qane += w[m] Qext[:, 9:] + Iane W;  Q += Qext[:, :9] + sum_{q,m} Iane[:, q, m] E[q, m, :]."""
    ext = np.zeros(Q.shape[:2] + (NE,), dtype=Q.dtype)
    for i in range(3):
        ext += np.matmul(tab["kDivM"][i], np.matmul(I, star[:, i]))
    for f in range(4):
        upd = np.matmul(tab["rDivM"][f], np.matmul(np.matmul(tab["fMrT"][f], I), Aplus[:, f]))
        if face_types is not None:
            upd = upd * (face_types[:, f] != 3)[:, None, None].astype(Q.dtype)
        ext += upd
    Qane_new = Qane + w[:, None, None, :] * ext[:, :, NQ:, None] + Iane * W[:, None, None, :]
    Q_new = Q + ext[:, :, :NQ] + np.einsum("nkqm,nqmp->nkp", Iane, E)
    return Q_new, Qane_new


def neighbor(tab, Q, Qane, Iall, neighbors, Aminus, w, face_rel, face_types=None, dr=None):
    """Neighbor::computeNeighborsIntegral (LinearCKAnelastic/Neighbor.cpp:71-130): neighborFluxExt(h, j, i) into
    Qext, then kernel::neighbor (viscoelastic2.py:317-334): Qane += w[m] Qext[:, 9:]; Q += Qext[:, :9].
    DynamicRupture faces (:107-115): Qext += V3mTo2nTWDivM(side, rel) . godunov . fluxSolver (dynamic_rupture.py:174-191
    with the 9 x 15 flux solver of the extended quantities); `dr` = dict(godunov [m, P, 9], fluxSolver [m, 9, 15],
    index [n, 4], side [n, 4], rel [n, 4])."""
    ext = np.zeros(Q.shape[:2] + (NE,), dtype=Q.dtype)
    if dr is not None and face_types is not None:
        for i in range(4):
            m = face_types[:, i] == 3
            if m.any():
                idx = dr["index"][m, i]
                mat = tab["V3mTo2nTWDivM"][dr["side"][m, i] + 4 * dr["rel"][m, i]]
                ext[m] += np.matmul(mat, np.matmul(dr["godunov"][idx], dr["fluxSolver"][idx]))
    for i in range(4):
        j = face_rel[:, i, 0]
        h = face_rel[:, i, 1]
        g = np.matmul(tab["rT"][j], Iall[neighbors[:, i]])
        g = np.matmul(np.matmul(tab["fP"][h], g), Aminus[:, i])
        upd = np.matmul(tab["rDivM"][i], g)
        if face_types is not None:
            upd = upd * (face_types[:, i] == 0)[:, None, None].astype(Q.dtype)
        ext += upd
    return Q + ext[:, :, :NQ], Qane + w[:, None, None, :] * ext[:, :, NQ:, None]


def proxy_data(order, cells, dtype, mech=3, seed=42, scale=1e-2):
    """Bounded synthetic element data."""
    rng = np.random.default_rng(seed)
    B = O.nbasis(order)
    d = {
        "Q": rng.uniform(-1, 1, (cells, B, NQ)).astype(dtype),
        "Qane": rng.uniform(-1, 1, (cells, B, NA, mech)).astype(dtype),
        "star": (rng.uniform(-1, 1, (cells, 3, 33)) * scale).astype(dtype),
        "Aplus": (rng.uniform(-1, 1, (cells, 4, 135)) * scale).astype(dtype),
        "Aminus": (rng.uniform(-1, 1, (cells, 4, 135)) * scale).astype(dtype),
        "E": (rng.uniform(-1, 1, (cells, NA * mech * NQ)) * scale).astype(dtype),
        "w": (rng.uniform(0.1, 1, (cells, mech)) * scale).astype(dtype),
        "W": (-rng.uniform(0.1, 1, (cells, mech)) * scale).astype(dtype),
    }
    d["face_rel"], d["neighbors"] = O.proxy_topology(cells)
    d["face_types"] = np.zeros((cells, 4), dtype=np.uint8)
    return d


def proxy_step(tab, d, coeffs, kernel="all", mech=3):
    """One proxy pass: all = (ader, local) per cell, then neigh (src/Proxy/KernelHost.cpp)."""
    star = star_dense(d["star"])
    E = e_dense(d["E"], mech)
    if kernel in ("all", "local", "ader"):
        d["I"], d["Iane"] = ader(tab, d["Q"], d["Qane"], star, E, d["w"], d["W"], coeffs)
    if kernel in ("all", "local", "localwoader"):
        d["Q"], d["Qane"] = local(tab, d["Q"], d["Qane"], d["I"], d["Iane"], star, flux_dense(d["Aplus"]), E, d["w"],
                                  d["W"], d["face_types"])
    if kernel in ("all", "neigh"):
        d["Q"], d["Qane"] = neighbor(tab, d["Q"], d["Qane"], d["I"], d["neighbors"], flux_dense(d["Aminus"]), d["w"],
                                     d["face_rel"], d["face_types"])
    return d
