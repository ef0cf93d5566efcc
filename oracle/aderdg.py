"""CPU ORACLE (numpy) for the ADER-DG element update — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may
import this module.  It is a plain restatement of the reference algorithm on
dense batched arrays; every function cites the reference file:line it follows.
Parity status: PINNED — `tests/test_oracle_vs_reference.py` checks every function
here against the reference's own generated CPU kernels (oracle/_ref, built by
oracle/build_ref.py from /root/reference/codegen) and against the golden vectors
committed under tests/golden/ (produced by those same reference kernels).

Logical layout used here: modal tensors are `[n, B, 9]` (basis function k,
quantity p) in the chosen dtype; the reference's memory layout is column-major
`T[k + ld*p]` with a padded leading dimension — see
`to_ref_layout` / `from_ref_layout`.
"""
import math
import os

import numpy as np

TABLE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "seissol_b200", "tables")
NQ = 9


def nbasis(order):
    return order * (order + 1) * (order + 2) // 6


def load_tables(order, dtype=np.float64):
    """Global operator matrices; kDivMT is NEGATED here exactly as the reference
    does once at start-up (src/Memory/GlobalData.cpp:51-62)."""
    z = np.load(os.path.join(TABLE_DIR, f"o{order}.npz"))
    t = {k: z[k] for k in z.files}
    t["kDivMT"] = -t["kDivMT"]
    for k in ("kDivM", "kDivMT", "rDivM", "fMrT", "rT", "fP", "V3mTo2n", "V3mTo2nTWDivM",
              "resample", "quadweights", "V3mTo2nFace", "V2nTo2m"):
        t[k] = t[k].astype(dtype)
    return t


# ----------------------------------------------------------------------------- layouts
def align_ld(rows, itemsize, alignment=64):
    per = alignment // itemsize
    return (rows + per - 1) // per * per


def to_ref_layout(x, ld):
    """[n, B, 9] -> reference memory [n, 9*ld] (column-major, zero padding rows)."""
    n, b, q = x.shape
    out = np.zeros((n, q, ld), dtype=x.dtype)
    out[:, :, :b] = x.transpose(0, 2, 1)
    return out.reshape(n, q * ld)


def from_ref_layout(x, b, ld):
    n = x.shape[0]
    return np.ascontiguousarray(x.reshape(n, -1, ld)[:, :, :b].transpose(0, 2, 1))


def derivative_layout(order, itemsize, alignment=64):
    """Offsets/ld of dQ(d) inside the reference derivative family: dQ(d) keeps only
    its non-zero rows B(O-d), padded (src/Kernels/LinearCK/Time.cpp:98-101,
    codegen/kernels/aderdg.py:507-517)."""
    offs, lds = [], []
    o = 0
    for d in range(order):
        ld = align_ld(nbasis(order - d), itemsize, alignment)
        offs.append(o)
        lds.append(ld)
        o += ld * NQ
    return offs, lds, o


def star_dense(star_vals, pattern):
    """[n,3,24] CSC values -> dense [n,3,9,9] (row q, column p); pattern from
    codegen/matrices/star.xml (LocalIntegrationData.starMatrices, Typedefs.h:187-196)."""
    n = star_vals.shape[0]
    out = np.zeros((n, 3, NQ, NQ), dtype=star_vals.dtype)
    out[:, :, pattern[:, 0], pattern[:, 1]] = star_vals
    return out


def flux_dense(vals):
    """[n,4,81] column-major `A[q + 9p]` -> dense [n,4,9(q),9(p)]."""
    n = vals.shape[0]
    return vals.reshape(n, -1, NQ, NQ).transpose(0, 1, 3, 2)


# ----------------------------------------------------------------------------- time basis
def time_integrate_coeffs(order, t0, t1, dtype=np.float64):
    """MonomialBasis::integrate (src/Numerical/TimeBasis.h:82-96):
    c_d = (t1^(d+1) - t0^(d+1)) / (d+1)!  (expansion point 0)."""
    return np.array([(t1 ** (d + 1) - t0 ** (d + 1)) / math.factorial(d + 1)
                     for d in range(order)], dtype=dtype)


def time_point_coeffs(order, tau, dtype=np.float64):
    """MonomialBasis::point (src/Numerical/TimeBasis.h:71-80): tau^d / d!."""
    return np.array([tau ** d / math.factorial(d) for d in range(order)], dtype=dtype)


def gauss_legendre_shifted(n, a, b):
    """quadrature::ShiftedGaussLegendre (src/Numerical/Quadrature.h:76-90)."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (b - a) * x + 0.5 * (a + b), 0.5 * (b - a) * w


def friction_delta_t(points, dt):
    """FrictionSolver::computeDeltaT (src/DynamicRupture/FrictionLaws/FrictionSolver.cpp:20-37)."""
    d = np.empty(len(points))
    d[0] = points[0]
    d[1:] = np.diff(points)
    d[-1] += dt - points[-1]
    return d


# ----------------------------------------------------------------------------- kernels
def ader(tab, Q, star, coeffs, want_derivatives=False):
    """Cauchy-Kovalevskaya predictor.  Spacetime::computeAder
    (src/Kernels/LinearCK/Time.cpp:65-140) + generated kernel::derivative
    (codegen/kernels/aderdg.py:463-532):
      dQ(0)=Q; dQ(d)=sum_j kDivMT_j . dQ(d-1) . star_j ; I = sum_d c_d dQ(d).
    dQ(d) is non-zero only in its first B(O-d) rows."""
    order = int(tab["order"])
    kt = tab["kDivMT"]
    dq = Q
    I = coeffs[0] * Q
    ders = [Q]
    for d in range(1, order):
        rows_in = nbasis(order - d + 1)
        rows_out = nbasis(order - d)
        new = np.zeros_like(Q)
        for j in range(3):
            tmp = np.matmul(dq[:, :rows_in, :], star[:, j])          # [n, rows_in, 9]
            new[:, :rows_out] += np.matmul(kt[j][:rows_out, :rows_in], tmp)
        I = I + coeffs[d] * new
        dq = new
        ders.append(new)
    return (I, ders) if want_derivatives else I


def taylor(ders, coeffs):
    """Time::evaluate (src/Kernels/LinearCK/Time.cpp:237-255): I = sum_d c_d dQ(d)."""
    out = coeffs[0] * ders[0]
    for d in range(1, len(ders)):
        out = out + coeffs[d] * ders[d]
    return out


def volume(tab, Q, I, star):
    """kernel::volume (codegen/kernels/aderdg.py:321-334; Local.cpp:132-148):
    Q += sum_i kDivM_i . I . star_i"""
    out = Q.copy()
    for i in range(3):
        out += np.matmul(tab["kDivM"][i], np.matmul(I, star[:, i]))
    return out


def local_flux(tab, Q, I, Aplus, face_types=None):
    """kernel::localFlux(f) (aderdg.py:373-385; Local.cpp:150-156):
    Q += rDivM_f . (fMrT_f . I . AplusT_f), skipped on DynamicRupture faces."""
    out = Q.copy()
    for f in range(4):
        upd = np.matmul(tab["rDivM"][f], np.matmul(np.matmul(tab["fMrT"][f], I), Aplus[:, f]))
        if face_types is not None:
            upd = upd * (face_types[:, f] != 3)[:, None, None].astype(Q.dtype)
        out += upd
    return out


def local(tab, Q, I, star, Aplus, face_types=None):
    """Local::computeIntegral (src/Kernels/LinearCK/Local.cpp:124-160)."""
    return local_flux(tab, volume(tab, Q, I, star), I, Aplus, face_types)


def neighbor(tab, Q, Iall, neighbors, Aminus, face_rel, face_types=None,
             dr=None):
    """Neighbor::computeNeighborsIntegral (src/Kernels/LinearCK/Neighbor.cpp:66-109):
    Regular face i: Q += rDivM_i . fP_h . rT_j . I_nb . AminusT_i with
    j = faceRelations[i][0], h = faceRelations[i][1] (kernel::neighboringFlux(h,j,i),
    aderdg.py:414-461).  DynamicRupture face: Q += V3mTo2nTWDivM(side,rel) . godunov .
    fluxSolver (dynamic_rupture.py:174-191); `dr` = dict(godunov [m,P,9], fluxSolver
    [m,9,9], index [n,4], side [n,4], rel [n,4])."""
    out = Q.copy()
    n = Q.shape[0]
    for i in range(4):
        j = face_rel[:, i, 0]
        h = face_rel[:, i, 1]
        Inb = Iall[neighbors[:, i]]
        g = np.matmul(tab["rT"][j], Inb)                 # [n, B2, 9]
        g = np.matmul(tab["fP"][h], g)
        g = np.matmul(g, Aminus[:, i])
        upd = np.matmul(tab["rDivM"][i], g)
        if face_types is not None:
            upd = upd * (face_types[:, i] == 0)[:, None, None].astype(Q.dtype)
        out += upd
        if dr is not None and face_types is not None:
            m = face_types[:, i] == 3
            if m.any():
                idx = dr["index"][m, i]
                mat = tab["V3mTo2nTWDivM"][dr["side"][m, i] + 4 * dr["rel"][m, i]]
                out[m] += np.matmul(mat, np.matmul(dr["godunov"][idx], dr["fluxSolver"][idx]))
    return out


def local_nodal_bc(tab, Q, I, Aminus, bc):
    """Nodal boundary conditions of the local integral (Local::computeIntegral, src/Kernels/LinearCK/Local.cpp:157-252;
    batched :350-455), one boundary face after the other:
      INodal = V3mTo2nFace_f . I . Tinv^T                    (projectToNodalBoundaryRotated, codegen/kernels/nodalbc.py:73-85)
      FaceType::Dirichlet (4): INodal(l, a) = sum_b map(a, b, l) INodal(l, b) + const(a, l)   (createEasiBoundaryGhostCells :55-60)
      FaceType::FreeSurfaceGravity (2): INodal(l, 0..2) = 2 p - INodal(l, 0..2), p = -rho g displacement(l)   (Local.cpp:168-199)
      Q += (rDivM_f . V2nTo2m) . INodal . AminusT_f          (localFluxNodal, codegen/kernels/aderdg.py:336-352)
    bc: dict(cell [m], face [m], kind [m], tinv [m, 9, 9], easiMap [m, 9, 9, B2], easiConst [m, 9, B2],
    displacement [m, B2], rhoG [m]); Aminus dense [n, 4, 9, 9] (flux_dense)."""
    out = Q.copy()
    for i in range(len(bc["cell"])):
        c, f = int(bc["cell"][i]), int(bc["face"][i])
        if bc["kind"][i] not in (2, 4):
            continue          # Local.cpp:247-250: no boundary condition, no nodal flux
        nodal = tab["V3mTo2nFace"][f] @ I[c] @ bc["tinv"][i].T
        if bc["kind"][i] == 4:
            nodal = np.einsum("abl,lb->la", bc["easiMap"][i], nodal) + bc["easiConst"][i].T
        elif bc["kind"][i] == 2:
            p = (-bc["rhoG"][i] * bc["displacement"][i]).astype(Q.dtype)
            nodal = nodal.copy()
            nodal[:, :3] = 2 * p[:, None] - nodal[:, :3]
        out[c] += (tab["rDivM"][f] @ tab["V2nTo2m"]) @ nodal @ Aminus[c, f]
    return out


def analytical_inodal(evaluate, nodes, start_time, time_step_width, order):
    """DirichletBoundary::evaluateTimeDependent (src/Kernels/LinearCK/DirichletBoundary.h:121-165) with
    ApplyAnalyticalSolution (Local.cpp:86-118): INodal_i = sum_j w_j evaluate(nodes_i, t_j), Gauss-Legendre points of
    [t, t + dt].  evaluate(nodes [B2, 3], t) -> [B2, 9]; returns [m, B2, 9]."""
    t, w = gauss_legendre_shifted(order, start_time, start_time + time_step_width)
    return np.stack([sum(wj * np.asarray(evaluate(nd, tj), dtype=np.float64) for tj, wj in zip(t, w)) for nd in nodes])


def local_nodal_flux(tab, Q, Aminus, cell, face, inodal):
    """FaceType::Analytical faces (Local.cpp:233-246, batched :409-452): localFluxNodal on GIVEN nodal values,
    Q += (rDivM_f . V2nTo2m) . INodal . AminusT_f (codegen/kernels/aderdg.py:336-352).  inodal [m, B2, 9]."""
    out = Q.copy()
    for i in range(len(cell)):
        c, f = int(cell[i]), int(face[i])
        out[c] += (tab["rDivM"][f] @ tab["V2nTo2m"]) @ inodal[i] @ Aminus[c, f]
    return out


def tetrahedron_quadrature(n):
    """seissol::quadrature::TetrahedronQuadrature (src/Numerical/Quadrature.h:156-200): conical product of n-point
    Gauss-Jacobi rules with weights (1-x)^2, (1-x), 1 mapped to [0, 1].  Returns points [n^3, 3], weights [n^3]."""
    from scipy.special import roots_jacobi
    p0, w0 = roots_jacobi(n, 2, 0)
    p1, w1 = roots_jacobi(n, 1, 0)
    p2, w2 = roots_jacobi(n, 0, 0)
    p0, p1, p2 = 0.5 * p0 + 0.5, 0.5 * p1 + 0.5, 0.5 * p2 + 0.5
    w0, w1, w2 = 0.125 * w0, 0.25 * w1, 0.5 * w2
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    pts = np.stack([p0[i], p1[j] * (1 - p0[i]), p2[k] * (1 - p1[j]) * (1 - p0[i])], -1)
    return pts, w0[i] * w1[j] * w2[k]


def dubiner_basis(order, pts):
    """Basis functions of the reference tetrahedron at pts [m, 3] -> [m, B] (src/Numerical/Functions.cpp:136-148,
    ordering src/Numerical/BasisFunction.h:135-143): the evalAtQP matrix of the energy output when pts are the
    quadrature points."""
    from scipy.special import eval_jacobi
    xi, eta, zeta = pts[:, 0], pts[:, 1], pts[:, 2]
    r, s, t = 2 * xi - 1 + eta + zeta, 2 * eta - 1 + zeta, 2 * zeta - 1
    st, th = 1 - eta - zeta, 1 - zeta

    def sfj(n, a, x, y):
        yy = np.where(np.abs(y) < 1e-300, 1e-300, y)
        return yy ** n * eval_jacobi(n, a, 0, x / yy)

    idx = [(o - j - k, j, k) for o in range(order) for k in range(o + 1) for j in range(o - k + 1)]
    return np.stack([sfj(i, 0, r, st) * sfj(j, 2 * i + 1, s, th) * eval_jacobi(k, 2 * i + 2 * j + 2, 0, t) for i, j, k in idx], -1)


def volume_energies(order, Q, volume, material):
    """EnergyOutput::computeVolumeEnergies (src/ResultWriter/EnergyOutput.cpp:347-486) without the gravitational and
    plastic parts: Q [n, B, 9] evaluated at the (O+1)^3 quadrature points (kernel::evalAtQP), per point
    kinetic 0.5 rho |v|^2, elastic 0.5 sigma : eps, acoustic p^2 / (2 lambda) for mu = 0, momentum rho v, weight 6 |V| w.
    material [n, 3] = lambda, mu, rho.  Returns [elasticKinetic, elastic, acousticKinetic, acoustic, mx, my, mz, 0]."""
    pts, w = tetrahedron_quadrature(order + 1)
    phi = dubiner_basis(order, pts)
    out = np.zeros(8)
    lookup = [[0, 3, 5], [3, 1, 4], [5, 4, 2]]
    for c in range(Q.shape[0]):
        num = phi @ Q[c].astype(np.float64)            # [qp, 9]
        cw = 6.0 * volume[c] * w
        lam, mu, rho = material[c]
        kin = 0.5 * rho * (num[:, 6] ** 2 + num[:, 7] ** 2 + num[:, 8] ** 2)
        if abs(mu) < 10e-14:
            out[3] += np.sum(cw * num[:, 0] ** 2 / (2 * lam))
            out[2] += np.sum(cw * kin)
            continue
        out[0] += np.sum(cw * kin)
        out[4:7] += rho * (cw @ num[:, 6:9])
        tr = num[:, 0] + num[:, 1] + num[:, 2]
        factor = -lam / (2.0 * mu * (3.0 * lam + 2.0 * mu))
        e = np.zeros(len(w))
        for i in range(3):
            for j in range(3):
                sij = num[:, lookup[i][j]]
                strain = sij / (2.0 * mu) + (factor * tr if i == j else 0.0)
                e += sij * strain
        out[1] += np.sum(cw * 0.5 * e)
    return out


def fault_energies(points, frictional_energy, accumulated_slip, doubled_surface_area, mu_plus, mu_minus, on_rank=None):
    """EnergyOutput::computeDynamicRuptureEnergies (src/ResultWriter/EnergyOutput.cpp:295-324) without the static work:
    per-point arrays [m, ppad] (only the first `points` entries count).  Returns [frictional work, potency, moment]."""
    out = np.zeros(3)
    for i in range(frictional_energy.shape[0]):
        if on_rank is not None and not on_rank[i]:
            continue
        out[0] += np.sum(frictional_energy[i, :points].astype(np.float64))
        mu = 2.0 * mu_plus[i] * mu_minus[i] / (mu_plus[i] + mu_minus[i])
        inc = np.sum(accumulated_slip[i, :points].astype(np.float64)) * 0.5 * float(doubled_surface_area[i]) / points
        out[1] += inc
        out[2] += inc * mu
    return out


def dr_interpolate(tab, ders_plus, ders_minus, tinvT, face_info, coeffs):
    """DynamicRupture::spaceTimeInterpolation (src/Kernels/DynamicRupture.cpp:57-108):
    for each time point s: Q±(t_s) = sum_d coeffs[s,d] dQ±(d);
    QInterp+[s] = V3mTo2n(plusSide,0) . Q+ . TinvT; QInterp-[s] = V3mTo2n(minusSide,rel) . Q- . TinvT.
    Returns arrays [nFaces, O, P, 9]."""
    order = int(tab["order"])
    vp = tab["V3mTo2n"][face_info[:, 0] + 0]
    vm = tab["V3mTo2n"][face_info[:, 1] + 4 * face_info[:, 2]]
    qp, qm = [], []
    for s in range(order):
        qp.append(np.matmul(np.matmul(vp, taylor(ders_plus, coeffs[s])), tinvT))
        qm.append(np.matmul(np.matmul(vm, taylor(ders_minus, coeffs[s])), tinvT))
    return np.stack(qp, 1), np.stack(qm, 1)


# ----------------------------------------------------------------------------- proxy
def proxy_topology(cells, seed=None):
    """Fake topology as Proxy fakeData (src/Proxy/Allocator.cpp:67-87): per cell and
    face draw (side U{0..3}, orientation U{0..2}, neighbor U{0..cells-1}).  The
    reference seeds std::mt19937(cells) and uses libstdc++ uniform_int_distribution;
    we use numpy's generator (documented deviation: same distribution, other stream)."""
    rng = np.random.default_rng(cells if seed is None else seed)
    rel = np.empty((cells, 4, 2), dtype=np.uint8)
    rel[:, :, 0] = rng.integers(0, 4, size=(cells, 4))
    rel[:, :, 1] = rng.integers(0, 3, size=(cells, 4))
    nb = rng.integers(0, cells, size=(cells, 4)).astype(np.int64)
    return rel, nb


def proxy_data(order, cells, dtype, seed=42, scale=1e-2):
    """Bounded synthetic element data: Q ~ U(-1,1); star and flux-solver
    entries ~ U(-1,1)*scale so that the update is a contraction."""
    rng = np.random.default_rng(seed)
    B = nbasis(order)
    d = {
        "Q": rng.uniform(-1, 1, (cells, B, NQ)).astype(dtype),
        "star": (rng.uniform(-1, 1, (cells, 3, 24)) * scale).astype(dtype),
        "Aplus": (rng.uniform(-1, 1, (cells, 4, 81)) * scale).astype(dtype),
        "Aminus": (rng.uniform(-1, 1, (cells, 4, 81)) * scale).astype(dtype),
    }
    d["face_rel"], d["neighbors"] = proxy_topology(cells)
    d["face_types"] = np.zeros((cells, 4), dtype=np.uint8)
    return d


def proxy_step(tab, d, coeffs, kernel="all"):
    """One proxy pass (src/Proxy/KernelHost.cpp): all = local ∘ neigh,
    local = ader ∘ localwoader per cell; `I` is kept in d['I'] (the LTS buffers)."""
    pattern = tab["star_pattern"]
    star = star_dense(d["star"], pattern)
    if kernel in ("all", "local", "ader"):
        d["I"] = ader(tab, d["Q"], star, coeffs)
    if kernel in ("all", "local", "localwoader"):
        d["Q"] = local(tab, d["Q"], d["I"], star, flux_dense(d["Aplus"]), d["face_types"])
    if kernel in ("all", "neigh"):
        d["Q"] = neighbor(tab, d["Q"], d["I"], d["neighbors"], flux_dense(d["Aminus"]),
                          d["face_rel"], d["face_types"])
    return d
