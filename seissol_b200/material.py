"""Cell-local matrices of the elastic wave equation from geometry and material (host logic, numpy): the star matrices and
the flux solvers AplusT / AminusT that the element-update kernels consume.

Restated from src/Initializer/Model/CellLocalMatrices.cpp:66-298 (star = sum_d grad[d] M_d^T; AplusT = fluxScale Tinv^T
(QgodLocal A~^T) T^T, fluxScale = -2 |S| / (6 |V|)), src/Equations/elastic/Model/ElasticSetup.h:28-215 (transposed
coefficient matrices, transposed Godunov state from the eigenvector basis, face rotation), src/Numerical/
Transformation.cpp:66-282 (reference Jacobian, rotation matrices), src/Geometry/MeshTools.cpp:44-132 (normals, areas).
Quantity order: sxx syy szz sxy syz sxz u v w."""
import numpy as np

FACE2NODES = np.array([[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]])


def coefficient_matrix_T(lam, mu, rho, dim):
    """getTransposedCoefficientMatrix (ElasticSetup.h:28-78): M(row, col), elastic (mu > 0)."""
    m = np.zeros((9, 9))
    l2m, ri = lam + 2.0 * mu, 1.0 / rho
    if dim == 0:
        m[6, 0], m[6, 1], m[6, 2], m[7, 3], m[8, 5] = -l2m, -lam, -lam, -mu, -mu
        m[0, 6], m[3, 7], m[5, 8] = -ri, -ri, -ri
    elif dim == 1:
        m[7, 0], m[7, 1], m[7, 2], m[6, 3], m[8, 4] = -lam, -l2m, -lam, -mu, -mu
        m[1, 7], m[3, 6], m[4, 8] = -ri, -ri, -ri
    else:
        m[8, 0], m[8, 1], m[8, 2], m[7, 4], m[6, 5] = -lam, -lam, -l2m, -mu, -mu
        m[2, 8], m[5, 6], m[4, 7] = -ri, -ri, -ri
    return m


def godunov_state_T(lam, mu, rho, lam_n, mu_n, rho_n):
    """getTransposedGodunovState (ElasticSetup.h:80-180), elastic-elastic: returns (QgodLocal, QgodNeighbor)."""
    R = np.zeros((9, 9))
    R[0, 0], R[1, 0], R[2, 0], R[6, 0] = lam + 2 * mu, lam, lam, np.sqrt((lam + 2 * mu) / rho)
    R[3, 1], R[7, 1] = mu, np.sqrt(mu / rho)
    R[5, 2], R[8, 2] = mu, np.sqrt(mu / rho)
    R[4, 3] = R[1, 4] = R[2, 5] = lam + 2 * mu
    R[5, 6], R[8, 6] = mu_n, -np.sqrt(mu_n / rho_n)
    R[3, 7], R[7, 7] = mu_n, -np.sqrt(mu_n / rho_n)
    R[0, 8], R[1, 8], R[2, 8], R[6, 8] = lam_n + 2 * mu_n, lam_n, lam_n, -np.sqrt((lam_n + 2 * mu_n) / rho_n)
    chi = np.zeros((9, 9))
    chi[0, 0] = chi[1, 1] = chi[2, 2] = 1.0
    god = np.linalg.solve(R.T, chi @ R.T)
    return np.eye(9) - god, god


def free_surface_godunov_state_T(lam, mu, rho):
    """getTransposedFreeSurfaceGodunovState (src/Model/Common.h:319-364), elastic: the traction components (0, 3, 5 in
    the face-aligned frame) are mirrored into the velocities, QgodLocal(t, v) = S(v, t), QgodLocal(v, v) = 1,
    S = -R21 R11^-1 with the outgoing eigenvectors of the local material."""
    cp, cs = np.sqrt((lam + 2 * mu) / rho), np.sqrt(mu / rho)
    r11 = np.diag([lam + 2 * mu, mu, mu])
    r21 = np.diag([cp, cs, cs])
    S = -(r21 @ np.linalg.inv(r11))
    q = np.zeros((9, 9))
    for col, t in enumerate((0, 3, 5)):
        for row, v in enumerate((6, 7, 8)):
            q[t, v] = S[row, col]
    for v in (6, 7, 8):
        q[v, v] = 1.0
    return q


def rotation_matrices(n, s, t):
    """getFaceRotationMatrix (ElasticSetup.h:182-198): T and Tinv for the normal / tangent basis."""
    T, Ti = np.zeros((9, 9)), np.zeros((9, 9))
    v = [n, s, t]
    pairs = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2)]       # Voigt order of this code: xx yy zz xy yz xz
    for col, (a, b) in enumerate(pairs):
        for row, (i, j) in enumerate(pairs):
            # global component (i, j) of the dyad basis a (x) b of the local frame (symmetrised)
            val = v[a][i] * v[b][j] if a == b else v[a][i] * v[b][j] + v[b][i] * v[a][j]
            T[row, col] = val
    for a in range(3):
        for i in range(3):
            T[6 + i, 6 + a] = v[a][i]
            Ti[6 + a, 6 + i] = v[a][i]
    for row, (a, b) in enumerate(pairs):
        for col, (i, j) in enumerate(pairs):
            val = v[a][i] * v[b][j] if i == j else v[a][i] * v[b][j] + v[a][j] * v[b][i]
            Ti[row, col] = val
    return T, Ti


def reference_jacobian(c):
    """tetrahedronGlobalToReferenceJacobian: rows grad xi, grad eta, grad zeta of one element ([4, 3] coordinates)."""
    return np.linalg.inv((c[1:] - c[:1]).T)


def cell_local_matrices(coords, neighbors, lam, mu, rho, star_pattern, face_types=None):
    """coords [n, 4, 3], neighbors [n, 4] (index or < 0), per-cell lam / mu / rho [n].
    Returns star [n, 3, 24] (CSC values), Aplus [n, 4, 81], Aminus [n, 4, 81] (column-major A[i + 9 j])."""
    n = coords.shape[0]
    star = np.zeros((n, 3, len(star_pattern)))
    aplus = np.zeros((n, 4, 81))
    aminus = np.zeros((n, 4, 81))
    for e in range(n):
        c = coords[e]
        grad = reference_jacobian(c)
        mats = [coefficient_matrix_T(lam[e], mu[e], rho[e], d) for d in range(3)]
        for i in range(3):
            dense = grad[i, 0] * mats[0] + grad[i, 1] * mats[1] + grad[i, 2] * mats[2]
            star[e, i] = dense[star_pattern[:, 0], star_pattern[:, 1]]
        volume = abs(np.linalg.det(c[1:] - c[:1])) / 6.0
        for f in range(4):
            a, b, cc = c[FACE2NODES[f]]
            normal = np.cross(b - a, cc - a)
            surface = 0.5 * np.linalg.norm(normal)
            t1 = b - a
            t2 = np.cross(normal, t1)
            nn, s, t = normal / np.linalg.norm(normal), t1 / np.linalg.norm(t1), t2 / np.linalg.norm(t2)
            nb = neighbors[e, f]
            m = nb if nb >= 0 else e
            q_loc, q_nb = godunov_state_T(lam[e], mu[e], rho[e], lam[m], mu[m], rho[m])
            if face_types is not None and face_types[e, f] == 1:          # FaceType::FreeSurface
                q_loc, q_nb = free_surface_godunov_state_T(lam[e], mu[e], rho[e]), np.zeros((9, 9))
            T, Ti = rotation_matrices(nn, s, t)
            at = coefficient_matrix_T(lam[e], mu[e], rho[e], 0)
            scale = -2.0 * surface / (6.0 * volume)
            if face_types is not None and face_types[e, f] == 3:
                ap = np.zeros((9, 9))                                  # DynamicRupture: no local flux matrix
            else:
                ap = scale * (Ti.T @ (q_loc @ at) @ T.T)
            am = scale * (Ti.T @ (q_nb @ at) @ T.T)
            aplus[e, f] = ap.T.reshape(81)       # memory A[i + 9 j]
            aminus[e, f] = am.T.reshape(81)
    return star, aplus, aminus


def face_frame(c, f):
    """Unit normal and tangents, area of face f and volume of the tetrahedron c [4, 3] (MeshTools.cpp:44-132)."""
    a, b, cc = c[FACE2NODES[f]]
    normal = np.cross(b - a, cc - a)
    t1 = b - a
    t2 = np.cross(normal, t1)
    volume = abs(np.linalg.det(c[1:] - c[:1])) / 6.0
    return (normal / np.linalg.norm(normal), t1 / np.linalg.norm(t1), t2 / np.linalg.norm(t2),
            0.5 * np.linalg.norm(normal), volume)


def dr_matrices(coords_plus, coords_minus, face_info, mat_plus, mat_minus):
    """Fault-face matrices (host restatement of src/Initializer/Model/DynamicRuptureMatrices.cpp:111-477): impedances and
    eta (:341-359), TinvT (:427), fluxSolverPlus / Minus = fluxScale star0 . T^T (:461-469, dynamic_rupture.py:88),
    doubledSurfaceArea.  coords_* [m, 4, 3], face_info [m, 3] = {plusSide, minusSide, faceRelation}, mat_* [m, 3] =
    lambda, mu, rho.  Returns tinvT [m, 9, 9], imp_and_eta [m, 11], fs_plus, fs_minus [m, 9, 9] (row = quantity of the
    imposed state, column = updated quantity), doubled_area [m]."""
    m = coords_plus.shape[0]
    tinv_t = np.zeros((m, 9, 9))
    imp = np.zeros((m, 11))
    fsp = np.zeros((m, 9, 9))
    fsm = np.zeros((m, 9, 9))
    area2 = np.zeros(m)
    for i in range(m):
        lp, mp, rp = mat_plus[i]
        lm, mm, rm = mat_minus[i]
        zp, zs = rp * np.sqrt((lp + 2 * mp) / rp), rp * np.sqrt(mp / rp)
        zpn, zsn = rm * np.sqrt((lm + 2 * mm) / rm), rm * np.sqrt(mm / rm)
        imp[i] = [zp, zs, zpn, zsn, 1.0 / (1.0 / zp + 1.0 / zpn), 1.0 / (1.0 / zs + 1.0 / zsn), 1.0 / zs + 1.0 / zsn,
                  1.0 / zp, 1.0 / zs, 1.0 / zpn, 1.0 / zsn]
        n, s, t, surf_p, vol_p = face_frame(coords_plus[i], face_info[i, 0])
        _, _, _, surf_m, vol_m = face_frame(coords_minus[i], face_info[i, 1])
        T, Ti = rotation_matrices(n, s, t)
        tinv_t[i] = Ti.T
        area2[i] = 2.0 * surf_m
        fsp[i] = (-2.0 * surf_p / (6.0 * vol_p)) * (coefficient_matrix_T(lp, mp, rp, 0) @ T.T)
        fsm[i] = (2.0 * surf_m / (6.0 * vol_m)) * (coefficient_matrix_T(lm, mm, rm, 0) @ T.T)
    return tinv_t, imp, fsp, fsm, area2


def plane_wave(lam, mu, rho, direction, wavelength_count, length, kind="P"):
    """Exact plane-wave solution Q(x, t) of the homogeneous elastic equations in a periodic box of edge `length`:
    returns (func(x, t) -> [.., 9], wave speed).  direction: integer vector (periodicity), kind 'P' or 'S'."""
    d = np.asarray(direction, dtype=np.float64)
    kvec = 2.0 * np.pi * wavelength_count * d / length
    k = np.linalg.norm(kvec)
    n = kvec / k
    helper = np.array([0.0, 0.0, 1.0]) if abs(n[2]) < 0.9 else np.array([1.0, 0.0, 0.0])
    s = np.cross(n, helper)
    s /= np.linalg.norm(s)
    t = np.cross(n, s)
    T, _ = rotation_matrices(n, s, t)
    if kind == "P":
        c = np.sqrt((lam + 2 * mu) / rho)
        r = np.array([lam + 2 * mu, lam, lam, 0, 0, 0, -c, 0, 0])
    else:
        c = np.sqrt(mu / rho)
        r = np.array([0, 0, 0, mu, 0, 0, 0, -c, 0])
    amp = T @ r

    def func(x, time=0.0):
        phase = np.tensordot(x, kvec, axes=([-1], [0])) - k * c * time
        return np.sin(phase)[..., None] * amp
    return func, c
