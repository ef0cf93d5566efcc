"""Modal basis of the reference tetrahedron (host utility, numpy): Dubiner polynomials in the reference's ordering, a
quadrature rule and the L2 projection used to set initial conditions / measure errors.

Restated from src/Numerical/Functions.cpp:136-148 (TetraDubinerP with singularity-free Jacobi polynomials),
src/Numerical/BasisFunction.h:135-143 (ordering: for ord, for k <= ord, for j <= ord - k: (ord - j - k, j, k)) and the
projection of src/Initializer/InitialFieldProjection.cpp (quadrature + diagonal mass matrix)."""
import numpy as np
from scipy.special import eval_jacobi


def basis_indices(order):
    return [(o - j - k, j, k) for o in range(order) for k in range(o + 1) for j in range(o - k + 1)]


def dubiner(order, xi, eta, zeta):
    """[len(xi), B] values of the B = O(O+1)(O+2)/6 basis functions at points of the reference tetrahedron."""
    xi, eta, zeta = (np.asarray(a, dtype=np.float64) for a in (xi, eta, zeta))
    r = 2.0 * xi - 1.0 + eta + zeta
    s = 2.0 * eta - 1.0 + zeta
    t = 2.0 * zeta - 1.0
    st = 1.0 - eta - zeta
    th = 1.0 - zeta

    def sfj(n, a, x, y):   # y^n P_n^{(a,0)}(x / y)
        yy = np.where(np.abs(y) < 1e-300, 1e-300, y)
        return yy ** n * eval_jacobi(n, a, 0, x / yy)

    cols = [sfj(i, 0, r, st) * sfj(j, 2 * i + 1, s, th) * eval_jacobi(k, 2 * i + 2 * j + 2, 0, t)
            for (i, j, k) in basis_indices(order)]
    return np.stack(cols, -1)


def tet_quadrature(n):
    """Gauss-Legendre^3 collapsed onto the reference tetrahedron: exact for polynomials of degree 2n - 3.
    Returns points [m, 3] and weights [m] (sum = 1/6)."""
    x, w = np.polynomial.legendre.leggauss(n)
    x, w = 0.5 * (x + 1.0), 0.5 * w
    a, b, c = np.meshgrid(x, x, x, indexing="ij")
    wa, wb, wc = np.meshgrid(w, w, w, indexing="ij")
    xi = a
    eta = b * (1.0 - a)
    zeta = c * (1.0 - a) * (1.0 - b)
    wt = wa * wb * wc * (1.0 - a) ** 2 * (1.0 - b)
    return np.stack([xi.ravel(), eta.ravel(), zeta.ravel()], -1), wt.ravel()


def physical_points(coords, ref_pts):
    """[n, 4, 3] vertex coordinates, [m, 3] reference points -> [n, m, 3]."""
    v0 = coords[:, 0][:, None, :]
    e = coords[:, 1:] - coords[:, :1]                       # [n, 3 (xi, eta, zeta), 3]
    return v0 + np.einsum("md,ndc->nmc", ref_pts, e)


def project(order, coords, func, nq=None):
    """L2 projection of func(x [.., 3]) -> [.., 9] onto the modal basis of every element: returns Q [n, B, 9]."""
    pts, w = tet_quadrature(nq or order + 2)
    phi = dubiner(order, pts[:, 0], pts[:, 1], pts[:, 2])    # [m, B]
    mass = np.einsum("m,mk,mk->k", w, phi, phi)              # diagonal (orthogonal basis)
    f = func(physical_points(coords, pts))                   # [n, m, 9]
    return np.einsum("m,mk,nmq->nkq", w, phi, f) / mass[None, :, None]


def l2_error(order, coords, Q, func, nq=None):
    """sqrt(sum_elements |J| int (Q_h - f)^2) per quantity."""
    pts, w = tet_quadrature(nq or order + 3)
    phi = dubiner(order, pts[:, 0], pts[:, 1], pts[:, 2])
    qh = np.einsum("mk,nkq->nmq", phi, Q)
    f = func(physical_points(coords, pts))
    jac = np.abs(np.linalg.det(coords[:, 1:] - coords[:, :1]))   # = 6 |V|
    return np.sqrt(np.einsum("n,m,nmq->q", jac, w, (qh - f) ** 2))
