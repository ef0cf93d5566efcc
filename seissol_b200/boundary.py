"""Host side of the FaceType::Analytical boundary condition.

The reference evaluates the analytical solution on the HOST (a user callback per node and time) even in its GPU build
(Local::evaluateBatchedTimeDependentBc, src/Kernels/LinearCK/Local.cpp:397-455): per face it integrates the callback over
the time step with DirichletBoundary::evaluateTimeDependent (src/Kernels/LinearCK/DirichletBoundary.h:121-165), copies the
nodal values to LTS::AnalyticScratch and launches localFluxNodal on the device.  `analytical_inodal` is that host loop;
`Context.local_nodal_flux` (ssb_local_nodal_flux) is the device launch."""
import numpy as np


def analytical_inodal(evaluate, nodes, start_time, time_step_width, order, dtype=np.float64):
    """Time-integrated boundary values at the face nodes.

    evaluate(nodes_i [B2, 3], time) -> [B2, 9]: the analytical solution at the nodes of ONE face in global coordinates
    (physics::InitialField::evaluate through ApplyAnalyticalSolution, Local.cpp:86-118); nodes [m, B2, 3]
    (CellBoundaryMapping::nodes).  Quadrature: `order` Gauss-Legendre points mapped to [t, t + dt]
    (DirichletBoundary.h:139-146): t_i = (dt x_i + 2 t + dt) / 2, w_i = dt w_i / 2.
    Returns [m, 9, B2] (quantity, node): the layout ssb_local_nodal_flux reads."""
    x, w = np.polynomial.legendre.leggauss(order)
    points = (time_step_width * x + 2.0 * start_time + time_step_width) / 2.0
    weights = 0.5 * time_step_width * w
    nodes = np.asarray(nodes, dtype=np.float64)
    out = np.zeros((nodes.shape[0], 9, nodes.shape[1]), dtype=np.float64)
    for i in range(nodes.shape[0]):
        for t, wt in zip(points, weights):
            out[i] += wt * np.asarray(evaluate(nodes[i], t), dtype=np.float64).T
    return out.astype(dtype)
