"""Time clusters of a mesh from its cell time steps (host logic, numpy): the input of seissol_b200/lts.py.

What it restates from the reference:
* the CFL time step of a cell from its insphere radius and P-wave speed (src/Initializer/TimeStepping/
  GlobalTimestep.cpp:27-47: dt = cfl * 2 r_in / (v_p (2 O - 1)));
* rate-r clustering: cluster k holds the cells with r^k <= dt / (lambda dt_min) < r^(k+1), lambda = wiggle factor
  (docs/local-timestepping.rst "Wiggle factor", src/Initializer/Clustering/Clustering.cpp);
* the maximum-difference property: face neighbors differ by at most one cluster, enforced by lowering cells until a fixed
  point (src/Initializer/Clustering/ClusterSmoother.cpp) -- a hard requirement of the scheme (lts.py raises otherwise);
* the cost of a clustering per unit of simulated time (docs/local-timestepping.rst "The cost model",
  src/Initializer/Clustering/ClusterCostModel.h):
      C = sum_k (c_launch + c_update max(W_k, T_fill)) / (n_k lambda dt_min),   n_k = r^k,
  with the two GPU terms of docs/gpus.rst:60-68: c_launch (a cluster update costs its launches whatever it carries) and
  T_fill (below this many cells an update does not get cheaper);
* the grid search over wiggle factors and the merging of clusters from the top (LtsMaxNumberOfClusters).

GPU_COST gives the two terms MEASURED for this library on B200: one cluster update is 7 launches
(2 x ader, local, accumulate, 2 x Taylor, neigh) of ~4 us each = 28 us, a cell update takes 8 ns at 126 M updates/s, so
c_launch = 3500 cell updates; one wave of the persistent kernels is 296 CTAs x 28 cells = 8288 cells = T_fill.
"""
import numpy as np

GPU_COST = dict(c_update=1.0, c_launch=3500.0, t_fill=8288.0)
CPU_COST = dict(c_update=1.0, c_launch=0.0, t_fill=0.0)


def insphere_radius(coords):
    """coords [n, 4, 3] -> insphere radius (GlobalTimestep.cpp:31-42): |det [x 1]| / sum of the doubled face areas."""
    x = np.asarray(coords, dtype=np.float64)
    a = np.concatenate([x, np.ones(x.shape[:2] + (1,))], axis=2)
    alpha = np.linalg.det(a)

    def area2(i, j, k):
        return np.linalg.norm(np.cross(x[:, j] - x[:, i], x[:, k] - x[:, i]), axis=1)
    return np.abs(alpha) / (area2(0, 1, 2) + area2(0, 1, 3) + area2(0, 2, 3) + area2(1, 2, 3))


def cell_timestep(coords, p_wave_speed, order, cfl=0.5, max_timestep=np.inf):
    """computeCellTimestep (GlobalTimestep.cpp:27-47)."""
    return np.minimum(max_timestep, cfl * 2.0 * insphere_radius(coords) / (np.asarray(p_wave_speed) * (2 * order - 1)))


def cluster_ids(dt, rate=2, wiggle=1.0, max_clusters=None):
    """Cluster of every cell: floor(log_rate(dt / (wiggle dt_min))), merged from the top into max_clusters clusters.
    Returns (cluster ids, base time step wiggle * dt_min)."""
    dt = np.asarray(dt, dtype=np.float64)
    base = wiggle * dt.min()
    k = np.floor(np.log(dt / base) / np.log(rate) + 1e-12).astype(np.int64)
    k = np.maximum(k, 0)
    if max_clusters is not None:
        k = np.minimum(k, max_clusters - 1)
    return k, base


def enforce_maximum_difference(cluster, neighbors, regular=None, max_iterations=10000):
    """Lower cells until face neighbors differ by at most one cluster (ClusterSmoother.cpp): cluster[e] <= cluster[nb] + 1.
    neighbors [n, 4] (index; ignored where `regular` is False or the index is negative).  Returns (ids, #sweeps)."""
    c = np.asarray(cluster, dtype=np.int64).copy()
    nb = np.asarray(neighbors)
    ok = nb >= 0 if regular is None else (np.asarray(regular) & (nb >= 0))
    nbc = np.where(ok, nb, 0)
    big = np.iinfo(np.int64).max // 2
    for sweep in range(max_iterations):
        lim = np.where(ok, c[nbc] + 1, big).min(axis=1)
        new = np.minimum(c, lim)
        if np.array_equal(new, c):
            return c, sweep
        c = new
    raise RuntimeError("maximum-difference smoothing did not converge")


def clustering_cost(cluster, base_dt, rate=2, weights=None, c_update=1.0, c_launch=0.0, t_fill=0.0):
    """Cost per unit of simulated time (docs/local-timestepping.rst "The cost model")."""
    cluster = np.asarray(cluster)
    w = np.ones(len(cluster)) if weights is None else np.asarray(weights, dtype=np.float64)
    cost = 0.0
    for k in range(int(cluster.max()) + 1):
        wk = w[cluster == k].sum()
        if wk > 0:
            cost += (c_launch + c_update * max(wk, t_fill)) / (float(rate) ** k * base_dt)
    return cost


def best_clustering(dt, neighbors, regular=None, rate=2, wiggle_min=1.0, wiggle_step=0.01, max_clusters=None,
                    enforce=True, **cost_terms):
    """Grid search over the wiggle factor in [wiggle_min, 1] (docs/local-timestepping.rst "Wiggle factor") and, with
    GPU cost terms, over the number of clusters (merging from the top): the clustering of least cost.
    Returns dict(cluster, base_dt, wiggle, n_clusters, cost)."""
    best = None
    lam = 1.0
    while lam >= wiggle_min - 1e-12:
        top, _ = cluster_ids(dt, rate, lam, max_clusters)
        limits = range(1, int(top.max()) + 2) if (cost_terms.get("c_launch", 0.0) > 0 or cost_terms.get("t_fill", 0.0) > 0) \
            else [int(top.max()) + 1]
        for ncl in limits:
            ids, base = cluster_ids(dt, rate, lam, ncl if max_clusters is None else min(ncl, max_clusters))
            if enforce:
                ids, _ = enforce_maximum_difference(ids, neighbors, regular)
            cost = clustering_cost(ids, base, rate, **cost_terms)
            if best is None or cost < best["cost"] * (1.0 - 1e-12):
                best = dict(cluster=ids, base_dt=base, wiggle=lam, n_clusters=int(ids.max()) + 1, cost=cost)
        lam -= wiggle_step
    return best


def gts_cost(dt, **cost_terms):
    """Cost of global time stepping (everything in one cluster at dt_min) for comparison."""
    dt = np.asarray(dt)
    return clustering_cost(np.zeros(len(dt), dtype=np.int64), dt.min(), **cost_terms)
