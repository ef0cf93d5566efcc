"""Partition of the cube mesh across the GPUs of one box + the LTS layer layout of one rank (host logic, numpy).

The reference partitions a mesh into MPI ranks and lays every rank's cells out as
(time cluster) x (Ghost | Copy | Interior) layers (src/Memory/Tree/Colormap.h:210-227, src/Initializer/
InitProcedure/Internal/MeshLayout.cpp): copy cells are the own cells with a face neighbor on
another rank, packed contiguously per (peer, cluster) region so that one message moves a whole region; ghost
cells mirror the peers' copy regions.  Here:

* slabs of cubes along x, periodic (rank r owns x in [r*nxl, (r+1)*nxl)), so every rank has exactly two peers
  (ring) and the same load -- the weak-scaling set-up of BASELINE.json configs[4];
* time clusters: horizontal bands in z (cluster c has time step 2^c dt; rate-2 LTS), neighbouring cells differ
  by at most one cluster as the reference enforces (src/Initializer/TimeStepping/LtsLayout / Clustering);
* own cells are ordered [copy regions ... | interior], regions ordered by (cluster, side), cells by global id --
  the peer orders its ghost region the same way, so a region is a contiguous slice on both ends and the halo
  exchange needs no pack/unpack kernels (the slice of `buffers` is sent from where it lies and received where
  the neighbor kernel reads it).
"""
from dataclasses import dataclass, field

import numpy as np

from .mesh import CubeMesh

OUTFLOW = 5  # FaceType::Outflow (src/Initializer/BasicTypedefs.h:23-48)


@dataclass
class Region:
    """A contiguous run of cells exchanged with one peer: [start, start+count) in the local buffer index space
    (own cells first, ghost cells after them)."""
    peer: int
    side: int        # 0: towards x-, 1: towards x+
    cluster: int
    start: int
    count: int


@dataclass
class RankLayout:
    rank: int
    world: int
    n_own: int
    n_ghost: int
    n_copy: int                      # own cells [0, n_copy) are copy cells, [n_copy, n_own) interior
    gid: np.ndarray                  # [n_own + n_ghost] global element ids
    cluster: np.ndarray              # [n_own + n_ghost] time cluster
    neighbors: np.ndarray            # [n_own, 4] local buffer index (own or ghost) or own index for boundary faces
    face_rel: np.ndarray             # [n_own, 4, 2]
    face_types: np.ndarray           # [n_own, 4]
    send: list = field(default_factory=list)   # copy regions  (Region.start in [0, n_copy))
    recv: list = field(default_factory=list)   # ghost regions (Region.start in [n_own, n_own + n_ghost))
    cluster_ranges: dict = field(default_factory=dict)  # cluster -> {"copy": (a, b), "interior": (a, b)}


def cluster_of_z(cube_z, nz, n_clusters):
    """Horizontal bands: the top band (largest z) is cluster 0 (smallest time step), each band below doubles the
    time step and the band height, mimicking a mesh graded towards the surface."""
    if n_clusters <= 1:
        return np.zeros_like(cube_z)
    depth = (nz - 1 - cube_z).astype(np.float64) / max(nz, 1)      # 0 at the top
    edges = np.array([(2.0 ** c - 1.0) / (2.0 ** n_clusters - 1.0) for c in range(1, n_clusters)])
    return np.searchsorted(edges, depth, side="right").astype(np.int64)


def cluster_of_x(cube_x, nx, n_clusters):
    """Clusters varying ALONG the partition direction (periodic in x): two-cube layers that walk 0, 1, .., C-1, .., 1, 0, ..
    so that EVERY even x is a cluster boundary.  Slab boundaries then separate cells of different time clusters: the halo
    has to carry derivative families and the receiver has to accumulate ghost buffers (the general exchange of
    seissol_b200/lts.py; nx must be a multiple of 4 (C - 1))."""
    if n_clusters <= 1:
        return np.zeros_like(cube_x)
    period = 2 * (n_clusters - 1)
    if nx % (2 * period):
        raise ValueError("nx must be a multiple of 4 (clusters - 1) for clusters along x")
    tri = np.array(list(range(n_clusters)) + list(range(n_clusters - 2, 0, -1)), dtype=np.int64)
    return tri[(cube_x // 2) % period]


def build_layout(nx, ny, nz, rank, world, n_clusters=1, cluster_axis="z", cluster_by_gid=None):
    """Layout of `rank`: which cells it owns, in which order, where its ghost cells sit and what is exchanged.
    cluster_axis: "z" = horizontal bands (remote neighbors share the cluster), "x" = bands across the slabs;
    cluster_by_gid: time cluster of every GLOBAL element id (e.g. seissol_b200/clustering.py::best_clustering on the
    whole mesh: clusters from the cells' CFL time steps), overrides the synthetic bands."""
    if nx % world:
        raise ValueError("nx must be a multiple of the number of ranks")
    nxl = nx // world
    if world > 1 and nxl < 2:
        raise ValueError("slabs must be at least two cubes thick")
    x_lo, x_hi = rank * nxl, (rank + 1) * nxl
    if world == 1:
        mesh = CubeMesh(nx, ny, nz)
        own = np.ones(mesh.n, dtype=bool)
    else:
        lo, hi = x_lo - 1, x_hi + 1
        if hi - lo > nx:        # two ranks, two-cube slabs: the halo cubes of both sides coincide
            lo, hi = 0, nx
        mesh = CubeMesh(nx, ny, nz, lo, hi)
        own = (mesh.cube_x >= x_lo) & (mesh.cube_x < x_hi)
    if cluster_by_gid is not None:
        cluster = np.asarray(cluster_by_gid, dtype=np.int64)[mesh.gid]
    else:
        cluster = cluster_of_z(mesh.cube_z, nz, n_clusters) if cluster_axis == "z" else cluster_of_x(mesh.cube_x, nx, n_clusters)
    nb = mesh.neighbors
    has = nb >= 0
    nb_own = np.where(has, own[np.maximum(nb, 0)], True)
    remote_face = own[:, None] & ~nb_own                               # own cell, neighbor on another rank
    # side of a remote neighbor: its cube is the layer just below (0) or just above (1) the slab
    nb_x = mesh.cube_x[np.maximum(nb, 0)]
    side_of_face = np.where(nb_x == (x_lo - 1) % nx, 0, 1)
    if world == 2 and nxl * 2 == nx:
        # ring of two: both sides lead to the same peer; distinguish by which own layer touches
        side_of_face = np.where(mesh.cube_x[:, None] == x_lo, 0, 1) * np.ones((1, 4), dtype=np.int64)
    is_copy = remote_face.any(axis=1)
    copy_side = np.where(is_copy, np.take_along_axis(side_of_face, remote_face.argmax(axis=1)[:, None], 1)[:, 0], -1)
    if world > 1:
        both = (remote_face & (side_of_face == 0)).any(1) & (remote_face & (side_of_face == 1)).any(1)
        assert not both.any(), "a copy cell may touch only one peer region"
    # ghost cells: not owned, neighbor of an own cell; their side is the side of the slab they lie on
    ghost = np.zeros(mesh.n, dtype=bool)
    ghost[nb[remote_face]] = True
    ghost_side = np.full(mesh.n, -1)
    e, f = np.nonzero(remote_face)
    ghost_side[nb[e, f]] = side_of_face[e, f]

    def ordered(mask, side):
        idx = np.nonzero(mask)[0]
        key = np.lexsort((mesh.gid[idx], side[idx], cluster[idx]))     # (cluster, side, gid)
        return idx[key]

    copy_cells = ordered(own & is_copy, copy_side)
    interior_cells = np.nonzero(own & ~is_copy)[0]
    interior_cells = interior_cells[np.lexsort((mesh.gid[interior_cells], cluster[interior_cells]))]
    ghost_cells = ordered(ghost, ghost_side)
    order = np.concatenate([copy_cells, interior_cells, ghost_cells])
    n_copy, n_own, n_ghost = len(copy_cells), len(copy_cells) + len(interior_cells), len(ghost_cells)
    new_index = np.full(mesh.n, -1, dtype=np.int64)
    new_index[order] = np.arange(len(order))

    own_order = order[:n_own]
    neighbors = new_index[np.maximum(nb[own_order], 0)]
    face_types = np.where(has[own_order], 0, OUTFLOW).astype(np.uint8)
    # boundary faces carry no neighbor: point them at the cell itself (never dereferenced for Outflow)
    neighbors = np.where(has[own_order], neighbors, np.arange(n_own)[:, None])
    assert (neighbors >= 0).all()

    lay = RankLayout(rank=rank, world=world, n_own=n_own, n_ghost=n_ghost, n_copy=n_copy,
                     gid=mesh.gid[order], cluster=cluster[order], neighbors=neighbors,
                     face_rel=mesh.face_rel[own_order], face_types=face_types)
    peers = ((rank - 1) % world, (rank + 1) % world)

    def regions(cells, sides, base):
        out = []
        if len(cells) == 0:
            return out
        keys = cluster[cells] * 2 + sides[cells]
        bounds = np.nonzero(np.diff(keys))[0] + 1
        starts = np.concatenate([[0], bounds])
        ends = np.concatenate([bounds, [len(cells)]])
        for a, b in zip(starts, ends):
            s = int(sides[cells[a]])
            out.append(Region(peer=peers[s], side=s, cluster=int(cluster[cells[a]]), start=base + int(a), count=int(b - a)))
        return out

    lay.send = regions(copy_cells, copy_side, 0)
    lay.recv = regions(ghost_cells, ghost_side, n_own)
    for c in range(max(n_clusters, 1)):
        cc = np.nonzero(cluster[copy_cells] == c)[0]
        ic = np.nonzero(cluster[interior_cells] == c)[0]
        lay.cluster_ranges[c] = {
            "copy": (int(cc[0]), int(cc[-1]) + 1) if len(cc) else (0, 0),
            "interior": (n_copy + int(ic[0]), n_copy + int(ic[-1]) + 1) if len(ic) else (n_copy, n_copy),
        }
    return lay


def export_layout(layout, path):
    """RankLayout -> the binary file host/ssb_lts.hpp::Layout::read loads (little endian; int64 header
    {magic "SSBLTS01", rank, world, n_own, n_ghost, n_clusters, #send, #recv}, then gid, cluster, neighbors (int64),
    face_rel, face_types (uint8), send and recv regions as {peer, side, cluster, start, count} int64 records)."""
    n_clusters = int(layout.cluster.max(initial=0)) + 1 if len(layout.cluster) else 1
    n_clusters = max(n_clusters, len(layout.cluster_ranges))
    with open(path, "wb") as f:
        np.array([0x5353424C54533031, layout.rank, layout.world, layout.n_own, layout.n_ghost, n_clusters,
                  len(layout.send), len(layout.recv)], dtype=np.int64).tofile(f)
        layout.gid.astype(np.int64).tofile(f)
        layout.cluster.astype(np.int64).tofile(f)
        np.ascontiguousarray(layout.neighbors, dtype=np.int64).tofile(f)
        np.ascontiguousarray(layout.face_rel, dtype=np.uint8).tofile(f)
        np.ascontiguousarray(layout.face_types, dtype=np.uint8).tofile(f)
        for regs in (layout.send, layout.recv):
            np.array([[r.peer, r.side, r.cluster, r.start, r.count] for r in regs], dtype=np.int64).reshape(-1, 5).tofile(f)
