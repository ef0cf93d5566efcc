"""Clustered local time stepping (rate 2) of the element update, one rank or an x-slab partition: host logic above the C ABI.

What it replaces in the reference: the LTS part of the time-cluster step (src/Solver/TimeStepping/TimeCluster.cpp:
computeLocalIntegration :358-538 with buffer reset/accumulation, computeNeighboringIntegration :556-998 with
TimeCommon::computeIntegrals, src/Kernels/TimeCommon.cpp:40-125) and the per-cell LtsSetup bits that drive it
(src/Initializer/LtsSetup.h:21-111, InitProcedure/Internal/LtsSetup.cpp):

  cluster c advances with dt_c = 2^c dt_0.  For a cell e in cluster c and a regular face with neighbor n:
    cl(n) == c     : the neighbor's time-integrated DOFs of the same interval.  Read from its buffer, unless the
                     neighbor ACCUMULATES its buffer for a larger cluster (then its buffer holds a sum over two
                     steps): integrate its derivatives over the full interval instead (neighborHasDerivatives + GTS).
    cl(n) == c + 1 : the neighbor's predictor covers two of my steps: integrate its derivatives over my sub-interval
                     [off, off + dt_c] relative to its expansion point (neighborHasDerivatives, no GTS relation).
    cl(n) == c - 1 : the neighbor makes two steps per step of mine and accumulates both integrals in its buffer
                     (accumulateBuffers; reset at the start of my step).
  Cells that somebody integrates in time keep their derivative family (hasDerivatives).

Own cells are re-ordered [cluster][accumulating | other cells with derivatives | rest], so that every kernel launch
of a cluster works on contiguous slices; only the Taylor-sum inputs and the neighbor pointers (LTS::FaceNeighbors,
src/Memory/Descriptor/LTS.h:107) are pointer tables.  The compute back-end is injected (`CudaLtsBackend` = the C ABI;
the CPU tests inject a numpy back-end from tests/), so this module holds no CPU compute path.
"""
import math
from dataclasses import dataclass, field

import numpy as np


def integrate_coeffs(order, t0, t1):
    """MonomialBasis::integrate (src/Numerical/TimeBasis.h:82-96) for [t0, t1] relative to the expansion point."""
    return np.array([(t1 ** (d + 1) - t0 ** (d + 1)) / math.factorial(d + 1) for d in range(order)])


@dataclass
class ClusterRange:
    start: int      # first cell of the cluster
    acc_end: int    # [start, acc_end): cells that accumulate their buffers
    der_end: int    # [start, der_end): cells that keep their derivative family
    end: int        # [start, end): all cells of the cluster


@dataclass
class LtsPlan:
    """Re-ordered single-rank mesh + everything the LTS step needs, as numpy arrays (host logic only)."""
    n: int
    n_clusters: int
    gid: np.ndarray          # [n] global element id of the re-ordered cells
    cluster: np.ndarray      # [n]
    neighbors: np.ndarray    # [n, 4] re-ordered cell index (own index on boundary faces)
    face_rel: np.ndarray     # [n, 4, 2]
    face_types: np.ndarray   # [n, 4]
    ranges: list             # ClusterRange per cluster
    der_row: np.ndarray      # [n] row in the derivative storage or -1
    acc_row: np.ndarray      # [n] row in the accumulation-buffer storage or -1
    kind: np.ndarray         # [n, 4] 0 = read a buffer, 1 = Taylor sum over the full interval, 2 = over a sub-interval
    scratch_row: np.ndarray  # [n, 4] row in the Taylor scratch storage or -1
    taylor_full: dict = field(default_factory=dict)   # cluster -> (cell, face) index arrays of kind 1
    taylor_sub: dict = field(default_factory=dict)    # cluster -> (cell, face) index arrays of kind 2
    n_der: int = 0
    n_acc: int = 0
    n_scratch: int = 0
    # partitioned mesh: ghost cells sit behind the own cells (index n + g); per cluster the copy cells to send
    n_ghost: int = 0
    world: int = 1
    rank: int = 0
    send: dict = field(default_factory=dict)   # cluster -> list of (peer, side, own cell indices in the peer's ghost order)
    recv: dict = field(default_factory=dict)   # cluster -> list of (peer, side, first ghost row, count)
    # remote neighbors in ANOTHER time cluster (the copy regions of the reference carry "buffers and / or derivatives",
    # InitProcedure/Internal/LtsSetup.cpp:115-189, DirectGhostTimeCluster.cpp:22-105):
    #  * a ghost cell with an own neighbor in the next SMALLER cluster sends its derivative family after every
    #    prediction; it lands in derivative rows behind the own ones (der_row is defined for ghost indices too);
    #  * a ghost cell with an own neighbor in the next LARGER cluster is read through the sum of its two steps: the
    #    RECEIVER accumulates the per-step rows it gets (acc_row is defined for ghost indices too).
    cluster_ext: np.ndarray = None             # [n + n_ghost]
    send_der: dict = field(default_factory=dict)    # cluster -> list of (peer, side, derivative rows to send)
    recv_der: dict = field(default_factory=dict)    # cluster -> list of (peer, side, first derivative row, count)
    ghost_acc: dict = field(default_factory=dict)   # cluster -> (ghost cell indices, accumulation rows)
    n_der_own: int = 0
    n_acc_own: int = 0

    def updates_per_macro_step(self):
        """Element updates in one step of the largest cluster."""
        return int(sum((r.end - r.start) * 2 ** (self.n_clusters - 1 - c) for c, r in enumerate(self.ranges)))


def build_plan(layout, n_clusters):
    """LtsPlan from a RankLayout (seissol_b200/partition.py::build_layout).  With more than one rank the ghost cells keep
    their layout order behind the re-ordered own cells.  The halo carries one time-integrated tensor per copy cell and
    step of its cluster; where a slab boundary separates cells of different clusters it also carries the derivative
    families of the larger cluster's copy cells, and the smaller cluster's ghost rows are accumulated on arrival."""
    n, ng = layout.n_own, layout.n_ghost
    cl_all = layout.cluster.astype(np.int64)
    cl = cl_all[:n]
    nb = layout.neighbors                      # index into [own | ghost] of the layout order
    regular = layout.face_types == 0
    remote = regular & (nb >= n)
    nb_cl = cl_all[nb]
    diff = np.where(regular, nb_cl - cl[:, None], 0)
    if np.abs(diff).max(initial=0) > 1:
        raise ValueError("neighboring cells may differ by at most one time cluster")
    larger, smaller = diff == 1, diff == -1
    accumulates = larger.any(1)
    has_derivs = accumulates | smaller.any(1)
    group = np.where(accumulates, 0, np.where(has_derivs, 1, 2))
    order = np.lexsort((layout.gid[:n], group, cl))
    new_index = np.empty(n + ng, dtype=np.int64)
    new_index[order] = np.arange(n)
    new_index[n:] = np.arange(n, n + ng)
    cl_o, group_o = cl[order], group[order]
    neighbors = new_index[nb[order]]
    face_types = layout.face_types[order]
    regular = face_types == 0
    local_nb = neighbors < n
    ranges = []
    for c in range(n_clusters):
        idx = np.nonzero(cl_o == c)[0]
        a = int(idx[0]) if len(idx) else (ranges[-1].end if ranges else 0)
        b = int(idx[-1]) + 1 if len(idx) else a
        g = group_o[a:b]
        ranges.append(ClusterRange(a, a + int((g == 0).sum()), a + int((g <= 1).sum()), b))
    acc_o, der_o = group_o == 0, group_o <= 1
    n_der_own, n_acc_own = int(der_o.sum()), int(acc_o.sum())
    cl_ext = np.concatenate([cl_o, cl_all[n:]])
    nb_cl = cl_ext[neighbors]
    # ghost cells seen from a smaller / larger own cluster (layout order of the ghost cells = order of the messages)
    ghost_der = np.zeros(ng, dtype=bool)
    ghost_accm = np.zeros(ng, dtype=bool)
    rem = regular & ~local_nb
    e, f = np.nonzero(rem & (nb_cl == cl_o[:, None] + 1))
    ghost_der[neighbors[e, f] - n] = True
    e, f = np.nonzero(rem & (nb_cl == cl_o[:, None] - 1))
    ghost_accm[neighbors[e, f] - n] = True
    der_row = np.concatenate([np.where(der_o, np.cumsum(der_o) - 1, -1),
                              np.where(ghost_der, n_der_own + np.cumsum(ghost_der) - 1, -1)])
    acc_row = np.concatenate([np.where(acc_o, np.cumsum(acc_o) - 1, -1),
                              np.where(ghost_accm, n_acc_own + np.cumsum(ghost_accm) - 1, -1)])
    acc_ext = np.concatenate([acc_o, np.zeros(ng, dtype=bool)])     # same-cluster ghost cells: their per-step integral
    kind = np.zeros((n, 4), dtype=np.int64)
    kind[regular & local_nb & (nb_cl == cl_o[:, None]) & acc_ext[neighbors]] = 1
    kind[regular & (nb_cl == cl_o[:, None] + 1)] = 2
    need = kind > 0
    scratch_row = np.where(need, np.cumsum(need.ravel()).reshape(n, 4) - 1, -1)
    assert np.all(der_row[neighbors[need]] >= 0), "a neighbor that is integrated in time must keep its derivatives"
    sm = regular & (nb_cl == cl_o[:, None] - 1)
    assert np.all(acc_row[neighbors[sm]] >= 0)
    plan = LtsPlan(n=n, n_clusters=n_clusters, gid=layout.gid[:n][order], cluster=cl_o, neighbors=neighbors,
                   face_rel=layout.face_rel[order], face_types=face_types, ranges=ranges, der_row=der_row, acc_row=acc_row,
                   kind=kind, scratch_row=scratch_row, n_der=n_der_own + int(ghost_der.sum()),
                   n_acc=n_acc_own + int(ghost_accm.sum()), n_scratch=int(need.sum()), n_ghost=ng, world=layout.world,
                   rank=layout.rank, cluster_ext=cl_ext, n_der_own=n_der_own, n_acc_own=n_acc_own)
    # own copy cells whose derivative family a peer needs: those with a remote neighbor in the next smaller cluster
    # (an x-slab copy cell touches ONE peer region, so "any remote neighbor" and "a neighbor on that peer" coincide)
    sends_der = (rem & (nb_cl == cl_o[:, None] - 1)).any(1)
    for c, r in enumerate(ranges):
        for k, store in ((1, plan.taylor_full), (2, plan.taylor_sub)):
            e, f = np.nonzero(kind[r.start:r.end] == k)
            store[c] = (e + r.start, f)
        plan.send[c] = [(reg.peer, reg.side, new_index[reg.start:reg.start + reg.count]) for reg in layout.send
                        if reg.cluster == c]
        plan.recv[c] = [(reg.peer, reg.side, reg.start, reg.count) for reg in layout.recv if reg.cluster == c]
        plan.send_der[c], plan.recv_der[c] = [], []
        for peer, side, idx in plan.send[c]:
            sub = idx[sends_der[idx]]
            if len(sub):
                assert np.all(der_row[sub] >= 0)
                plan.send_der[c].append((peer, side, der_row[sub]))
        for peer, side, start, count in plan.recv[c]:
            rows = der_row[start:start + count]
            rows = rows[rows >= 0]
            if len(rows):
                assert np.array_equal(rows, np.arange(rows[0], rows[0] + len(rows)))
                plan.recv_der[c].append((peer, side, int(rows[0]), len(rows)))
        g = np.nonzero(ghost_accm & (cl_all[n:] == c))[0] + n
        plan.ghost_acc[c] = (g, acc_row[g])
    return plan


class LtsSolver:
    """Rate-2 clustered LTS schedule.  One call of `macro_step` advances every cluster to t + dt_{C-1}:

        for every fine sub-step s:  predict the clusters whose step starts at s (ader + volume/local flux, derivatives,
                                    buffer reset/accumulation), then correct the clusters whose step ends at s + 1
                                    (Taylor sums of the neighbors' derivatives + neighbor flux)

    which is the order the reference's actor scheduler produces for rate 2 (AbstractTimeCluster.cpp:35-199: a cluster
    may correct when its neighbors' predictions cover its interval, and may predict again when its neighbors have
    consumed the previous prediction)."""

    def __init__(self, plan, backend, order, dt0=1.0):
        self.plan, self.be, self.order, self.dt0 = plan, backend, order, dt0
        self.fine_steps = 2 ** (plan.n_clusters - 1)

    def macro_step(self):
        plan, be = self.plan, self.be
        for s in range(self.fine_steps):
            for c, r in enumerate(plan.ranges):
                if s % (2 ** c) == 0:
                    if r.end > r.start:
                        dt = self.dt0 * 2 ** c
                        be.predict(c, r, integrate_coeffs(self.order, 0.0, dt), reset=(s % (2 ** (c + 1)) == 0))
                    if plan.world > 1 and (plan.send[c] or plan.recv[c]):
                        # copy cells of cluster c -> the peers' ghost rows (one message per region; derivative
                        # families where a peer integrates them); ghost rows read by a larger cluster are accumulated.
                        # A rank without own cells of cluster c may still hold ghost cells of it.
                        be.exchange(c, reset=(s % (2 ** (c + 1)) == 0))
            for c, r in enumerate(plan.ranges):
                if (s + 1) % (2 ** c) == 0 and r.end > r.start:
                    dt = self.dt0 * 2 ** c
                    start = (s + 1) - 2 ** c                      # my step starts at fine sub-step `start`
                    off = (start % (2 ** (c + 1))) * self.dt0     # offset to the larger cluster's expansion point
                    be.correct(c, r, integrate_coeffs(self.order, 0.0, dt), integrate_coeffs(self.order, off, off + dt))


class CudaLtsBackend:
    """predict / correct of one cluster through the C ABI (seissol_b200/api.py) on the re-ordered device arrays.

    faceproj (default at order >= 6; SSB_FACEPROJ=0 switches it off): the neighbor flux runs on face projections
    (include/seissol_b200.h): the local kernel stores F_f = fMrT_f I for the four faces of every own cell, cells that
    accumulate for a larger cluster accumulate these blocks, neighbors that are read through their derivatives get
    the Taylor sum AND the projection of the one face the reader sees in one kernel (ssb_taylor_faceproj), and the halo
    carries the copy cells' projection blocks.  FaceNeighbors then points at projection blocks instead of time-integrated
    DOFs; the schedule is unchanged."""

    def __init__(self, torch, ctx, plan, dofs, local_int, neigh_int, dist=None, faceproj=None):
        import os
        from . import api
        t = torch
        self.torch, self.api, self.ctx, self.plan = t, api, ctx, plan
        dev, dt = dofs.device, dofs.dtype
        if faceproj is None:
            faceproj = os.environ.get("SSB_FACEPROJ", "1" if ctx.order >= 6 else "0") != "0"
        self.faceproj = bool(faceproj)
        self.dofs, self.local_int, self.neigh_int = dofs, local_int, neigh_int
        n, w = plan.n, dofs.shape[1]
        es = dofs.element_size()
        self.derivs = t.zeros((max(plan.n_der, 1), ctx.dersize), dtype=dt, device=dev)
        self.face_types = t.from_numpy(np.ascontiguousarray(plan.face_types)).to(dev)
        self.face_rel = t.from_numpy(np.ascontiguousarray(plan.face_rel)).to(dev)
        nb = plan.neighbors
        cl_ext = plan.cluster_ext if plan.cluster_ext is not None else plan.cluster
        # neighbors (own or ghost cells) in the next smaller cluster: read through their accumulation row
        smaller = (cl_ext[nb] == plan.cluster[:, None] - 1) & (plan.face_types == 0)
        if self.faceproj:
            fp = ctx.face_proj_reals()
            prj = fp // 4
            self.fp, self.prj = fp, prj
            self.idofs = t.zeros((n, w), dtype=dt, device=dev)                 # own cells only: no ghost I is needed
            self.face_proj = t.zeros((n + plan.n_ghost, fp), dtype=dt, device=dev)
            self.acc = t.zeros((max(plan.n_acc, 1), fp), dtype=dt, device=dev)
            self.scratch = t.zeros((max(plan.n_scratch, 1), prj), dtype=dt, device=dev)
            self.halo = self.face_proj
            j = plan.face_rel[:, :, 0].astype(np.int64)
            # LTS::FaceNeighbors for the projection path: the kernel adds j * prj reals to the entry
            ptr = np.where(plan.kind > 0, self.scratch.data_ptr() + (plan.scratch_row * prj - j * prj) * es,
                           np.where(smaller, self.acc.data_ptr() + plan.acc_row[nb] * fp * es,
                                    self.face_proj.data_ptr() + nb * fp * es))
            scratch_width = prj
        else:
            self.idofs = t.zeros((n + plan.n_ghost, w), dtype=dt, device=dev)   # own cells, then the ghost layer
            self.acc = t.zeros((max(plan.n_acc, 1), w), dtype=dt, device=dev)
            self.scratch = t.zeros((max(plan.n_scratch, 1), w), dtype=dt, device=dev)
            self.halo = self.idofs
            # LTS::FaceNeighbors: where the neighbor flux of (cell, face) reads the neighbor's time-integrated DOFs
            # kind 0: a smaller-cluster neighbor's accumulation buffer, else the neighbor's own integral of this step
            # (same cluster, local or ghost row); kind 1/2: the Taylor scratch row of this (cell, face)
            ptr = np.where(plan.kind > 0, self.scratch.data_ptr() + plan.scratch_row * w * es,
                           np.where(smaller, self.acc.data_ptr() + plan.acc_row[nb] * w * es,
                                    self.idofs.data_ptr() + nb * w * es))
            scratch_width = w
        self.nb_ptrs = t.from_numpy(np.ascontiguousarray(ptr.astype(np.int64))).to(dev)
        self.taylor = {}
        for name, store in (("full", plan.taylor_full), ("sub", plan.taylor_sub)):
            for c, (e, f) in store.items():
                if len(e) == 0:
                    continue
                src = self.derivs.data_ptr() + plan.der_row[nb[e, f]] * ctx.dersize * es
                dst = self.scratch.data_ptr() + plan.scratch_row[e, f] * scratch_width * es
                face_j = t.from_numpy(np.ascontiguousarray(plan.face_rel[e, f, 0].astype(np.uint8))).to(dev)
                self.taylor[(name, c)] = (t.from_numpy(src.astype(np.int64)).to(dev), t.from_numpy(dst.astype(np.int64)).to(dev),
                                          len(e), face_j)
        self.perm = {}
        for c, r in enumerate(plan.ranges):
            rel = plan.face_rel[r.start:r.end].astype(np.int64)
            key = ((rel[:, 0, 0] + 4 * rel[:, 1, 0] + 16 * rel[:, 2, 0] + 64 * rel[:, 3, 0]) * 81
                   + rel[:, 0, 1] + 3 * rel[:, 1, 1] + 9 * rel[:, 2, 1] + 27 * rel[:, 3, 1])
            self.perm[c] = t.from_numpy(np.argsort(key, kind="stable").astype(np.int32)).to(dev)
        self.launches = 0
        self.dist = dist
        self.pending = set()
        # halo: persistent pack buffers (one per cluster and region), communication stream and events created once
        self.send_idx = {c: [(peer, side, t.from_numpy(np.ascontiguousarray(idx)).to(dev),
                              t.empty((len(idx), self.halo.shape[1]), dtype=dt, device=dev)) for peer, side, idx in regs]
                         for c, regs in plan.send.items()}
        # ... derivative families of the copy cells a peer integrates in time, and the ghost rows a larger own cluster reads
        # as a sum over two steps (accumulated on arrival)
        self.send_der = {c: [(peer, side, t.from_numpy(np.ascontiguousarray(rows)).to(dev),
                              t.empty((len(rows), ctx.dersize), dtype=dt, device=dev)) for peer, side, rows in regs]
                         for c, regs in plan.send_der.items()}
        self.ghost_acc = {}
        for c, (g, rows) in plan.ghost_acc.items():
            if len(g):
                assert np.array_equal(rows, np.arange(rows[0], rows[0] + len(rows)))
                src = self.halo.data_ptr() + g * self.halo.shape[1] * es
                self.ghost_acc[c] = (t.from_numpy(src.astype(np.int64)).to(dev), int(rows[0]), len(g))
        self.comm = None
        if dist is not None and dofs.is_cuda and plan.world > 1:
            self.comm = t.cuda.Stream(dev)
            self.ev_pred = {k: t.cuda.Event() for k in range(plan.n_clusters)}
            self.ev_halo = {k: t.cuda.Event() for k in range(plan.n_clusters)}

    def exchange(self, c, reset=True):
        """Halo of cluster c: pack the copy cells' rows (time-integrated DOFs, or face projections) per region into a
        persistent buffer, one NCCL send/recv pair per region, received straight into the ghost rows the neighbor kernel
        reads.  Where a slab boundary separates two clusters the copy cells of the larger one also send their derivative
        families (received straight into derivative rows behind the own ones), and the ghost rows of the smaller one are
        accumulated over its two steps on arrival (DirectGhostTimeCluster.cpp:22-105).  Runs on a communication stream
        behind the prediction of cluster c; only the corrections that read these rows wait for it, so the other
        predictions and corrections overlap the transfer."""
        t, dist, plan = self.torch, self.dist, self.plan
        main = t.cuda.current_stream(self.halo.device)
        self.ev_pred[c].record(main)
        with t.cuda.stream(self.comm):
            self.comm.wait_event(self.ev_pred[c])
            ops = []
            for peer, side, idx, buf in sorted(self.send_idx[c], key=lambda r: r[1]):
                t.index_select(self.halo, 0, idx, out=buf)
                ops.append(dist.P2POp(dist.isend, buf, peer, tag=2 * c + side))
            for peer, side, rows, buf in sorted(self.send_der.get(c, []), key=lambda r: r[1]):
                t.index_select(self.derivs, 0, rows, out=buf)
                ops.append(dist.P2POp(dist.isend, buf, peer, tag=64 + 2 * c + side))
            for peer, side, start, count in sorted(plan.recv[c], key=lambda r: 1 - r[1]):
                ops.append(dist.P2POp(dist.irecv, self.halo[start:start + count], peer, tag=2 * c + (1 - side)))
            for peer, side, row0, count in sorted(plan.recv_der.get(c, []), key=lambda r: 1 - r[1]):
                ops.append(dist.P2POp(dist.irecv, self.derivs[row0:row0 + count], peer, tag=64 + 2 * c + (1 - side)))
            if ops:
                for wk in dist.batch_isend_irecv(ops):
                    wk.wait()
            if c in self.ghost_acc:
                src, a0, cnt = self.ghost_acc[c]
                if self.faceproj:
                    self.ctx.accumulate_rows(self._s(self.acc, a0, a0 + cnt), self.api.pointer_table(src), reset, cnt, self.fp)
                else:
                    self.ctx.lts_accumulate(self._s(self.acc, a0, a0 + cnt), self.api.pointer_table(src), reset, cnt)
                self.launches += 1
            self.ev_halo[c].record(self.comm)
        self.pending.add(c)

    def _s(self, x, a, b):
        return self.api.strided(x[a:b])

    def predict(self, c, r, coeffs, reset):
        ctx, plan = self.ctx, self.plan
        if r.der_end > r.start:
            d0 = int(plan.der_row[r.start])
            ctx.ader(self._s(self.dofs, r.start, r.der_end), self._s(self.idofs, r.start, r.der_end),
                     self._s(self.derivs, d0, d0 + r.der_end - r.start), self._s(self.local_int, r.start, r.der_end), coeffs,
                     r.der_end - r.start)
            self.launches += 1
        if r.end > r.der_end:
            ctx.ader(self._s(self.dofs, r.der_end, r.end), self._s(self.idofs, r.der_end, r.end), None,
                     self._s(self.local_int, r.der_end, r.end), coeffs, r.end - r.der_end)
            self.launches += 1
        if self.faceproj:
            ctx.local_faceproj(self._s(self.dofs, r.start, r.end), self._s(self.idofs, r.start, r.end),
                               self._s(self.local_int, r.start, r.end), self._s(self.face_proj, r.start, r.end),
                               self.face_types[r.start:r.end], r.end - r.start)
        else:
            ctx.local(self._s(self.dofs, r.start, r.end), self._s(self.idofs, r.start, r.end),
                      self._s(self.local_int, r.start, r.end), self.face_types[r.start:r.end], r.end - r.start)
        self.launches += 1
        if r.acc_end > r.start:
            a0 = int(plan.acc_row[r.start])
            if self.faceproj:
                ctx.accumulate_rows(self._s(self.acc, a0, a0 + r.acc_end - r.start), self._s(self.face_proj, r.start, r.acc_end),
                                    reset, r.acc_end - r.start, self.fp)
            else:
                ctx.lts_accumulate(self._s(self.acc, a0, a0 + r.acc_end - r.start), self._s(self.idofs, r.start, r.acc_end),
                                   reset, r.acc_end - r.start)
            self.launches += 1

    def correct(self, c, r, coeffs_full, coeffs_sub):
        ctx, api = self.ctx, self.api
        for k in (c - 1, c, c + 1):   # the ghost rows this correction reads (own cluster, accumulated smaller, derivatives)
            if k in self.pending:
                self.torch.cuda.current_stream(self.halo.device).wait_event(self.ev_halo[k])
                if k == c:
                    self.pending.discard(k)
        for name, coeffs in (("full", coeffs_full), ("sub", coeffs_sub)):
            if (name, c) in self.taylor:
                src, dst, cnt, face_j = self.taylor[(name, c)]
                if self.faceproj:
                    ctx.taylor_faceproj(api.pointer_table(src), api.pointer_table(dst), face_j, coeffs, cnt)
                else:
                    ctx.taylor(api.pointer_table(src), api.pointer_table(dst), coeffs, cnt)
                self.launches += 1
        if self.faceproj:
            ctx.neigh_faceproj(self._s(self.dofs, r.start, r.end), None, self._s(self.neigh_int, r.start, r.end),
                               self.face_types[r.start:r.end], self.face_rel[r.start:r.end], r.end - r.start,
                               nb_ptrs=self.nb_ptrs[r.start:r.end], perm=self.perm[c])
        else:
            ctx.neigh(self._s(self.dofs, r.start, r.end), api.NONE, self._s(self.neigh_int, r.start, r.end),
                      self.face_types[r.start:r.end], self.face_rel[r.start:r.end], r.end - r.start,
                      nb_ptrs=self.nb_ptrs[r.start:r.end], perm=self.perm[c])
        self.launches += 1


__all__ = ["LtsPlan", "ClusterRange", "build_plan", "LtsSolver", "CudaLtsBackend", "integrate_coeffs"]
