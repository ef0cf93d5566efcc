"""Partitioned mesh across the GPUs of one box: halo exchange of copy-layer buffers overlapped with the
interior element update (host logic above the C ABI; torch.distributed is the plumbing).

What it replaces in the reference: the MPI ghost/copy-layer exchange of `src/Parallel` +
`src/Solver/TimeStepping/{TimeCluster,GhostTimeClusterWithCopy,DirectGhostTimeCluster}.cpp` -- per time cluster,
copy-layer cells are predicted first, their time-integrated buffers travel to the neighbouring rank's ghost layer
while the interior cells are predicted, and the neighbor (corrector) kernel waits for the ghost data
(SURVEY.md §8e).  B200-native shape: one process per GPU, copy and ghost regions are contiguous slices of the
`buffers` array on both ends (seissol_b200/partition.py), so a region moves with ONE NCCL send/recv pair
straight from/to the memory the kernels use (no pack/unpack kernels, no host staging), issued on a second
CUDA stream and fenced with events; NVLink/NVSwitch carries it while the interior kernels run.

The compute back-end is injected (`CudaBackend` = the C ABI; the CPU tests inject an oracle back-end from
tests/), so that this module stays free of any CPU compute path.
"""
import numpy as np

from . import mesh as _mesh

_M1 = 0x9E3779B97F4A7C15
_M2 = 0xC2B2AE3D27D4EB4F
_M3 = 0xFF51AFD7ED558CCD
_M4 = 0xC4CEB9FE1A85EC53


def _s64(x):
    """Two's-complement int64 value of an unsigned 64-bit constant."""
    return x - (1 << 64) if x >= (1 << 63) else x


def element_values_torch(torch, gid, count, stream, scale, dtype, chunk=1 << 16):
    """mesh.element_values on a torch device: int64 arithmetic wraps like uint64, logical shifts are emulated.
    gid: int64 tensor.  Returns [len(gid), count] of `dtype` in [-scale, scale)."""
    dev = gid.device
    col = (torch.arange(count, dtype=torch.int64, device=dev) + (stream << 20)) * _s64(_M2)
    out = torch.empty((gid.shape[0], count), dtype=dtype, device=dev)
    for a in range(0, gid.shape[0], chunk):
        h = gid[a:a + chunk, None] * _s64(_M1) + col[None, :]
        h = h ^ ((h >> 33) & 0x7FFFFFFF)
        h = h * _s64(_M3)
        h = h ^ ((h >> 33) & 0x7FFFFFFF)
        h = h * _s64(_M4)
        h = h ^ ((h >> 33) & 0x7FFFFFFF)
        u = ((h >> 40) & 0xFFFFFF).to(torch.float64) / float(1 << 24)
        out[a:a + chunk] = ((2.0 * u - 1.0) * scale).to(dtype)
    return out


class MeshData:
    """Per-rank storage of the partitioned mesh in the device layout of seissol_b200/proxy.py::ProxyData:
    dofs [n_own, 9*ld], buffers [n_own + n_ghost, 9*ld] (ghost layer behind the own cells), local_int [n_own, 396],
    neigh_int [n_own, 324], neighbors [n_own, 4] (index into buffers), face_rel, face_types.
    Values are a function of the global element id only (mesh.element_values), so every partitioning of the same
    mesh computes on identical data."""

    def __init__(self, torch, layout, B, ld, dtype, device, scale=1e-2):
        t = torch
        self.torch, self.layout, self.B, self.ld = t, layout, B, ld
        self.device, self.dtype = t.device(device), dtype
        n_own, n_all = layout.n_own, layout.n_own + layout.n_ghost
        self.cells = n_own
        gid = t.from_numpy(layout.gid[:n_own]).to(self.device)
        q = element_values_torch(t, gid, 9 * B, 0, 1.0, dtype).reshape(n_own, 9, B)
        dofs = t.zeros((n_own, 9, ld), dtype=dtype, device=self.device)
        dofs[:, :, :B] = q
        self.dofs = dofs.reshape(n_own, 9 * ld).contiguous()
        self.buffers = t.zeros((n_all, 9 * ld), dtype=dtype, device=self.device)
        self.local_int = element_values_torch(t, gid, 396, 1, scale, dtype).contiguous()
        self.neigh_int = element_values_torch(t, gid, 324, 2, scale, dtype).contiguous()
        self.neighbors = t.from_numpy(np.ascontiguousarray(layout.neighbors)).to(self.device)
        self.face_rel = t.from_numpy(np.ascontiguousarray(layout.face_rel)).to(self.device)
        self.face_types = t.from_numpy(np.ascontiguousarray(layout.face_types)).to(self.device)
        self.perm = None

    def sort_by_face_relation(self):
        """Processing order of the neighbor kernel: uniform (j, h) operator selection inside a warp
        (see proxy.ProxyData.sort_by_face_relation); boundary faces sort as relation (0, 0)."""
        t = self.torch
        j = self.face_rel[:, :, 0].to(t.int64)
        h = self.face_rel[:, :, 1].to(t.int64)
        key = (j[:, 0] + 4 * j[:, 1] + 16 * j[:, 2] + 64 * j[:, 3]) * 81 + h[:, 0] + 3 * h[:, 1] + 9 * h[:, 2] + 27 * h[:, 3]
        self.perm = t.argsort(key, stable=True).to(t.int32).contiguous()
        return self.perm

    def logical_dofs(self):
        """[n_own, B, 9] numpy (the oracle's layout)."""
        x = self.dofs.detach().cpu().numpy().reshape(self.cells, 9, self.ld)
        return np.ascontiguousarray(x[:, :, :self.B].transpose(0, 2, 1))


class CudaBackend:
    """The element-update kernels through the C ABI (seissol_b200/api.py) on sub-ranges of the own cells.

    faceproj (default at order >= 6, SSB_FACEPROJ=0 switches it off): the local kernel stores the face projections
    F_f = fMrT_f I of every own cell, the neighbor kernel gathers those blocks, and the halo carries the copy cells'
    face projections instead of their time-integrated DOFs (`halo_tensor`)."""

    def __init__(self, ctx, data, coeffs, faceproj=None):
        import os
        from . import api
        self.api, self.ctx, self.d, self.coeffs = api, ctx, data, coeffs
        if faceproj is None:
            faceproj = os.environ.get("SSB_FACEPROJ", "1" if ctx.order >= 6 else "0") != "0"
        self.faceproj = bool(faceproj)
        if self.faceproj and getattr(data, "face_proj", None) is None:
            data.face_proj = data.torch.zeros((data.buffers.shape[0], ctx.face_proj_reals()), dtype=data.dtype,
                                              device=data.device)

    def halo_tensor(self):
        """The per-cell array whose copy regions travel to the peers' ghost regions."""
        return self.d.face_proj if self.faceproj else self.d.buffers

    def _rows(self, t, a, b):
        return self.api.strided(t[a:b])

    def local(self, a, b, stream=None):
        """ader + volume + local flux for own cells [a, b): writes buffers[a:b] (and face_proj[a:b]) and dofs[a:b]."""
        if b <= a:
            return
        d = self.d
        self.ctx.ader(self._rows(d.dofs, a, b), self._rows(d.buffers, a, b), None, self._rows(d.local_int, a, b),
                      self.coeffs, b - a, stream=stream)
        if self.faceproj:
            self.ctx.local_faceproj(self._rows(d.dofs, a, b), self._rows(d.buffers, a, b), self._rows(d.local_int, a, b),
                                    self._rows(d.face_proj, a, b), d.face_types[a:b], b - a, stream=stream)
        else:
            self.ctx.local(self._rows(d.dofs, a, b), self._rows(d.buffers, a, b), self._rows(d.local_int, a, b),
                           d.face_types[a:b], b - a, stream=stream)

    def neigh(self, stream=None):
        """neighbor flux for all own cells; neighbors index the whole per-cell array (own + ghost)."""
        d = self.d
        if self.faceproj:
            self.ctx.neigh_faceproj(self.api.strided(d.dofs), self.api.strided(d.face_proj), self.api.strided(d.neigh_int),
                                    d.face_types, d.face_rel, d.cells, nb_idx=d.neighbors, perm=d.perm, stream=stream)
            return
        self.ctx.neigh(self.api.strided(d.dofs), self.api.strided(d.buffers), self.api.strided(d.neigh_int),
                       d.face_types, d.face_rel, d.cells, nb_idx=d.neighbors, perm=d.perm, stream=stream)

    launches_per_step = 5   # (ader, local) x (copy, interior) + neigh


class HaloExchange:
    """Moves every copy region into the matching ghost region of the peer (torch.distributed point-to-point:
    NCCL over NVLink on the GPUs, gloo in the CPU tests).  Regions are contiguous row slices of `buffers`."""

    def __init__(self, torch, dist, layout, buffers, group=None):
        self.torch, self.dist, self.layout, self.buffers, self.group = torch, dist, layout, buffers, group
        # issue order must match between the two ends when both regions of a rank go to the same peer (ring of
        # two): my region on side s is the peer's ghost region on side 1-s
        self.sends = sorted(layout.send, key=lambda r: (r.cluster, r.side))
        self.recvs = sorted(layout.recv, key=lambda r: (r.cluster, 1 - r.side))
        self.bytes_per_exchange = sum(r.count for r in self.sends) * buffers.shape[1] * buffers.element_size()

    def _ops(self, clusters=None):
        dist, buf = self.dist, self.buffers
        ops = []
        for r in self.sends:
            if clusters is None or r.cluster in clusters:
                ops.append(dist.P2POp(dist.isend, buf[r.start:r.start + r.count], r.peer, self.group,
                                      tag=2 * r.cluster + r.side))
        for r in self.recvs:
            if clusters is None or r.cluster in clusters:
                ops.append(dist.P2POp(dist.irecv, buf[r.start:r.start + r.count], r.peer, self.group,
                                      tag=2 * r.cluster + (1 - r.side)))
        return ops

    def exchange(self, clusters=None):
        """Enqueue the exchange on the current stream (NCCL) / run it to completion (gloo)."""
        if self.layout.world == 1:
            return
        ops = self._ops(clusters)
        if not ops:
            return
        for w in self.dist.batch_isend_irecv(ops):
            w.wait()


class PartitionedSolver:
    """Global-time-stepping update of one rank's partition with the exchange overlapped:

        main stream : local(copy cells) | local(interior cells) ........ | wait halo | neigh(all own cells)
        comm stream :                   | wait copy | send/recv regions  |

    (`TimeCluster::computeLocalIntegration` for the copy layer before the interior layer and the ghost cluster's
    send/receive in between, src/Solver/TimeStepping/TimeCluster.cpp:440-538, AbstractGhostTimeCluster.cpp.)"""

    def __init__(self, torch, layout, data, backend, exchange, overlap=True):
        self.torch, self.layout, self.d, self.backend, self.x = torch, layout, data, backend, exchange
        self.cuda = data.device.type == "cuda"
        self.overlap = overlap and self.cuda and layout.world > 1
        if self.cuda:
            self.main = torch.cuda.current_stream(data.device)
            self.comm = torch.cuda.Stream(data.device) if self.overlap else self.main
            self.ev_copy = torch.cuda.Event()
            self.ev_halo = torch.cuda.Event()

    def step(self):
        lay, be = self.layout, self.backend
        if not self.cuda:
            be.local(0, lay.n_copy)
            self.x.exchange()
            be.local(lay.n_copy, lay.n_own)
            be.neigh()
            return
        torch = self.torch
        be.local(0, lay.n_copy)
        if self.overlap:
            self.ev_copy.record(self.main)
            self.comm.wait_event(self.ev_copy)
            with torch.cuda.stream(self.comm):
                self.x.exchange()
                self.ev_halo.record(self.comm)
            be.local(lay.n_copy, lay.n_own)
            self.main.wait_event(self.ev_halo)
        else:
            self.x.exchange()
            be.local(lay.n_copy, lay.n_own)
        be.neigh()


def build_partitioned(torch, dist, nx, ny, nz, rank, world, B, ld, dtype, device, n_clusters=1, scale=1e-2):
    """Layout + data of `rank` (convenience used by bench.py and the tests)."""
    from .partition import build_layout
    layout = build_layout(nx, ny, nz, rank, world, n_clusters)
    data = MeshData(torch, layout, B, ld, dtype, device, scale)
    return layout, data


__all__ = ["MeshData", "CudaBackend", "HaloExchange", "PartitionedSolver", "build_partitioned",
           "element_values_torch"]
_ = _mesh
