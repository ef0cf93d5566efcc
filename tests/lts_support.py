"""Test support for the LTS scheduler (tests only; imports oracle/):

* `brute_force_macro_step` — the DEFINITION of clustered LTS, with no buffers / accumulation / sub-interval plumbing:
  every cell keeps the history of its Taylor predictors, and the neighbor flux of a step [a, b] uses
  int_a^b of the neighbor's piecewise predictor, summed over whatever pieces overlap [a, b].
* `NumpyLtsBackend` — the numpy oracle behind the predict/correct interface of seissol_b200.lts.CudaLtsBackend, with
  the same storage plumbing (own integrals, accumulation buffers, derivative rows, Taylor scratch rows)."""
import numpy as np

from oracle import aderdg as O
from seissol_b200.lts import integrate_coeffs


def brute_force_macro_step(tab, order, plan, Q, star, Aplus, Aminus, dt0=1.0):
    n, C = plan.n, plan.n_clusters
    Q = Q.copy()
    fine = 2 ** (C - 1)
    history = [[] for _ in range(n)]          # per cell: (start, length, derivatives list of [B, 9])
    cl, nb = plan.cluster, plan.neighbors
    own4 = np.arange(4 * n).reshape(n, 4)
    for s in range(fine):
        for c in range(C):
            if s % (2 ** c):
                continue
            cells = np.nonzero(cl == c)[0]
            if len(cells) == 0:
                continue
            coeffs = integrate_coeffs(order, 0.0, dt0 * 2 ** c).astype(Q.dtype)
            I, ders = O.ader(tab, Q[cells], star[cells], coeffs, True)
            Q[cells] = O.local(tab, Q[cells], I, star[cells], Aplus[cells], plan.face_types[cells])
            for i, e in enumerate(cells):
                history[e].append((s, 2 ** c, [d[i] for d in ders]))
        for c in range(C):
            if (s + 1) % (2 ** c):
                continue
            cells = np.nonzero(cl == c)[0]
            if len(cells) == 0:
                continue
            a, b = s + 1 - 2 ** c, s + 1
            Inb = np.zeros((len(cells), 4) + Q.shape[1:], dtype=Q.dtype)
            for i, e in enumerate(cells):
                for f in range(4):
                    if plan.face_types[e, f] != 0:
                        continue
                    for (t0, ln, ders) in history[nb[e, f]]:
                        lo, hi = max(a, t0), min(b, t0 + ln)
                        if hi > lo:
                            cf = integrate_coeffs(order, (lo - t0) * dt0, (hi - t0) * dt0).astype(Q.dtype)
                            Inb[i, f] += O.taylor([d[None] for d in ders], cf)[0]
            m = len(cells)
            Q[cells] = O.neighbor(tab, Q[cells], Inb.reshape(4 * m, *Q.shape[1:]), own4[:m], Aminus[cells],
                                  plan.face_rel[cells], plan.face_types[cells])
    return Q


class NumpyLtsBackend:
    def __init__(self, tab, order, plan, Q, star, Aplus, Aminus, dist=None):
        self.tab, self.order, self.plan = tab, order, plan
        self.Q, self.star, self.Aplus, self.Aminus = Q.copy(), star, Aplus, Aminus
        shp = Q.shape[1:]
        self.idofs = np.zeros((plan.n + plan.n_ghost,) + Q.shape[1:], dtype=Q.dtype)   # own cells, then the ghost layer
        self.dist = dist
        self.acc = np.zeros((max(plan.n_acc, 1),) + shp, dtype=Q.dtype)
        self.derivs = [np.zeros((max(plan.n_der, 1),) + shp, dtype=Q.dtype) for _ in range(order)]
        self.scratch = np.zeros((max(plan.n_scratch, 1),) + shp, dtype=Q.dtype)

    def predict(self, c, r, coeffs, reset):
        p, sl = self.plan, slice(r.start, r.end)
        I, ders = O.ader(self.tab, self.Q[sl], self.star[sl], coeffs.astype(self.Q.dtype), True)
        self.idofs[sl] = I
        nd = r.der_end - r.start
        if nd:
            d0 = p.der_row[r.start]
            for d in range(self.order):
                self.derivs[d][d0:d0 + nd] = ders[d][:nd]
        self.Q[sl] = O.local(self.tab, self.Q[sl], I, self.star[sl], self.Aplus[sl], p.face_types[sl])
        na = r.acc_end - r.start
        if na:
            a0 = p.acc_row[r.start]
            self.acc[a0:a0 + na] = I[:na] if reset else self.acc[a0:a0 + na] + I[:na]

    def exchange(self, c, reset=True):
        import torch
        dist, p = self.dist, self.plan
        ops, bufs, dbufs = [], [], []
        tdt = torch.from_numpy(self.idofs[:1]).dtype
        for peer, side, idx in sorted(p.send[c], key=lambda r: r[1]):
            ops.append(dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(self.idofs[idx])), peer, tag=2 * c + side))
        for peer, side, rows in sorted(p.send_der[c], key=lambda r: r[1]):
            fam = np.stack([d[rows] for d in self.derivs], 1)           # [cells, order, B, 9]
            ops.append(dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(fam)), peer, tag=64 + 2 * c + side))
        for peer, side, start, count in sorted(p.recv[c], key=lambda r: 1 - r[1]):
            t = torch.empty((count,) + self.idofs.shape[1:], dtype=tdt)
            bufs.append((start, count, t))
            ops.append(dist.P2POp(dist.irecv, t, peer, tag=2 * c + (1 - side)))
        for peer, side, row0, count in sorted(p.recv_der[c], key=lambda r: 1 - r[1]):
            t = torch.empty((count, self.order) + self.idofs.shape[1:], dtype=tdt)
            dbufs.append((row0, count, t))
            ops.append(dist.P2POp(dist.irecv, t, peer, tag=64 + 2 * c + (1 - side)))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for start, count, t in bufs:
            self.idofs[start:start + count] = t.numpy()
        for row0, count, t in dbufs:
            for d in range(self.order):
                self.derivs[d][row0:row0 + count] = t.numpy()[:, d]
        g, rows = p.ghost_acc[c]
        if len(g):      # ghost cells read by a larger own cluster: sum of their two steps
            self.acc[rows] = self.idofs[g] if reset else self.acc[rows] + self.idofs[g]

    def correct(self, c, r, coeffs_full, coeffs_sub):
        p, sl = self.plan, slice(r.start, r.end)
        nb = p.neighbors
        for store, coeffs in ((p.taylor_full, coeffs_full), (p.taylor_sub, coeffs_sub)):
            e, f = store[c]
            if len(e):
                rows = p.der_row[nb[e, f]]
                self.scratch[p.scratch_row[e, f]] = O.taylor([d[rows] for d in self.derivs], coeffs.astype(self.Q.dtype))
        m = r.end - r.start
        Inb = np.zeros((m, 4) + self.Q.shape[1:], dtype=self.Q.dtype)
        for f in range(4):
            k = p.kind[sl, f]
            n_f = nb[sl, f]
            buf = self.idofs[n_f]                      # same cluster (local or ghost row): the integral of this step
            accm = (k == 0) & (p.cluster_ext[n_f] == c - 1) & (p.face_types[sl, f] == 0)
            buf[accm] = self.acc[p.acc_row[n_f][accm]]  # smaller cluster (own or ghost cell): sum over its two steps
            tm = k > 0
            buf[tm] = self.scratch[p.scratch_row[sl, f][tm]]
            Inb[:, f] = buf
        own4 = np.arange(4 * m).reshape(m, 4)
        self.Q[sl] = O.neighbor(self.tab, self.Q[sl], Inb.reshape(4 * m, *self.Q.shape[1:]), own4, self.Aminus[sl],
                                p.face_rel[sl], p.face_types[sl])
