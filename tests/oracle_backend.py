"""CPU compute back-end for the distributed host-logic tests: the numpy ORACLE behind the interface of
seissol_b200.distributed.CudaBackend.  Test infrastructure only (lives under tests/, imports oracle/)."""
import numpy as np

from oracle import aderdg as O


class OracleBackend:
    def __init__(self, order, data, coeffs, np_dtype):
        self.d, self.coeffs, self.dt = data, np.asarray(coeffs, dtype=np_dtype), np_dtype
        self.tab = O.load_tables(order, np_dtype)
        self.B, self.ld = data.B, data.ld

    def _logical(self, t, a, b):
        x = t[a:b].numpy().reshape(b - a, 9, self.ld)
        return np.ascontiguousarray(x[:, :, :self.B].transpose(0, 2, 1))

    def _store(self, t, a, b, x):
        out = np.zeros((b - a, 9, self.ld), dtype=self.dt)
        out[:, :, :self.B] = x.transpose(0, 2, 1)
        t[a:b] = self.d.torch.from_numpy(out.reshape(b - a, 9 * self.ld))

    def _int(self, a, b):
        li = self.d.local_int[a:b].numpy()
        star = O.star_dense(li[:, :72].reshape(-1, 3, 24), self.tab["star_pattern"])
        return star, O.flux_dense(li[:, 72:].reshape(-1, 4, 81))

    def local(self, a, b, stream=None):
        if b <= a:
            return
        d = self.d
        Q = self._logical(d.dofs, a, b)
        star, aplus = self._int(a, b)
        I = O.ader(self.tab, Q, star, self.coeffs)
        self._store(d.buffers, a, b, I)
        self._store(d.dofs, a, b, O.local(self.tab, Q, I, star, aplus, d.face_types[a:b].numpy()))

    def neigh(self, stream=None):
        d = self.d
        n = d.cells
        Q = self._logical(d.dofs, 0, n)
        Iall = self._logical(d.buffers, 0, d.buffers.shape[0])
        aminus = O.flux_dense(d.neigh_int.numpy().reshape(-1, 4, 81))
        self._store(d.dofs, 0, n, O.neighbor(self.tab, Q, Iall, d.neighbors.numpy(), aminus, d.face_rel.numpy(),
                                             d.face_types.numpy()))
