"""ctypes binding of oracle/_ref/libssref_<cfg>.so — TEST INFRASTRUCTURE ONLY.

The library holds the reference's own generated CPU kernels plus our thin
wrapper restatement (oracle/ref_harness.cpp, built by oracle/build_ref.py).
Arrays cross this boundary in the LOGICAL layout of oracle/aderdg.py
([n,B,9]); conversion to the reference's padded column-major layout happens here.
"""
import ctypes
import os

import numpy as np

from . import aderdg as O

HERE = os.path.dirname(os.path.abspath(__file__))


def available(cfg):
    return os.path.exists(os.path.join(HERE, "_ref", f"libssref_{cfg}.so"))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class RefLib:
    def __init__(self, cfg):
        path = os.path.join(HERE, "_ref", f"libssref_{cfg}.so")
        self.lib = ctypes.CDLL(path)
        info = np.zeros(8, dtype=np.int64)
        self.lib.ssref_info(_p(info))
        self.order, size, self.qsize, self.isize, self.dersize, self.qinterp, self.B, self.ld = \
            [int(x) for x in info]
        self.dtype = np.float32 if size == 4 else np.float64
        offs = np.zeros(self.order, dtype=np.int64)
        sizes = np.zeros(self.order, dtype=np.int64)
        self.lib.ssref_der_layout(_p(offs), _p(sizes))
        self.der_offsets, self.der_sizes = offs, sizes
        self.lib.ssref_proxy.restype = ctypes.c_double
        self.ppad = self.qinterp // 9

    # -- layout helpers
    def to_ref(self, x):
        return O.to_ref_layout(np.ascontiguousarray(x, dtype=self.dtype), self.ld)

    def from_ref(self, x):
        return O.from_ref_layout(x, self.B, self.ld)

    def ders_to_ref(self, ders):
        n = ders[0].shape[0]
        out = np.zeros((n, self.dersize), dtype=self.dtype)
        for d, dq in enumerate(ders):
            ld = int(self.der_sizes[d]) // 9
            rows = min(ld, O.nbasis(self.order - d))
            blk = np.zeros((n, 9, ld), dtype=self.dtype)
            blk[:, :, :rows] = dq[:, :rows, :].transpose(0, 2, 1)
            out[:, self.der_offsets[d]:self.der_offsets[d] + 9 * ld] = blk.reshape(n, -1)
        return out

    def ders_from_ref(self, x):
        n = x.shape[0]
        ders = []
        for d in range(self.order):
            ld = int(self.der_sizes[d]) // 9
            rows = O.nbasis(self.order - d)
            blk = x[:, self.der_offsets[d]:self.der_offsets[d] + 9 * ld].reshape(n, 9, ld)
            dq = np.zeros((n, self.B, 9), dtype=self.dtype)
            dq[:, :rows] = blk[:, :, :rows].transpose(0, 2, 1)
            ders.append(dq)
        return ders

    def flops(self):
        nz = np.zeros(55, dtype=np.int64)
        hw = np.zeros(55, dtype=np.int64)
        self.lib.ssref_flops(_p(nz), _p(hw))
        return {"derivative": (int(nz[0]), int(hw[0])), "volume": (int(nz[1]), int(hw[1])),
                "localFlux": (nz[2:6].copy(), hw[2:6].copy()),
                "neighboringFlux": (nz[6:54].copy(), hw[6:54].copy()),
                "taylor": (int(nz[54]), int(hw[54]))}

    # -- kernels (logical layout in/out)
    def ader(self, Q, star, coeffs, want_derivatives=False):
        n = Q.shape[0]
        q = self.to_ref(Q)
        I = np.zeros((n, self.isize), dtype=self.dtype)
        der = np.zeros((n, self.dersize), dtype=self.dtype) if want_derivatives else None
        c = np.ascontiguousarray(coeffs, dtype=self.dtype)
        s = np.ascontiguousarray(star, dtype=self.dtype)
        self.lib.ssref_ader(ctypes.c_int64(n), _p(c), _p(q), _p(s), _p(I), _p(der))
        if want_derivatives:
            return self.from_ref(I), self.ders_from_ref(der)
        return self.from_ref(I)

    def taylor(self, ders, coeffs):
        n = ders[0].shape[0]
        der = self.ders_to_ref(ders)
        I = np.zeros((n, self.isize), dtype=self.dtype)
        c = np.ascontiguousarray(coeffs, dtype=self.dtype)
        self.lib.ssref_taylor(ctypes.c_int64(n), _p(c), _p(der), _p(I))
        return self.from_ref(I)

    def local(self, Q, I, star, Aplus, face_types):
        n = Q.shape[0]
        q, i = self.to_ref(Q), self.to_ref(I)
        s = np.ascontiguousarray(star, dtype=self.dtype)
        a = np.ascontiguousarray(Aplus, dtype=self.dtype)
        ft = np.ascontiguousarray(face_types, dtype=np.uint8)
        self.lib.ssref_local(ctypes.c_int64(n), _p(q), _p(i), _p(s), _p(a), _p(ft))
        return self.from_ref(q)

    def neighbor(self, Q, Iall, neighbors, Aminus, face_rel, face_types, dr=None):
        n = Q.shape[0]
        q, i = self.to_ref(Q), self.to_ref(Iall)
        a = np.ascontiguousarray(Aminus, dtype=self.dtype)
        nb = np.ascontiguousarray(neighbors, dtype=np.int64)
        fr = np.ascontiguousarray(face_rel, dtype=np.uint8)
        ft = np.ascontiguousarray(face_types, dtype=np.uint8)
        if dr is None:
            args = [None] * 5
        else:
            god = self.points_to_ref(dr["godunov"])
            fs = np.ascontiguousarray(dr["fluxSolver"].transpose(0, 2, 1), dtype=self.dtype)
            args = [god, fs, np.ascontiguousarray(dr["index"], dtype=np.int64),
                    np.ascontiguousarray(dr["side"], dtype=np.uint8),
                    np.ascontiguousarray(dr["rel"], dtype=np.uint8)]
        self.lib.ssref_neigh(ctypes.c_int64(n), _p(q), _p(i), _p(nb), _p(a), _p(fr), _p(ft),
                             *[_p(x) for x in args])
        return self.from_ref(q)

    def nodal_bc(self, Q, I, Aminus, bc):
        """ssref_nodal_bc: the reference's generated projectToNodalBoundaryRotated / createEasiBoundaryGhostCells /
        localFluxNodal kernels on the boundary faces of `bc` (oracle.aderdg.local_nodal_bc)."""
        q, i = self.to_ref(Q), self.to_ref(I)
        m = len(bc["cell"])
        c = lambda x: np.ascontiguousarray(x, dtype=self.dtype)  # noqa: E731
        self.lib.ssref_nodal_bc(ctypes.c_int64(m), _p(q), _p(i), _p(c(Aminus)), _p(np.ascontiguousarray(bc["cell"], dtype=np.int64)),
                                _p(np.ascontiguousarray(bc["face"], dtype=np.uint8)),
                                _p(np.ascontiguousarray(bc["kind"], dtype=np.uint8)), _p(c(bc["tinv"])), _p(c(bc["easiMap"])),
                                _p(c(bc["easiConst"])), _p(c(bc["displacement"])),
                                _p(np.ascontiguousarray(bc["rhoG"], dtype=np.float64)))
        return self.from_ref(q)

    def points_to_ref(self, x):
        """[m, P, 9] -> [m, 9*Ppad] (point fastest)."""
        m, P, _ = x.shape
        out = np.zeros((m, 9, self.ppad), dtype=self.dtype)
        out[:, :, :P] = x.transpose(0, 2, 1)
        return out.reshape(m, -1)

    def dr_interpolate(self, ders_plus, ders_minus, tinvT, face_info, coeffs):
        n = ders_plus[0].shape[0]
        P = (self.order + 1) ** 2
        dp, dm = self.ders_to_ref(ders_plus), self.ders_to_ref(ders_minus)
        # TinvT dense [n,9(q),9(p)] -> column-major TinvT[q + 9p]
        t = np.ascontiguousarray(tinvT.transpose(0, 2, 1), dtype=self.dtype)
        fi = np.ascontiguousarray(face_info, dtype=np.uint8)
        c = np.ascontiguousarray(coeffs, dtype=self.dtype)
        qp = np.zeros((n, self.order, self.qinterp), dtype=self.dtype)
        qm = np.zeros_like(qp)
        self.lib.ssref_dr_interp(ctypes.c_int64(n), _p(dp), _p(dm), _p(t), _p(fi), _p(c),
                                 _p(qp), _p(qm))
        conv = lambda x: np.ascontiguousarray(
            x.reshape(n, self.order, 9, self.ppad)[:, :, :, :P].transpose(0, 1, 3, 2))
        return conv(qp), conv(qm)

    def resample(self, x):
        n, P = x.shape
        xin = np.zeros((n, self.ppad), dtype=self.dtype)
        xin[:, :P] = x
        out = np.zeros_like(xin)
        self.lib.ssref_resample(ctypes.c_int64(n), _p(xin), _p(out))
        return out[:, :P]

    def proxy(self, kernel, d, coeffs, steps):
        """Timed proxy passes on reference-layout copies; returns (seconds, Q, I)."""
        kid = {"all": 0, "local": 1, "neigh": 2, "ader": 3, "localwoader": 4}[kernel]
        n = d["Q"].shape[0]
        q = self.to_ref(d["Q"])
        buf = self.to_ref(d["I"]) if "I" in d else np.zeros((n, self.isize), dtype=self.dtype)
        c = np.ascontiguousarray(coeffs, dtype=self.dtype)
        arrs = [np.ascontiguousarray(d["star"], dtype=self.dtype),
                np.ascontiguousarray(d["Aplus"], dtype=self.dtype),
                np.ascontiguousarray(d["Aminus"], dtype=self.dtype),
                np.ascontiguousarray(d["neighbors"], dtype=np.int64),
                np.ascontiguousarray(d["face_rel"], dtype=np.uint8),
                np.ascontiguousarray(d["face_types"], dtype=np.uint8)]
        t = self.lib.ssref_proxy(ctypes.c_int(kid), ctypes.c_int64(n), ctypes.c_int(steps), _p(c),
                                 _p(q), _p(buf), *[_p(a) for a in arrs])
        return t, self.from_ref(q), self.from_ref(buf)

    def threads(self):
        return int(self.lib.ssref_threads())

    def set_threads(self, n):
        """OpenMP threads of the reference kernels (torchrun exports OMP_NUM_THREADS=1 to every rank)."""
        self.lib.ssref_set_threads(ctypes.c_int(int(n)))
        return self.threads()

    def dr_interpolate_timed(self, ders_plus, ders_minus, tinvT, face_info, coeffs, passes=1):
        """Seconds per pass of the reference's batched space-time interpolation on the host cores (no layout conversion
        inside the timed region)."""
        import time
        n = ders_plus[0].shape[0]
        dp, dm = self.ders_to_ref(ders_plus), self.ders_to_ref(ders_minus)
        t = np.ascontiguousarray(tinvT.transpose(0, 2, 1), dtype=self.dtype)
        fi = np.ascontiguousarray(face_info, dtype=np.uint8)
        c = np.ascontiguousarray(coeffs, dtype=self.dtype)
        qp = np.zeros((n, self.order, self.qinterp), dtype=self.dtype)
        qm = np.zeros_like(qp)
        args = (ctypes.c_int64(n), _p(dp), _p(dm), _p(t), _p(fi), _p(c), _p(qp), _p(qm))
        self.lib.ssref_dr_interp(*args)
        t0 = time.perf_counter()
        for _ in range(passes):
            self.lib.ssref_dr_interp(*args)
        return (time.perf_counter() - t0) / passes


class ViscoRefLib:
    """The reference's generated viscoelastic2 kernels (oracle/ref_harness_visco.cpp).  Arrays cross in the
    logical layout of oracle/visco.py; conversion to the reference's padded column-major layout happens here."""

    def __init__(self, cfg):
        path = os.path.join(HERE, "_ref", f"libssref_{cfg}.so")
        self.lib = ctypes.CDLL(path)
        info = np.zeros(12, dtype=np.int64)
        self.lib.ssrefv_info(_p(info))
        (self.order, size, self.qsize, self.qanesize, self.qextsize, self.B, self.ld, self.starsize, self.asize,
         self.esize, self.wsize, self.Wsize) = [int(x) for x in info]
        self.dtype = np.float32 if size == 4 else np.float64
        self.mech = self.wsize
        self.lib.ssrefv_proxy_all.restype = ctypes.c_double
        assert self.starsize == 33 and self.asize == 135 and self.Wsize == self.mech

    def to_ref(self, x):
        return O.to_ref_layout(np.ascontiguousarray(x, dtype=self.dtype), self.ld)

    def from_ref(self, x):
        return O.from_ref_layout(x, self.B, self.ld)

    def ane_to_ref(self, x):
        """[n, B, 6, M] -> [n, M*6*ld] with index k + ld (p + 6 m)."""
        n, b, p, m = x.shape
        out = np.zeros((n, m, p, self.ld), dtype=self.dtype)
        out[:, :, :, :b] = x.transpose(0, 3, 2, 1)
        return out.reshape(n, -1)

    def ane_from_ref(self, x):
        n = x.shape[0]
        return np.ascontiguousarray(x.reshape(n, self.mech, 6, self.ld)[:, :, :, :self.B].transpose(0, 3, 2, 1))

    def flops(self):
        nz = np.zeros(56, dtype=np.int64)
        self.lib.ssrefv_flops(_p(nz))
        return {"derivative": int(nz[0]), "volumeExt": int(nz[1]), "localFluxExt": [int(v) for v in nz[2:6]],
                "local": int(nz[6]), "neighborFluxExt": [int(v) for v in nz[7:55]], "neighbor": int(nz[55])}

    def _c(self, x):
        return np.ascontiguousarray(x, dtype=self.dtype)

    def ader(self, Q, Qane, star, E, w, W, coeffs):
        n = Q.shape[0]
        q, qa = self.to_ref(Q), self.ane_to_ref(Qane)
        I = np.zeros_like(q)
        Ia = np.zeros_like(qa)
        c = self._c(coeffs)
        self.lib.ssrefv_ader(ctypes.c_int64(n), _p(c), _p(q), _p(qa), _p(self._c(star)), _p(self._c(E)), _p(self._c(w)),
                             _p(self._c(W)), _p(I), _p(Ia))
        return self.from_ref(I), self.ane_from_ref(Ia)

    def local(self, Q, Qane, I, Iane, star, Aplus, E, w, W, face_types):
        n = Q.shape[0]
        q, qa = self.to_ref(Q), self.ane_to_ref(Qane)
        ft = np.ascontiguousarray(face_types, dtype=np.uint8)
        self.lib.ssrefv_local(ctypes.c_int64(n), _p(q), _p(qa), _p(self.to_ref(I)), _p(self.ane_to_ref(Iane)),
                              _p(self._c(star)), _p(self._c(Aplus)), _p(self._c(E)), _p(self._c(w)), _p(self._c(W)), _p(ft))
        return self.from_ref(q), self.ane_from_ref(qa)

    def neigh(self, Q, Qane, Iall, neighbors, Aminus, w, face_rel, face_types):
        n = Q.shape[0]
        q, qa = self.to_ref(Q), self.ane_to_ref(Qane)
        nb = np.ascontiguousarray(neighbors, dtype=np.int64)
        fr = np.ascontiguousarray(face_rel, dtype=np.uint8)
        ft = np.ascontiguousarray(face_types, dtype=np.uint8)
        self.lib.ssrefv_neigh(ctypes.c_int64(n), _p(q), _p(qa), _p(self.to_ref(Iall)), _p(nb), _p(self._c(Aminus)),
                              _p(self._c(w)), _p(fr), _p(ft))
        return self.from_ref(q), self.ane_from_ref(qa)

    def neigh_dr(self, Q, Qane, Iall, neighbors, Aminus, w, face_rel, face_types, dr):
        """ssrefv_neigh with the DynamicRupture branch; dr as oracle.visco.neighbor."""
        n = Q.shape[0]
        q, qa = self.to_ref(Q), self.ane_to_ref(Qane)
        nb = np.ascontiguousarray(neighbors, dtype=np.int64)
        fr = np.ascontiguousarray(face_rel, dtype=np.uint8)
        ft = np.ascontiguousarray(face_types, dtype=np.uint8)
        self.lib.ssrefv_qinterp_size.restype = ctypes.c_int64
        ldp = int(self.lib.ssrefv_qinterp_size()) // 9
        g = dr["godunov"]
        god = np.zeros((g.shape[0], 9, ldp), dtype=self.dtype)
        god[:, :, :g.shape[1]] = g.transpose(0, 2, 1)
        fs = np.ascontiguousarray(dr["fluxSolver"].transpose(0, 2, 1), dtype=self.dtype)     # column-major 9 x 15
        self.lib.ssrefv_neigh_dr(ctypes.c_int64(n), _p(q), _p(qa), _p(self.to_ref(Iall)), _p(nb), _p(self._c(Aminus)),
                                 _p(self._c(w)), _p(fr), _p(ft), _p(god), _p(fs),
                                 _p(np.ascontiguousarray(dr["index"], dtype=np.int64)),
                                 _p(np.ascontiguousarray(dr["side"], dtype=np.uint8)),
                                 _p(np.ascontiguousarray(dr["rel"], dtype=np.uint8)))
        return self.from_ref(q), self.ane_from_ref(qa)

    def threads(self):
        return int(self.lib.ssrefv_threads())

    def set_threads(self, n):
        self.lib.ssrefv_set_threads(ctypes.c_int(int(n)))
        return self.threads()

    def proxy_all(self, d, coeffs, passes):
        """Proxy `all` on the host cores; returns (seconds, Q, Qane)."""
        n = d["Q"].shape[0]
        q, qa = self.to_ref(d["Q"]), self.ane_to_ref(d["Qane"])
        I = np.zeros_like(q)
        Ia = np.zeros_like(qa)
        c = self._c(coeffs)
        t = self.lib.ssrefv_proxy_all(
            ctypes.c_int64(n), ctypes.c_int(passes), _p(c), _p(q), _p(qa), _p(I), _p(Ia), _p(self._c(d["star"])),
            _p(self._c(d["Aplus"])), _p(self._c(d["Aminus"])), _p(self._c(d["E"])), _p(self._c(d["w"])), _p(self._c(d["W"])),
            _p(np.ascontiguousarray(d["neighbors"], dtype=np.int64)), _p(np.ascontiguousarray(d["face_rel"], dtype=np.uint8)),
            _p(np.ascontiguousarray(d["face_types"], dtype=np.uint8)))
        return float(t), self.from_ref(q), self.ane_from_ref(qa)
