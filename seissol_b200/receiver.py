"""Receivers: point samples of the wave field at a fixed sampling interval (host logic, numpy).

Restated from src/Kernels/Receiver.cpp:45-230: a receiver lives in one element; every time step the element's ADER
derivatives dQ(d) (the predictor's Taylor expansion) are evaluated at the sample times inside the step,
Q(x_r, t) = sum_k phi_k(xi_r) sum_d (t - t0)^d / d! dQ(d)[k], with the basis functions sampled once at the receiver's
reference coordinates (SampledBasisFunctions, src/Numerical/BasisFunction.h:120-143).  The derivatives of the few
receiver elements come from the predictor run on a pointer-table batch of just those elements (`ssb_ader` with a
derivative array) or, in the CPU tests, from the oracle."""
import math

import numpy as np

from . import basis as _basis


def locate(coords, points):
    """Element index and reference coordinates of every point: coords [n, 4, 3]; returns (elem [m], xi [m, 3]).
    tetrahedronGlobalToReference (src/Numerical/Transformation.cpp:43-63) on every element, first containing one wins."""
    inv = np.linalg.inv(np.transpose(coords[:, 1:] - coords[:, :1], (0, 2, 1)))     # [n, 3, 3]
    elem, xis = [], []
    for p in np.atleast_2d(points):
        xi = np.einsum("nij,nj->ni", inv, p[None, :] - coords[:, 0])
        inside = np.all(xi >= -1e-12, axis=1) & (xi.sum(axis=1) <= 1 + 1e-12)
        idx = np.nonzero(inside)[0]
        if len(idx) == 0:
            raise ValueError(f"receiver {p} is outside the mesh")
        elem.append(int(idx[0]))
        xis.append(xi[idx[0]])
    return np.array(elem), np.array(xis)


class ReceiverCluster:
    """ReceiverCluster (src/Kernels/Receiver.cpp:63-230) for one time cluster."""

    def __init__(self, order, coords, points, sampling_interval):
        self.order, self.dt_sample = order, float(sampling_interval)
        self.elements, self.xi = locate(coords, points)
        self.phi = _basis.dubiner(order, self.xi[:, 0], self.xi[:, 1], self.xi[:, 2])     # [m, B]
        self.next_time = 0.0
        self.times = []
        self.samples = []          # list of [m, 9]

    def step(self, ders, t0, dt):
        """ders: list over d of [m, B, 9] derivatives of the receiver elements at expansion point t0.  Appends the samples
        with t0 <= t < t0 + dt (the reference's loop :186-217)."""
        while self.next_time < t0 + dt - 1e-12 * dt:
            tau = self.next_time - t0
            q = sum((tau ** d / math.factorial(d)) * np.einsum("mk,mkq->mq", self.phi, ders[d]) for d in range(self.order))
            self.times.append(self.next_time)
            self.samples.append(q)
            self.next_time += self.dt_sample

    def series(self):
        return np.array(self.times), np.stack(self.samples, 1)      # [T], [m, T, 9]


def receiver_derivatives_cuda(torch, api, ctx, data, elements, coeffs):
    """Derivative families of the receiver elements through the C ABI: the predictor on a pointer-table batch of just those
    elements (their time integral goes to a scratch array).  Returns a list over d of [m, B, 9] numpy arrays."""
    dev = data.dofs.device
    m = len(elements)
    idx = torch.from_numpy(np.asarray(elements, dtype=np.int64)).to(dev)
    es = data.dofs.element_size()
    dptr = (data.dofs.data_ptr() + idx * (data.dofs.stride(0) * es)).contiguous()
    lptr = (data.local_int.data_ptr() + idx * (data.local_int.stride(0) * es)).contiguous()
    scratch = torch.empty((m, data.dofs.shape[1]), dtype=data.dofs.dtype, device=dev)
    derivs = torch.zeros((m, ctx.dersize), dtype=data.dofs.dtype, device=dev)
    ctx.ader(api.pointer_table(dptr), api.strided(scratch), api.strided(derivs), api.pointer_table(lptr), coeffs, m)
    host = derivs.cpu().numpy()
    out = []
    for d in range(ctx.order):
        rows = (ctx.order - d) * (ctx.order - d + 1) * (ctx.order - d + 2) // 6
        blk = host[:, ctx.der_offset[d]:ctx.der_offset[d] + 9 * ctx.der_ld[d]].reshape(m, 9, ctx.der_ld[d])
        full = np.zeros((m, ctx.B, 9), dtype=host.dtype)
        full[:, :rows] = blk[:, :, :rows].transpose(0, 2, 1)
        out.append(full)
    return out
