import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from seissol_b200 import api, proxy
from oracle import aderdg as O
order, dt, n = 4, np.float32, 40
ctx = api.Context(order, 4)
host = O.proxy_data(order, n, dt, seed=5)
tab = O.load_tables(order, dt)
c = O.time_integrate_coeffs(order, 0.0, 1.0, dt)
star = O.star_dense(host["star"], tab["star_pattern"])
I, ders = O.ader(tab, host["Q"], star, c, True)
for want in (False, True):
    data = proxy.ProxyData.from_host(ctx, host)
    ders_dev = torch.zeros((n, ctx.dersize), dtype=data.dtype, device=data.device)
    ctx.ader(api.strided(data.dofs), api.strided(data.buffers), api.strided(ders_dev) if want else None, api.strided(data.local_int), c, n)
    torch.cuda.synchronize()
    got = data.to_logical(data.buffers)
    bad = np.abs(got - I) > 1e-5
    print("wantDer", want, "bad entries", bad.sum(), "elements", np.unique(np.nonzero(bad)[0])[:20], "cols", np.unique(np.nonzero(bad)[2]), "rows", np.unique(np.nonzero(bad)[1]))
    if want:
        dd = ders_dev.cpu().numpy()
        for d in range(order):
            rows = O.nbasis(order - d)
            blk = dd[:, ctx.der_offset[d]:ctx.der_offset[d] + 9 * ctx.der_ld[d]].reshape(n, 9, ctx.der_ld[d])
            g = blk[:, :, :rows].transpose(0, 2, 1); w = ders[d][:, :rows]
            b = np.abs(g - w) > 1e-5
            print(" dQ", d, "bad", b.sum(), "cols", np.unique(np.nonzero(b)[2]), "rows", np.unique(np.nonzero(b)[1]), "elems", np.unique(np.nonzero(b)[0])[:10])
        e = np.nonzero(bad)[0][0] if bad.sum() else 0
        print(" sample got-I col0:", (got - I)[e, :, 0]); print(" c1*dQ1 col0:", c[1]*ders[1][e,:,0]); print(" c2*dQ2 col0:", c[2]*ders[2][e,:,0], c[3]*ders[3][e,:,0])
