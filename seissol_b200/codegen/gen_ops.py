#!/usr/bin/env python3
"""Emit the sparse reference-element operators as straight-line CUDA device code.

B200-first design decision: every constant operator of the ADER-DG
update (kDivM, kDivMT, rDivM, fMrT, rT, fP) is sparse with a pattern known at
build time, so instead of storing matrices in memory we emit one FMA per
non-zero with the matrix value as an *immediate operand*.  A thread owns one
column (one quantity of one element) held in registers; the emitted code is a
sequence `acc[k] = fma(c, x[l], acc[k])`.  This removes all matrix loads and all
index arithmetic from the inner loops; shared memory is touched only for the
element data itself (128-bit vector loads along the basis-function index).

Input : seissol_b200/tables/o{order}.npz  (tools/make_tables.py)
Output: seissol_b200/csrc/gen/ops_o{order}.cuh (git-ignored, rebuilt by build())

Math restated from the reference's tensor expressions (codegen/kernels/aderdg.py:
volume :321-334, localFlux :373-385, neighboringFlux :414-461, derivative
:463-532); kDivMT is negated here as src/Memory/GlobalData.cpp:51-62 does.
"""
import argparse
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TABLES = os.path.join(HERE, "..", "tables")
OUTDIR = os.path.join(HERE, "..", "csrc", "gen")


def nb3(o):
    return o * (o + 1) * (o + 2) // 6


def nb2(o):
    return o * (o + 1) // 2


def pad4(n):
    return (n + 3) // 4 * 4


def lit(v):
    return f"R({float(v)!r})"


class Emitter:
    def __init__(self):
        self.lines = []

    def w(self, s=""):
        self.lines.append(s)

    def text(self):
        return "\n".join(self.lines) + "\n"


LS_PERIOD = 64   # a lock-step marker SSB_LS(i) every LS_PERIOD FMAs (common.cuh decides whether it is a barrier)


def emit_acc(em, touched, name, k, coef, operand):
    """acc[k] (+)= coef*operand; first touch of a zero-initialised output is a plain mul."""
    if getattr(em, "lockstep", False):
        em.nacc = getattr(em, "nacc", 0) + 1
        if em.nacc % LS_PERIOD == 0:
            em.w(f"  SSB_LS({em.nacc // LS_PERIOD});")
    if touched is not None and k not in touched:
        touched.add(k)
        em.w(f"  {name}[{k}] = pk_mulc({lit(coef)}, {operand});")
    else:
        em.w(f"  {name}[{k}] = pk_fmac({lit(coef)}, {operand}, {name}[{k}]);")


def gen_star_sweep(em, fname, mats, rows_in, rows_out, out_pad, accumulate):
    """out[k] (+)= sum_j sum_l M_j[k,l] * t_j[l],  t_j[l] = sum_{c<3} s[j][c] * a_c[l]
    where a_c = column r_c of the source tensor (pointers a0,a1,a2 in shared memory).
    mats: list of 3 dense matrices (already restricted to [rows_out, rows_in])."""
    em.lockstep, em.nacc = True, 0
    em.w("template <typename R>")
    em.w(f"__device__ __forceinline__ void {fname}(typename Pk<R>::T (&out)[{out_pad}], const R* __restrict__ a0,"
         " const R* __restrict__ a1, const R* __restrict__ a2, const typename Pk<R>::T (&s)[3][3]) {")
    em.w("  using P = typename Pk<R>::T; constexpr int V = Pk<R>::V;")
    touched = None if accumulate else set()
    nfma = 0
    for l0 in range(0, pad4(rows_in), 4):
        cols = [l for l in range(l0, min(l0 + 4, rows_in))
                if any(np.any(m[:rows_out, l] != 0) for m in mats)]
        if not cols:
            continue
        em.w("  {")
        em.w(f"  P x0[4], x1[4], x2[4]; ldv4(x0, a0 + {l0} * V); ldv4(x1, a1 + {l0} * V); ldv4(x2, a2 + {l0} * V);")
        for l in cols:
            for j, m in enumerate(mats):
                ks = np.nonzero(m[:rows_out, l])[0]
                if len(ks) == 0:
                    continue
                i = l - l0
                em.w(f"  {{ const P t = pk_fma(s[{j}][2], x2[{i}], pk_fma(s[{j}][1], x1[{i}], pk_mul(s[{j}][0], x0[{i}])));")
                nfma += 3
                for k in ks:
                    emit_acc(em, touched, "out", int(k), m[k, l], "t")
                    nfma += 1
                em.w("  }")
        em.w("  }")
    if touched is not None:
        for k in range(out_pad):
            if k not in touched:
                em.w(f"  out[{k}] = pk_zero<R>();")
    em.w("}")
    em.w()
    return nfma


def gen_spmv_smem(em, fname, m, out_pad, accumulate, lockstep=False):
    """out (+)= M . x, x read from shared memory with 128-bit loads (column of length cols)."""
    rows, cols = m.shape
    em.lockstep, em.nacc = lockstep, 0
    em.w("template <typename R, typename Src>")
    em.w(f"__device__ __forceinline__ void {fname}(typename Pk<R>::T (&out)[{out_pad}], const Src& x) {{")
    em.w("  using P = typename Pk<R>::T;")
    touched = None if accumulate else set()
    nfma = 0
    for l0 in range(0, pad4(cols), 4):
        ls = [l for l in range(l0, min(l0 + 4, cols)) if np.any(m[:, l] != 0)]
        if not ls:
            continue
        em.w(f"  {{ P v[4]; x.load4(v, {l0});")
        for l in ls:
            for k in np.nonzero(m[:, l])[0]:
                emit_acc(em, touched, "out", int(k), m[k, l], f"v[{l - l0}]")
                nfma += 1
        em.w("  }")
    if touched is not None:
        for k in range(out_pad):
            if k not in touched:
                em.w(f"  out[{k}] = pk_zero<R>();")
    em.w("}")
    em.w()
    return nfma


def gen_spmv_reg(em, fname, m, in_pad, out_pad, accumulate):
    """out (+)= M . x with x in registers."""
    em.lockstep = False
    em.w("template <typename R>")
    em.w(f"__device__ __forceinline__ void {fname}(typename Pk<R>::T (&out)[{out_pad}], const typename Pk<R>::T (&x)[{in_pad}]) {{")
    touched = None if accumulate else set()
    nfma = 0
    rows, cols = m.shape
    for l in range(cols):
        for k in np.nonzero(m[:, l])[0]:
            emit_acc(em, touched, "out", int(k), m[k, l], f"x[{l}]")
            nfma += 1
    if touched is not None:
        for k in range(out_pad):
            if k not in touched:
                em.w(f"  out[{k}] = pk_zero<R>();")
    em.w("}")
    em.w()
    return nfma


def gen_spmv_chunks(em, fname, m, in_pad, out_pad):
    """out += M . x where x arrives in chunks of 4 entries from a kernel-supplied source
    (js.template chunk<c>(x4) yields x[4c .. 4c+3]): the full x column never sits in registers."""
    em.lockstep, em.nacc = True, 0
    em.w("template <typename R, typename JS>")
    em.w(f"__device__ __forceinline__ void {fname}(typename Pk<R>::T (&out)[{out_pad}], const JS& js) {{")
    nfma = 0
    rows, cols = m.shape
    for c in range(in_pad // 4):
        ls = [l for l in range(4 * c, min(4 * c + 4, cols)) if np.any(m[:, l] != 0)]
        if not ls:
            continue
        em.w(f"  {{ typename Pk<R>::T x[4]; js.template chunk<{c}>(x);")
        for l in ls:
            for k in np.nonzero(m[:, l])[0]:
                emit_acc(em, None, "out", int(k), m[k, l], f"x[{l - 4 * c}]")
                nfma += 1
        em.w("  }")
    em.w("}")
    em.w()
    return nfma


def gen_dispatch(em, base, count, sig_decl, sig_call, src=False):
    em.w(f"template <typename R, typename {src}>" if src else "template <typename R>")
    em.w(f"__device__ __forceinline__ void {base}(int i, {sig_decl}) {{")
    em.w("  switch (i) {")
    for i in range(count):
        em.w(f"    case {i}: {base}_{i}<R>({sig_call}); break;")
    em.w("    default: break;")
    em.w("  }")
    em.w("}")
    em.w()


def generate(order):
    z = np.load(os.path.join(TABLES, f"o{order}.npz"))
    B, B2 = nb3(order), nb2(order)
    kdm, kdmt = z["kDivM"], -z["kDivMT"]
    em = Emitter()
    em.w(f"// GENERATED by seissol_b200/codegen/gen_ops.py for order {order} -- do not edit.")
    em.w("#pragma once")
    em.w('#include "../common.cuh"')
    em.w(f"namespace ssb {{ namespace gen_o{order} {{")
    em.w(f"constexpr int kOrder = {order};")
    em.w(f"constexpr int kB = {B};        // 3D basis functions")
    em.w(f"constexpr int kB2 = {B2};      // 2D (face) basis functions")
    em.w(f"constexpr int kBP = {pad4(B)};   // padded to a multiple of 4")
    em.w(f"constexpr int kB2P = {pad4(B2)};")
    rows = [nb3(order - d) for d in range(order)]
    em.w("// rows of dQ(d) (SURVEY.md Appendix A: dQ(d) keeps B(O-d) rows)")
    em.w(f"__device__ constexpr int kDerRows[{order}] = {{{', '.join(str(r) for r in rows)}}};")
    stats = {}
    # --- ADER derivative steps d = 1..O-1
    for d in range(1, order):
        rin, rout = nb3(order - d + 1), nb3(order - d)
        mats = [kdmt[j][:rout, :rin] for j in range(3)]
        assert all(np.count_nonzero(kdmt[j][rout:, :rin]) == 0 for j in range(3))
        stats[f"ader_step_{d}"] = gen_star_sweep(em, f"ader_step_{d}", mats, rin, rout, pad4(rout), False)
    # --- volume
    stats["volume"] = gen_star_sweep(em, "volume", [kdm[i] for i in range(3)], B, B, pad4(B), True)
    # --- face operators
    for f in range(4):
        stats[f"fMrT_{f}"] = gen_spmv_smem(em, f"fMrT_{f}", z["fMrT"][f], pad4(B2), False, lockstep=True)
        stats[f"rT_{f}"] = gen_spmv_smem(em, f"rT_{f}", z["rT"][f], pad4(B2), False)
        stats[f"rDivM_{f}"] = gen_spmv_chunks(em, f"rDivM_{f}", z["rDivM"][f], pad4(B2), pad4(B))
    for h in range(3):
        stats[f"fP_{h}"] = gen_spmv_reg(em, f"fP_{h}", z["fP"][h], pad4(B2), pad4(B2), False)
    # --- face-projection path (kernels_faceproj.cuh): the local kernel stores F_f = fMrT_f . I = D . (rT_f . I) per face
    #     (the face mass matrix D is diagonal and the same for every face), the neighbor kernel applies
    #     fPD_h = fP_h . D^-1 and rDivM_i with its input in registers.
    dvec = np.zeros(B2)
    for f in range(4):
        rt, fm = z["rT"][f], z["fMrT"][f]
        assert np.array_equal(rt != 0, fm != 0)
        for k in range(B2):
            nzc = np.nonzero(rt[k])[0]
            ratio = fm[k, nzc] / rt[k, nzc]
            assert np.allclose(ratio, ratio[0], rtol=1e-12, atol=0), "fMrT is not a row scaling of rT"
            assert dvec[k] == 0 or np.isclose(dvec[k], ratio[0], rtol=1e-12), "face mass matrix differs between faces"
            dvec[k] = ratio[0]
    for h in range(3):
        stats[f"fPD_{h}"] = gen_spmv_reg(em, f"fPD_{h}", z["fP"][h] / dvec[None, :], pad4(B2), pad4(B2), False)
    for f in range(4):
        stats[f"rDivMr_{f}"] = gen_spmv_reg(em, f"rDivMr_{f}", z["rDivM"][f], pad4(B2), pad4(B), True)
    PT = "typename Pk<R>::T"
    gen_dispatch(em, "fMrT", 4, f"{PT} (&out)[{pad4(B2)}], const Src& x", "out, x", src="Src")
    gen_dispatch(em, "rT", 4, f"{PT} (&out)[{pad4(B2)}], const Src& x", "out, x", src="Src")
    gen_dispatch(em, "rDivM", 4, f"{PT} (&out)[{pad4(B)}], const JS& js", "out, js", src="JS")
    gen_dispatch(em, "fP", 3, f"{PT} (&out)[{pad4(B2)}], const {PT} (&x)[{pad4(B2)}]", "out, x")
    gen_dispatch(em, "fPD", 3, f"{PT} (&out)[{pad4(B2)}], const {PT} (&x)[{pad4(B2)}]", "out, x")
    gen_dispatch(em, "rDivMr", 4, f"{PT} (&out)[{pad4(B)}], const {PT} (&x)[{pad4(B2)}]", "out, x")
    # --- ADER driver: the kernel supplies a context with src()/finish() (hand-written)
    em.w("template <typename R, typename Ctx>")
    em.w("__device__ __forceinline__ void ader_run(Ctx& c) {")
    for d in range(1, order):
        n = pad4(nb3(order - d))
        em.w(f"  {{ typename Pk<R>::T out[{n}]; ader_step_{d}<R>(out, c.src({d}, 0), c.src({d}, 1), c.src({d}, 2), c.s);"
             f" c.template finish<{n}>({d}, out); }}")
    em.w("}")
    em.w()
    em.w("// FMA instruction counts per thread (one column):")
    for k, v in stats.items():
        em.w(f"//   {k}: {v}")
    em.w("}}  // namespace ssb::gen")
    return em.text(), stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--orders", type=int, nargs="*", default=[4, 5, 6])
    args = ap.parse_args()
    os.makedirs(OUTDIR, exist_ok=True)
    for o in args.orders:
        text, stats = generate(o)
        path = os.path.join(OUTDIR, f"ops_o{o}.cuh")
        old = open(path).read() if os.path.exists(path) else None
        if old != text:
            open(path, "w").write(text)
        print(f"[gen_ops] order {o}: {len(text.splitlines())} lines, {sum(stats.values())} FMA/column")


if __name__ == "__main__":
    main()
