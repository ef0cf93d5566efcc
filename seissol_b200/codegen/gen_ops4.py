#!/usr/bin/env python3
"""Operator code of kernel generation 4: every operator split into two ROW HALVES.

Same idea as gen_ops.py (one FMA per non-zero, matrix value as an immediate operand),
but a thread now owns only HALF of a quantity column: rows k with (k >> 1) & 1 == H.  Two threads
(H = 0, 1) share a column, each keeps 28 instead of 56 accumulators (order 6), so that

  * single precision runs on packed FFMA2 (two elements per thread, `Pk<float>::T = float2`) inside the
    128-register budget of 16 resident warps per SM -- half the instruction bytes per flop, which is what
    the instruction-fetch-bound straight-line code needs (tools/ubench);
  * double precision fits the same budget (28 doubles = 56 registers) instead of spilling at 255.

Row k of a half lives in local slot loc(k) = 2 (k >> 2) + (k & 1); a 128-bit shared-memory access covers the
two rows (4p + 2H, 4p + 2H + 1) of "piece" p for both packed elements (float) or one element (double).

Input : seissol_b200/tables/o{order}.npz     Output: seissol_b200/csrc/gen/ops4_o{order}.cuh
Math restated from the reference's tensor expressions (codegen/kernels/aderdg.py: volume :321-334,
localFlux :373-385, neighboringFlux :414-461, derivative :463-532); kDivMT negated as
src/Memory/GlobalData.cpp:51-62 does.
"""
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


def owner(k):
    return (k >> 1) & 1


def loc(k):
    return 2 * (k >> 2) + (k & 1)


def lit(v):
    return f"R({float(v)!r})"


class Emitter:
    def __init__(self):
        self.lines = []
        self.consts = None     # constants of the function being emitted, in order of use (constant-table variant)
        self.tab = None
        self.tables = []       # (name, values) of all functions

    def begin_table(self, name):
        self.consts, self.tab = [], name

    def end_table(self):
        vals = self.consts + [0.0] * (len(self.consts) % 2)
        self.tables.append((self.tab, vals))
        n = len(vals) // 2
        self.consts, self.tab = None, None
        return n

    def w(self, s=""):
        self.lines.append(s)

    def text(self):
        return "\n".join(self.lines) + "\n"


def emit_acc(em, touched, name, i, coef, operand):
    """One FMA per non-zero.  The matrix value is a literal (single precision: FFMA with an immediate operand) or, with
    SSB_KTAB = 1 in double precision, half of a 16-byte pair fetched from the function's constant table in order of use
    (DFMA has no immediate form: a literal costs two 32-bit moves per FMA, the table one 128-bit load per two FMAs)."""
    j = len(em.consts)
    em.consts.append(float(coef))
    if j % 2 == 0:
        em.w(f"    kp[{j // 2}] = kload<R>({em.tab}, {j // 2});")
    c = f"ksel<R, {j % 2}>(kp[{j // 2}], {lit(coef)})"
    if touched is not None and i not in touched:
        touched.add(i)
        em.w(f"    {name}[{i}] = pk_mulc({c}, {operand});")
    else:
        em.w(f"    {name}[{i}] = pk_fmac({c}, {operand}, {name}[{i}]);")


def finish_untouched(em, touched, n_local):
    if touched is not None:
        for i in range(n_local):
            if i not in touched:
                em.w(f"    out[{i}] = pk_zero<R>();")


def rows_of(h, n):
    return [k for k in range(n) if owner(k) == h]


def gen_star_sweep(em, fname, mats, rows_in, rows_out, accumulate):
    """out[loc(k)] (+)= sum_j sum_l M_j[k,l] t_j[l] for the rows k of half H,
    t_j[l] = sum_c s[j][c] a_c[l]; a_c = interleaved source columns in shared memory."""
    n_local = pad4(rows_out) // 2
    em.w("template <typename R, int H>")
    em.w(f"__device__ __forceinline__ void {fname}(typename Pk<R>::T (&out)[{n_local}], const R* __restrict__ a0,"
         " const R* __restrict__ a1, const R* __restrict__ a2, const typename Pk<R>::T (&s)[3][3]) {")
    em.w("  using P = typename Pk<R>::T; constexpr int V = Pk<R>::V;")
    counts = []
    for h in (0, 1):
        em.w(f"  if constexpr (H == {h}) {{")
        em.begin_table(f"ktab_{fname}_h{h}")
        mark = len(em.lines)
        em.w("KP_PLACEHOLDER")
        mine = rows_of(h, rows_out)
        touched = None if accumulate else set()
        nfma = 0
        for l0 in range(0, pad4(rows_in), 2):
            cols = [l for l in range(l0, min(l0 + 2, rows_in)) if any(np.any(m[mine, l] != 0) for m in mats)]
            if not cols:
                continue
            em.w("    {")
            em.w(f"    P x0[2], x1[2], x2[2]; ldv2(x0, a0 + {l0} * V); ldv2(x1, a1 + {l0} * V); ldv2(x2, a2 + {l0} * V);")
            for l in cols:
                for j, m in enumerate(mats):
                    ks = [k for k in mine if m[k, l] != 0]
                    if not ks:
                        continue
                    i = l - l0
                    em.w(f"    {{ const P t = pk_fma(s[{j}][2], x2[{i}], pk_fma(s[{j}][1], x1[{i}], pk_mul(s[{j}][0], x0[{i}])));")
                    nfma += 3
                    for k in ks:
                        emit_acc(em, touched, "out", loc(k), m[k, l], "t")
                        nfma += 1
                    em.w("    }")
            em.w("    }")
        finish_untouched(em, touched, n_local)
        em.w("  }")
        em.lines[mark] = f"    KPair<R> kp[{max(1, em.end_table())}];"
        counts.append(nfma)
    em.w("}")
    em.w()
    return counts


def gen_spmv_src(em, fname, m, accumulate=False):
    """out[loc(k)] (+)= sum_l M[k,l] x[l] for the rows of half H; x = full column read through Src."""
    rows, cols = m.shape
    n_local = pad4(rows) // 2
    em.w("template <typename R, int H, typename Src>")
    em.w(f"__device__ __forceinline__ void {fname}(typename Pk<R>::T (&out)[{n_local}], const Src& x) {{")
    em.w("  using P = typename Pk<R>::T;")
    counts = []
    for h in (0, 1):
        em.w(f"  if constexpr (H == {h}) {{")
        em.begin_table(f"ktab_{fname}_h{h}")
        mark = len(em.lines)
        em.w("KP_PLACEHOLDER")
        mine = rows_of(h, rows)
        touched = None if accumulate else set()
        nfma = 0
        for l0 in range(0, pad4(cols), 4):
            need = [l for l in range(l0, min(l0 + 4, cols)) if np.any(m[mine, l] != 0)]
            if not need:
                continue
            lo, hi = any(l < l0 + 2 for l in need), any(l >= l0 + 2 for l in need)
            if lo and hi:
                em.w(f"    {{ P v[4]; x.load4(v, {l0});")
                base = l0
            else:
                base = l0 if lo else l0 + 2
                em.w(f"    {{ P v[2]; x.load2(v, {base});")
            for l in need:
                for k in mine:
                    if m[k, l] != 0:
                        emit_acc(em, touched, "out", loc(k), m[k, l], f"v[{l - base}]")
                        nfma += 1
            em.w("    }")
        finish_untouched(em, touched, n_local)
        em.w("  }")
        em.lines[mark] = f"    KPair<R> kp[{max(1, em.end_table())}];"
        counts.append(nfma)
    em.w("}")
    em.w()
    return counts


def gen_spmv_chunks(em, fname, m):
    """out[loc(k)] += sum_m M[k,m] J[m] for the rows of half H; J arrives in chunks of 4 rows from a
    kernel-supplied source (js.template chunk<c>(x) fills x[4] = J[4c .. 4c+3])."""
    rows, cols = m.shape
    n_local = pad4(rows) // 2
    em.w("template <typename R, int H, typename JS>")
    em.w(f"__device__ __forceinline__ void {fname}(typename Pk<R>::T (&out)[{n_local}], const JS& js) {{")
    em.w("  using P = typename Pk<R>::T;")
    counts = []
    for h in (0, 1):
        em.w(f"  if constexpr (H == {h}) {{")
        em.begin_table(f"ktab_{fname}_h{h}")
        mark = len(em.lines)
        em.w("KP_PLACEHOLDER")
        mine = rows_of(h, rows)
        nfma = 0
        for c in range(pad4(cols) // 4):
            need = [l for l in range(4 * c, min(4 * c + 4, cols)) if np.any(m[mine, l] != 0)]
            if not need:
                continue
            em.w(f"    {{ P x[4]; js.template chunk<{c}>(x);")
            for l in need:
                for k in mine:
                    if m[k, l] != 0:
                        emit_acc(em, None, "out", loc(k), m[k, l], f"x[{l - 4 * c}]")
                        nfma += 1
            em.w("    }")
        em.w("  }")
        em.lines[mark] = f"    KPair<R> kp[{max(1, em.end_table())}];"
        counts.append(nfma)
    em.w("}")
    em.w()
    return counts


def gen_dispatch(em, base, count, n_local, arg_decl, arg_call, extra_tpl):
    em.w(f"template <typename R, int H, typename {extra_tpl}>")
    em.w(f"__device__ __forceinline__ void {base}(int i, typename Pk<R>::T (&out)[{n_local}], {arg_decl}) {{")
    em.w("  switch (i) {")
    for i in range(count):
        em.w(f"    case {i}: {base}_{i}<R, H>(out, {arg_call}); break;")
    em.w("    default: break;")
    em.w("  }")
    em.w("}")
    em.w()


def generate(order):
    z = np.load(os.path.join(TABLES, f"o{order}.npz"))
    B, B2 = nb3(order), nb2(order)
    kdm, kdmt = z["kDivM"], -z["kDivMT"]
    em = Emitter()
    em.w(f"// GENERATED by seissol_b200/codegen/gen_ops4.py for order {order} -- do not edit.")
    em.w("#pragma once")
    em.w('#include "../common.cuh"')
    em.w(f"namespace ssb {{ namespace gen4_o{order} {{")
    em.w(f"constexpr int kOrder = {order};")
    em.w(f"constexpr int kB = {B};")
    em.w(f"constexpr int kB2 = {B2};")
    em.w(f"constexpr int kBP = {pad4(B)};")
    em.w(f"constexpr int kB2P = {pad4(B2)};")
    stats = {}
    head_end = len(em.lines)
    for d in range(1, order):
        rin, rout = nb3(order - d + 1), nb3(order - d)
        mats = [kdmt[j][:rout, :rin] for j in range(3)]
        assert all(np.count_nonzero(kdmt[j][rout:, :rin]) == 0 for j in range(3))
        stats[f"ader_step_{d}"] = gen_star_sweep(em, f"ader_step_{d}", mats, rin, rout, False)
    stats["volume"] = gen_star_sweep(em, "volume", [kdm[i] for i in range(3)], B, B, True)
    for f in range(4):
        stats[f"fMrT_{f}"] = gen_spmv_src(em, f"fMrT_{f}", z["fMrT"][f])
        stats[f"rT_{f}"] = gen_spmv_src(em, f"rT_{f}", z["rT"][f])
        stats[f"rDivM_{f}"] = gen_spmv_chunks(em, f"rDivM_{f}", z["rDivM"][f])
    for h in range(3):
        stats[f"fP_{h}"] = gen_spmv_src(em, f"fP_{h}", z["fP"][h])
    nl2, nl3 = pad4(B2) // 2, pad4(B) // 2
    gen_dispatch(em, "fMrT", 4, nl2, "const X& x", "x", "X")
    gen_dispatch(em, "rT", 4, nl2, "const X& x", "x", "X")
    gen_dispatch(em, "fP", 3, nl2, "const X& x", "x", "X")
    gen_dispatch(em, "rDivM", 4, nl3, "const X& x", "x", "X")
    em.w("template <typename R, int H, typename Ctx>")
    em.w("__device__ __forceinline__ void ader_run(Ctx& c) {")
    for d in range(1, order):
        n = pad4(nb3(order - d))
        em.w(f"  {{ typename Pk<R>::T out[{n // 2}]; ader_step_{d}<R, H>(out, c.src({d}, 0), c.src({d}, 1), c.src({d}, 2), c.s);"
             f" c.template finish<{n}>({d}, out); }}")
    em.w("}")
    em.w()
    # constant tables of the functions above (used by the SSB_KTAB = 1 build only; unused tables are dropped)
    tabs = ["// matrix values in order of use, one table per function and row half (constant-table variant, SSB_KTAB = 1)"]
    for name, vals in em.tables:
        body = ", ".join(repr(v) for v in vals) if vals else "0.0, 0.0"
        tabs.append(f"static __device__ __align__(16) const double {name}[{max(2, len(vals))}] = {{{body}}};")
    em.lines[head_end:head_end] = tabs
    em.w("// FMA instruction counts per thread (half column), H = 0 / 1:")
    for k, v in stats.items():
        em.w(f"//   {k}: {v[0]} / {v[1]}")
    em.w("}}  // namespace ssb::gen4")
    return em.text(), stats


def emit_all(orders=(4, 5, 6)):
    os.makedirs(OUTDIR, exist_ok=True)
    for o in orders:
        text, stats = generate(o)
        path = os.path.join(OUTDIR, f"ops4_o{o}.cuh")
        if not os.path.exists(path) or open(path).read() != text:
            open(path, "w").write(text)


if __name__ == "__main__":
    emit_all()
    for o in (4, 5, 6):
        _, st = generate(o)
        print(f"[gen_ops4] order {o}: " + ", ".join(f"{k}={v[0]}/{v[1]}" for k, v in st.items()))
