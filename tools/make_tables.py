#!/usr/bin/env python3
"""Convert the reference's constant operator matrices into our own table files.

The ADER-DG reference-element operators (stiffness, face-restriction, flux and
fault-interpolation matrices) are mathematical constants of the Dubiner basis on
the unit tetrahedron.  The reference ships them as sparse XML / JSON entry lists
(`/root/reference/codegen/matrices/matrices_{B}.xml`, `dr_stroud_matrices_{O}.json`,
`mass_{O}.json`; format described in SURVEY.md Appendix A, parsed there by
`submodules/yateto/yateto/input.py:83-140`).  This script reads those lists and
writes ONE compressed ``.npz`` per order under ``seissol_b200/tables/`` holding
dense float64 arrays with our naming (``kDivM`` -> shape (3,B,B) ...).  Those
``.npz`` files are what the oracle, the code generator and the tests consume; the
reference tree is not needed at build or run time.

Run (only in a container that has the reference mounted):
    python tools/make_tables.py [--ref /root/reference] [--orders 2 3 4 5 6 7]
"""
import argparse
import json
import os
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "seissol_b200", "tables")


def nbasis3(order):
    return order * (order + 1) * (order + 2) // 6


def nbasis2(order):
    return order * (order + 1) // 2


def read_xml(path):
    out = {}
    root = ET.parse(path).getroot()
    for node in root:
        name = node.get("name")
        rows, cols = int(node.get("rows")), int(node.get("columns"))
        m = np.zeros((rows, cols))
        for e in node:
            m[int(e.get("row")) - 1, int(e.get("column")) - 1] = float(e.get("value"))
        out[name] = m
    return out


def read_json(path):
    out = {}
    for m in json.load(open(path)):
        rows = m["rows"]
        cols = m.get("columns", 1)
        a = np.zeros((rows, cols))
        for e in m["entries"]:
            a[int(e[0]) - 1, int(e[1]) - 1 if len(e) > 2 else 0] = float(e[-1])
        out[m["name"]] = a
    return out


def family(db, base, n):
    return np.stack([db[f"{base}({i})"] for i in range(n)])


def build(ref, order):
    B = nbasis3(order)
    mdir = os.path.join(ref, "codegen", "matrices")
    db = read_xml(os.path.join(mdir, f"matrices_{B}.xml"))
    dr = read_json(os.path.join(mdir, f"dr_stroud_matrices_{order}.json"))
    mass = read_json(os.path.join(mdir, f"mass_{order}.json"))
    nodal = read_json(os.path.join(mdir, "nodal", f"nodalBoundary_matrices_{order}.json"))
    t = {
        "order": np.int64(order),
        "kDivM": family(db, "kDivM", 3),
        # NOTE: stored as in the matrix file; the solver uses the NEGATED kDivMT
        # (reference: src/Memory/GlobalData.cpp:51-62).  Consumers must negate.
        "kDivMT": family(db, "kDivMT", 3),
        "rDivM": family(db, "rDivM", 4),
        "fMrT": family(db, "fMrT", 4),
        "rT": family(db, "rT", 4),
        "fP": family(db, "fP", 3),
        "M2": mass["M2"],
        "M3": mass["M3"],
        "quadpoints": dr["quadpoints"],
        "quadweights": dr["quadweights"][:, 0],
        "resample": dr["resample"],
        # flat index side + 4*rel, as the generated family index
        "V3mTo2n": np.stack([dr[f"V3mTo2n({s},{r})"] for r in range(4) for s in range(4)]),
        "V3mTo2nTWDivM": np.stack(
            [dr[f"V3mTo2nTWDivM({s},{r})"] for r in range(4) for s in range(4)]
        ),
        # nodal boundary conditions (codegen/kernels/nodalbc.py, aderdg.py:123-131): modal volume -> face nodes, face
        # nodes -> modal face; project2nFaceTo3m(i) = rDivM(i) . V2nTo2m is formed by the consumers
        "V3mTo2nFace": family(nodal, "V3mTo2nFace", 4),
        "V2nTo2m": nodal["V2nTo2m"],
        "nodes2D": nodal["nodes2D"],
    }
    return t


def star_pattern(ref):
    root = ET.parse(os.path.join(ref, "codegen", "matrices", "star.xml")).getroot()
    ent = []
    for node in root:
        for e in node:
            ent.append((int(e.get("row")) - 1, int(e.get("column")) - 1))
    # CSC order (column-major over the pattern) = order of the 24 stored values
    ent.sort(key=lambda rc: (rc[1], rc[0]))
    return np.array(ent, dtype=np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--orders", type=int, nargs="*", default=[2, 3, 4, 5, 6, 7])
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    sp = star_pattern(args.ref)
    for o in args.orders:
        t = build(args.ref, o)
        t["star_pattern"] = sp
        path = os.path.join(OUT, f"o{o}.npz")
        np.savez_compressed(path, **t)
        nnz = {k: int(np.count_nonzero(v)) for k, v in t.items() if v.ndim >= 2}
        print(f"order {o}: {path} ({os.path.getsize(path)} B) nnz={nnz}")


if __name__ == "__main__":
    main()
