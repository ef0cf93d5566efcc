#!/usr/bin/env python3
"""Run the C++ LTS host (seissol_b200/lib/ssb_lts_proxy) on N GPUs of this box: export the rank layouts of one cube mesh
(seissol_b200/partition.py), start one process per GPU, print their JSON lines and the aggregate rate."""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from seissol_b200 import partition as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--dims", type=int, nargs=3, default=[100, 70, 72])   # cubes per GPU along x, y, z (x is multiplied by --gpus)
    ap.add_argument("--clusters", type=int, default=3)
    ap.add_argument("--order", type=int, default=6)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--faceproj", type=int, default=1)
    a = ap.parse_args()
    nx, ny, nz = a.dims[0] * a.gpus, a.dims[1], a.dims[2]
    tmp = tempfile.mkdtemp(prefix="ssb_lts_")
    prefix = os.path.join(tmp, "layout")
    for r in range(a.gpus):
        P.export_layout(P.build_layout(nx, ny, nz, r, a.gpus, a.clusters), f"{prefix}_{r}.bin")
    exe = os.path.join(ROOT, "seissol_b200", "lib", "ssb_lts_proxy")
    procs = []
    for r in range(a.gpus):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(a.gpus), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([exe, "--layout", prefix, "--order", str(a.order), "--steps", str(a.steps), "--warmup",
                                       str(a.warmup), "--faceproj", str(a.faceproj), "--idfile", os.path.join(tmp, "id")],
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    total, tmax = 0.0, 0.0
    for p in procs:
        so, se = p.communicate()
        if p.returncode != 0:
            print(se, file=sys.stderr)
            sys.exit(1)
        line = json.loads(so.strip().splitlines()[-1])
        print(json.dumps(line))
        total += line["elem-updates-per-s"] * line["time"]
        tmax = max(tmax, line["time"])
    print(json.dumps({"host": "C++ (ssb_lts_proxy)", "gpus": a.gpus, "mesh": [nx, ny, nz], "clusters": a.clusters,
                      "element-updates-per-s": total / tmax if tmax > 0 else 0.0, "seconds": tmax}))


if __name__ == "__main__":
    main()
