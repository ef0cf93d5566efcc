#!/usr/bin/env python3
"""Per-kernel SASS summary of libseissol_b200.so (cuobjdump -sass): instruction count, code bytes (16 B per
instruction) and the counts of the mnemonics that identify the execution resources (FFMA/FFMA2/DFMA: CUDA-core FMAs,
UBLKCP: 1-D bulk copies (TMA unit), UTC*MMA: tcgen05.mma, LDTM/STTM: tensor-memory loads/stores, UMOV: constant
materialisation of the double-precision code).  Usage: tools/sass_summary.py [lib] > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "seissol_b200", "lib", "libseissol_b200.so")
KEYS = ["FFMA", "FFMA2", "DFMA", "LDS", "STS", "LDG", "STG", "UBLKCP", "UBLKRED", "UTCMMA", "LDTM", "STTM", "UMOV", "BAR", "SYNCS"]
proc = subprocess.Popen(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True)
name, rows, cur = None, [], None
ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
for line in proc.stdout:
    if "Function :" in line:
        if name is not None:
            rows.append((name, cur))
        name, cur = line.split("Function :")[1].strip(), collections.Counter()
        continue
    m = ins.match(line)
    if m and name is not None:
        op = m.group(1)
        cur["n"] += 1
        if op.startswith("UTC") and "MMA" in op:
            cur["UTCMMA"] += 1
        elif op in ("UBLKCP", "UBLKRED", "UBLKPF"):
            cur["UBLKRED" if op == "UBLKRED" else "UBLKCP"] += 1
        elif op in KEYS:
            cur[op] += 1
if name is not None:
    rows.append((name, cur))
dem = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.basename(lib)}: {len(rows)} kernels; instructions are 16 bytes; the measured instruction cache is 32 KB")
print(f"{'instr':>8} {'KB':>7} " + " ".join(f"{k:>7}" for k in KEYS) + "  kernel")
for (n, c), d in sorted(zip(rows, dem), key=lambda t: -t[0][1]["n"]):
    short = re.sub(r"\(anonymous namespace\)::|ssb::|\(ssb::\w+\)$", "", d)
    short = re.sub(r"^void ", "", short)
    print(f"{c['n']:8d} {c['n'] * 16 / 1024:7.1f} " + " ".join(f"{c[k]:7d}" for k in KEYS) + "  " + short[:110])
