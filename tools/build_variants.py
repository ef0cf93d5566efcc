#!/usr/bin/env python3
"""Build A/B variants of the library (development aid): tools/build_variants.py name=flag,flag[:source,source] ...
Each variant becomes seissol_b200/lib/libseissol_b200_<name>.so (select with SSB_LIB=<path>)."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from seissol_b200 import build as b  # noqa: E402


def one(spec):
    name, rest = spec.split("=", 1)
    flags, _, srcs = rest.partition(":")
    srcs = srcs.split(",") if srcs else ["inst_o6_f32.cu"]
    return b.build(verbose=False, extra_flags=tuple("-D" + f for f in flags.split(",") if f), variant=name,
                   variant_sources=srcs)


if __name__ == "__main__":
    with ThreadPoolExecutor(max_workers=6) as ex:
        for lib in ex.map(one, sys.argv[1:]):
            print("built", lib, flush=True)
