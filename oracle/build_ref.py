#!/usr/bin/env python3
"""Build oracle/_ref: the REFERENCE's own CPU kernels for the hot path.

TEST INFRASTRUCTURE ONLY.  What this does (only in a container where the
reference is mounted at /root/reference; the GPU box uses the prebuilt .so):

 1. runs the reference's kernel generator `codegen/generate.py` with the
    generic C++ loop back-end (`--gemm_tools none --device_backend none`), the
    same invocation the reference's own pytest uses
    (tests/python/codegen/test_generate_goldenfile.py:37-76), writing into
    oracle/_ref/gen/<cfg>/ (git-ignored; no reference source is copied into
    the repository history),
 2. compiles the generated kernel.cpp/init.cpp/tensor.cpp together with
    oracle/ref_harness.cpp (our thin restatement of the pointer-binding
    wrappers) into oracle/_ref/libssref_<cfg>.so.

LIBXSMM / PSpaMM are not installed (no network), so the GEMMs are the
generator's plain-loop back-end: identical tensor expressions, constants and
layouts; only the micro-kernel summation order differs from a production CPU
build.  `-march=x86-64-v3` (AVX2+FMA) is used instead of `-march=native` so the
library also runs on the GPU host's CPU.  `-fno-gnu-unique -Wl,-Bsymbolic`
keep the per-configuration `seissol::tensor::*` constants private to each .so so
several configurations can be loaded into one Python process.
"""
import argparse
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")

# name -> (equations, order, precision, mechanisms)
CONFIGS = {
    "elastic_o4_f32": ("elastic", 4, "s", 0),
    "elastic_o5_f32": ("elastic", 5, "s", 0),
    "elastic_o6_f32": ("elastic", 6, "s", 0),
    "elastic_o6_f64": ("elastic", 6, "d", 0),
    "elastic_o4_f64": ("elastic", 4, "d", 0),
    "visco_o5_f32": ("viscoelastic2", 5, "s", 3),
    "visco_o5_f64": ("viscoelastic2", 5, "d", 3),
}


def libpath(cfg):
    return os.path.join(REFDIR, f"libssref_{cfg}.so")


def generate(ref, cfg):
    eq, order, prec, mech = CONFIGS[cfg]
    out = os.path.join(REFDIR, "gen", cfg)
    sub = os.path.join(out, f"equation-{eq}-{order}-{'single' if prec == 's' else 'double'}")
    if os.path.exists(os.path.join(sub, "kernel.cpp")):
        return sub
    os.makedirs(out, exist_ok=True)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(ref, "submodules", "yateto") + ":."
    cmd = [sys.executable, "generate.py", "--equations", eq, "--matricesDir", "matrices",
           "--outputDir", out, "--host_arch", "skx", "--device_backend", "none",
           "--order", str(order), "--precision", prec, "--numberOfMechanisms", str(mech),
           "--memLayout", "auto", "--multipleSimulations", "1", "--PlasticityMethod", "nb",
           "--gemm_tools", "none", "--device_codegen", "none", "--drQuadRule", "stroud"]
    t0 = time.time()
    subprocess.run(cmd, cwd=os.path.join(ref, "codegen"), env=env, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    print(f"[oracle/_ref] generated {cfg} in {time.time() - t0:.1f}s", flush=True)
    return sub


def compile_cfg(ref, cfg, force=False):
    eq, order, prec, mech = CONFIGS[cfg]
    lib = libpath(cfg)
    harness = os.path.join(HERE, "ref_harness_visco.cpp" if eq == "viscoelastic2" else "ref_harness.cpp")
    if (not force and os.path.exists(lib)
            and os.path.getmtime(lib) > max(os.path.getmtime(harness),
                                            os.path.getmtime(os.path.join(HERE, "ref_rs.cpp")))):
        return lib
    sub = generate(ref, cfg)
    objs = []
    flags = ["-std=c++17", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-w", "-fno-gnu-unique", "-fvisibility-inlines-hidden",
             f"-I{sub}", f"-I{os.path.join(ref, 'submodules', 'yateto', 'include')}",
             f"-DSSREF_ORDER={order}"] + (["-DSSREF_DOUBLE"] if prec == "d" else []) \
        + ([f"-DSSREF_MECHANISMS={mech}"] if mech else [])
    procs = []
    t0 = time.time()
    rs_flags = [f"-I{os.path.join(HERE, 'stub')}", f"-I{os.path.join(ref, 'src')}"]
    sources = [os.path.join(sub, "kernel.cpp"), os.path.join(sub, "init.cpp"), os.path.join(sub, "tensor.cpp"), harness]
    if eq != "viscoelastic2":
        sources.append(os.path.join(HERE, "ref_rs.cpp"))
    for src in sources:
        obj = os.path.join(REFDIR, "gen", cfg, os.path.basename(src) + ".o")
        objs.append(obj)
        extra = rs_flags if src.endswith("ref_rs.cpp") else []
        procs.append(subprocess.Popen(["g++", *flags, *extra, "-c", src, "-o", obj]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"compile failed for {cfg}")
    subprocess.run(["g++", "-shared", "-fopenmp", "-Wl,-Bsymbolic", "-o", lib, *objs], check=True)
    print(f"[oracle/_ref] built {lib} in {time.time() - t0:.1f}s", flush=True)
    return lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("configs", nargs="*", default=None)
    args = ap.parse_args()
    if not os.path.isdir(args.ref):
        print("[oracle/_ref] reference tree not present; using prebuilt libraries only")
        return 0
    for cfg in (args.configs or ["elastic_o4_f32", "elastic_o6_f32", "elastic_o6_f64"]):
        compile_cfg(args.ref, cfg, args.force)
    return 0


if __name__ == "__main__":
    sys.exit(main())
