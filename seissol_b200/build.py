"""Build libseissol_b200.so in-tree (seissol_b200/lib/), sm_100a only.

Steps: (1) regenerate the immediate-operand operator code (codegen/gen_ops.py),
(2) compile every translation unit under csrc/ with
`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo` in parallel,
(3) link the shared library.  Objects are rebuilt only when a source is newer.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "..", "build", "obj")
LIB = os.path.join(LIBDIR, "libseissol_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


_INCLUDE = __import__("re").compile(r'^\s*#\s*include\s+"([^"]+)"', __import__("re").M)


def _includes(path, seen=None):
    """Files a translation unit depends on: its quoted includes, recursively (so that a change to one kernel family
    recompiles only the units that use it)."""
    seen = set() if seen is None else seen
    try:
        text = open(path).read()
    except OSError:
        return []
    for inc in _INCLUDE.findall(text):
        dep = os.path.normpath(os.path.join(os.path.dirname(path), inc))
        if dep not in seen and os.path.exists(dep):
            seen.add(dep)
            _includes(dep, seen)
    return sorted(seen)


def build(verbose=True, force=False, extra_flags=(), variant=None, variant_sources=None):
    """variant: build a differently-configured library libseissol_b200_<variant>.so (development aid for
    A/B measurements; selected at run time with the environment variable SSB_LIB).  Only the translation
    units named in `variant_sources` are recompiled with `extra_flags`; the rest is linked from the main build."""
    sys.path.insert(0, os.path.join(HERE, "codegen"))
    import gen_ops  # noqa: E402

    os.makedirs(gen_ops.OUTDIR, exist_ok=True)
    for o in (4, 5, 6):
        text, _ = gen_ops.generate(o)
        path = os.path.join(gen_ops.OUTDIR, f"ops_o{o}.cuh")
        if not os.path.exists(path) or open(path).read() != text:
            open(path, "w").write(text)
    extra = sorted(glob.glob(os.path.join(HERE, "codegen", "gen_*.py")))
    for mod in extra:
        name = os.path.splitext(os.path.basename(mod))[0]
        if name != "gen_ops":
            m = __import__(name)
            if hasattr(m, "emit_all"):
                m.emit_all()

    os.makedirs(OBJDIR, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(CSRC, "gen", "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    # generation-4 kernels are the product path in double precision only; the single-precision instantiations
    # (slower than generation 3) are built on request for A/B measurements
    v4_f32 = os.environ.get("SSB_BUILD_V4_F32") == "1"
    if not v4_f32:
        sources = [s for s in sources if not (os.path.basename(s).startswith("inst4_") and s.endswith("_f32.cu"))]
    extra_flags = tuple(extra_flags) + (("-DSSB_HAVE_V4_F32",) if v4_f32 else ())
    jobs = []
    objdir = OBJDIR if variant is None else OBJDIR + "_" + variant
    os.makedirs(objdir, exist_ok=True)
    lib = LIB if variant is None else LIB.replace(".so", f"_{variant}.so")

    def obj_of(src):
        base = os.path.basename(src)
        if variant is not None and (variant_sources is None or base in variant_sources):
            return os.path.join(objdir, base[:-3] + ".o")
        return os.path.join(OBJDIR, base[:-3] + ".o")

    # a library that is newer than every source it is built from is up to date even when the object files are not
    # there (build/ does not travel to the GPU box; the in-tree .so does): nothing to compile, nothing to link
    all_deps = sources + headers
    if variant is None and not force and os.path.exists(lib) and not _newer(lib, all_deps):
        build_host(lib, verbose)
        build_lts_host(lib, verbose)
        return lib
    for src in sources:
        obj = obj_of(src)
        if force or _newer(obj, [src] + _includes(src)):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        flags = extra_flags if obj.startswith(objdir + os.sep) or variant is None else ()
        if v4_f32 and "-DSSB_HAVE_V4_F32" not in flags:
            flags = tuple(flags) + ("-DSSB_HAVE_V4_F32",)
        cmd = ["nvcc", *NVCC_FLAGS, *flags, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[build] compiled {os.path.basename(src)}", flush=True)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    objs = [obj_of(s) for s in sources]
    if jobs or not os.path.exists(lib):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[build] linked {lib}", flush=True)
    if variant is None:
        build_host(lib, verbose)
        build_lts_host(lib, verbose)
    return lib


def build_host(lib, verbose=True):
    """C++ host above the C ABI: host/ssb_host.hpp (the reference's batched kernel classes) + the proxy executable."""
    root = os.path.join(HERE, "..")
    src = os.path.join(root, "host", "proxy_main.cpp")
    exe = os.path.join(LIBDIR, "ssb_proxy")
    deps = [src, os.path.join(root, "host", "ssb_host.hpp"), os.path.join(root, "include", "seissol_b200.h"), lib]
    if not _newer(exe, deps):
        return exe
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-std=c++17", "-O2", src, f"-I{os.path.join(root, 'include')}", f"-I{os.path.join(root, 'host')}",
           f"-I{os.path.join(cuda, 'include')}", f"-L{LIBDIR}", "-lseissol_b200", f"-L{os.path.join(cuda, 'lib64')}",
           "-lcudart", "-Wl,-rpath,$ORIGIN", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"host build failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[build] built {exe}", flush=True)
    return exe


def build_lts_host(lib, verbose=True):
    """C++ LTS host (host/ssb_lts.hpp: plan, rate-2 schedule, NCCL halo) + the ssb_lts_proxy executable.  NCCL comes
    from the system (libnccl-dev) or from the wheel torch brings; without either the single-rank build remains."""
    root = os.path.join(HERE, "..")
    src = os.path.join(root, "host", "lts_main.cpp")
    exe = os.path.join(LIBDIR, "ssb_lts_proxy")
    deps = [src, os.path.join(root, "host", "ssb_lts.hpp"), os.path.join(root, "host", "ssb_host.hpp"),
            os.path.join(root, "include", "seissol_b200.h"), lib]
    if not _newer(exe, deps):
        return exe
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    nccl_flags = []
    candidates = [("/usr/include", "/usr/lib/x86_64-linux-gnu")]
    try:
        import nvidia.nccl as _n
        base = os.path.dirname(_n.__file__) if getattr(_n, "__file__", None) else list(_n.__path__)[0]
        candidates.append((os.path.join(base, "include"), os.path.join(base, "lib")))
    except Exception:
        pass
    for inc, libdir in candidates:
        so = [x for x in (os.listdir(libdir) if os.path.isdir(libdir) else []) if x.startswith("libnccl.so")]
        if os.path.exists(os.path.join(inc, "nccl.h")) and so:
            link = "-lnccl" if "libnccl.so" in so else f"-l:{sorted(so)[0]}"
            nccl_flags = ["-DSSB_WITH_NCCL", f"-I{inc}", f"-L{libdir}", link, f"-Wl,-rpath,{libdir}"]
            break
    cmd = ["g++", "-std=c++17", "-O2", src, f"-I{os.path.join(root, 'include')}", f"-I{os.path.join(root, 'host')}",
           f"-I{os.path.join(cuda, 'include')}", f"-L{LIBDIR}", "-lseissol_b200", f"-L{os.path.join(cuda, 'lib64')}",
           "-lcudart", "-Wl,-rpath,$ORIGIN", "-pthread", *nccl_flags, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"LTS host build failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[build] built {exe}" + (" (NCCL)" if nccl_flags else " (single rank: no NCCL found)"), flush=True)
    return exe


if __name__ == "__main__":
    build(force="--force" in sys.argv)
