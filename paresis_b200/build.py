"""Build the CUDA library in-tree: nvcc, sm_100a only, C ABI, no torch headers.

    python -m paresis_b200.build [--force] [--verbose] [--bounds-check] [--variant NAME -DMACRO=VALUE ...]

The resulting ``paresis_b200/libparesis_b200.so`` is git-ignored but travels with the tree.  ``--bounds-check`` builds
``libparesis_b200_checked.so`` instead (-DPARESIS_BOUNDS_CHECK: device asserts on the shared-memory and queue indices of the
newer kernels); ``PARESIS_B200_LIB=libparesis_b200_checked.so`` makes the binding load it.
"""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libparesis_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--use_fast_math=false", "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(PKG, "..", "include", "paresis_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(PKG, "..", "include", "paresis_b200.h")]


def build_library(force=False, verbose=False, bounds_check=False, variant=None, defines=()):
    """Compile every .cu to an object (in parallel, only the stale ones) and link the shared library.
    ``variant`` + ``defines``: an experimental build next to the production one (libparesis_b200_<variant>.so, its own object
    directory, extra -D flags), to be A/B-ed on one box with PARESIS_B200_LIB."""
    lib = os.path.join(PKG, "libparesis_b200_checked.so") if bounds_check else LIB
    if variant:
        lib = os.path.join(PKG, "libparesis_b200_%s.so" % variant)
    if not force and not bounds_check and not variant and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(CSRC, "_obj_" + variant if variant else ("_obj_checked" if bounds_check else "_obj"))
    os.makedirs(objdir, exist_ok=True)
    flags = ([f for f in FLAGS if f not in ("--use_fast_math=false", "-shared")] + (["-DPARESIS_BOUNDS_CHECK"] if bounds_check else [])
             + ["-D" + d for d in defines])
    newest_header = max(os.path.getmtime(h) for h in _headers())

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_header):
            return obj
        cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources()))
    # --cudart shared: the runtime is the image's libcudart.so, not a private static copy inside the library
    # the arch on the link line too: without it nvcc embeds an (empty) cubin of its default architecture, sm_52
    cmd = [NVCC, "-shared", "--cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-lcufft", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    _defines = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv, bounds_check="--bounds-check" in sys.argv,
                        variant=_variant, defines=_defines))
