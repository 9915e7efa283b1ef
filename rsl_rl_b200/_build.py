"""Builds the in-tree CUDA library `rsl_rl_b200/lib/librslrl_b200.so` with nvcc for sm_100a.

The library is plain CUDA C++ behind the C ABI of `include/rslrl_b200.h` (no torch headers), so it is
compiled directly with nvcc -- seconds per file, cross-compiles without a GPU.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIBPATH = os.path.join(LIBDIR, "librslrl_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math=false",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build librslrl_b200.so")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def is_stale() -> bool:
    if not os.path.exists(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(PKG, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, out: str = None, extra_flags=()) -> str:
    """Compile every .cu under csrc/ into one shared library (object files in build/, parallel).
    `out` / `extra_flags` build an experimental variant next to the product library (kernel A/B runs, tracing)."""
    variant = out is not None or bool(extra_flags)
    out = out or LIBPATH
    if not force and not variant and not is_stale():
        return LIBPATH
    os.makedirs(os.path.dirname(out), exist_ok=True)
    objdir = os.path.join(PKG, "..", "build", "obj_variant" if variant else "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + list(extra_flags)
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *flags, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        log, _ = p.communicate()
        if verbose and log:
            print(log)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
    link = [nvcc, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc link failed:\n{r.stdout}")
    return out


if __name__ == "__main__":
    import sys

    # python -m rsl_rl_b200._build [--force] [-v] [--out PATH] [-DFLAG ...]
    argv = sys.argv[1:]
    out_path = argv[argv.index("--out") + 1] if "--out" in argv else None
    print(build_library(force="--force" in argv, verbose="-v" in argv, out=out_path,
                        extra_flags=[a for a in argv if a.startswith("-D")]))
