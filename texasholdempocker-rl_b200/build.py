"""Builds the sm_100a library (and the test-only host-check library) in-tree with nvcc / g++.

Outputs land in texasholdempocker-rl_b200/lib/ (git-ignored, shipped to the GPU box by gpurun).
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libphe_b200.so")
HOSTCHECK = os.path.join(LIBDIR, "libphe_hostcheck.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall", "--shared", "-cudart", "static",
]


def _sources(names):
    return [os.path.join(CSRC, n) for n in names]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _all_deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "phe_b200.h"))
    return deps


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_library(force=False, verbose=False, extra_flags=(), out=None):
    """`extra_flags` / `out` build an experiment variant next to the product library (tools/lab/mc_variants.py)."""
    os.makedirs(LIBDIR, exist_ok=True)
    target = out or LIB
    os.makedirs(os.path.dirname(target), exist_ok=True)
    if force or _stale(target, _all_deps()):
        cmd = ([nvcc_path()] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", target] +
               _sources(["phe_kernels.cu", "phe_nn.cu", "phe_obs.cu", "phe_tables.cpp"]))
        subprocess.check_call(cmd)
    return target


def build_hostcheck(force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    if force or _stale(HOSTCHECK, _all_deps()):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-x", "c++"] + _sources(["phe_hostcheck.cpp", "phe_tables.cpp"]) + ["-o", HOSTCHECK]
        subprocess.check_call(cmd)
    return HOSTCHECK


PEAKS = os.path.join(LIBDIR, "phe_peaks")


def build_peaks(force=False):
    """On-box microbenchmarks for the integer / shared-memory roofline denominators (csrc/phe_peaks.cu)."""
    os.makedirs(LIBDIR, exist_ok=True)
    src = os.path.join(CSRC, "phe_peaks.cu")
    if force or _stale(PEAKS, [src]):
        subprocess.check_call([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-o", PEAKS, src])
    return PEAKS


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_hostcheck(force=True))
    print(build_peaks(force=True))
