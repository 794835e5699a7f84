"""Builds the CUDA library (and the C++ host program) in-tree with nvcc for sm_100a.

The shared library is the product: a C ABI (include/asw_b200.h) over hand-written CUDA
kernels.  It is built in-tree (stereo_matchin_b200/libasw_b200.so) so that it travels to
the GPU box with the repository snapshot; nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libasw_b200.so")
HOST_BIN = os.path.join(ROOT, "src", "host", "stereo_matching")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build libasw_b200.so")


def _host_cxx() -> str:
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); the system g++ is the safe choice
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def lib_sources() -> list[str]:
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(ROOT, "include", "asw_b200.h"))
    return srcs


def build_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = lib_sources()
    if not force and not _stale(LIB, srcs):
        return LIB
    cus = [s for s in srcs if s.endswith(".cu")]
    cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", _host_cxx(), "-shared", "-o", LIB, *cus]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


UBENCH_LIB = os.path.join(PKG, "libasw_ubench.so")


def build_ubench(force: bool = False) -> str:
    """Measurement helpers (FP32 peak calibration, shared-memory probes); not part of the ABI."""
    src = os.path.join(CSRC, "ubench", "asw_ubench.cu")
    if not force and not _stale(UBENCH_LIB, [src]):
        return UBENCH_LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", _host_cxx(), "-shared", "-o", UBENCH_LIB, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return UBENCH_LIB


def build_host(force: bool = False) -> str:
    hdir = os.path.join(ROOT, "src", "host")
    srcs = [os.path.join(hdir, f) for f in sorted(os.listdir(hdir)) if f.endswith((".cpp", ".h"))]
    if not force and not _stale(HOST_BIN, srcs + [LIB]):
        return HOST_BIN
    cpps = [s for s in srcs if s.endswith(".cpp")]
    cmd = [_host_cxx(), "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", HOST_BIN, *cpps,
           "-L", PKG, "-lasw_b200", "-lz", "-Wl,-rpath,$ORIGIN/../../stereo_matchin_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return HOST_BIN


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_ubench(force="--force" in sys.argv))
    if os.path.isdir(os.path.join(ROOT, "src", "host")) and any(f.endswith(".cpp") for f in os.listdir(os.path.join(ROOT, "src", "host"))):
        print(build_host(force="--force" in sys.argv))
