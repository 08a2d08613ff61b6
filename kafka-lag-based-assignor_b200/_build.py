"""Builds the CUDA library in-tree (nvcc, sm_100a only).

    python -m kafka_lag_based_assignor_b200._build        (or __graft_entry__.build())

Outputs kafka-lag-based-assignor_b200/liblagassign.so (kernels + C ABI of
include/lagassign.h) and kafka-lag-based-assignor_b200/liblagassign_host.so (the
C++ host mirror of LagBasedPartitionAssignor, which links the former).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
ROOT = os.path.dirname(PKG_DIR)
LIB = os.path.join(PKG_DIR, "liblagassign.so")
HOST_LIB = os.path.join(PKG_DIR, "liblagassign_host.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def _sources(sub, exts):
    d = os.path.join(CSRC, sub) if sub else CSRC
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    deps = _sources("", (".cu", ".cuh")) + [os.path.join(ROOT, "include", "lagassign.h")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(deps):
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB, os.path.join(CSRC, "lagassign_api.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


def build_host(force: bool = False) -> str:
    hdir = os.path.join(CSRC, "host")
    if not os.path.isdir(hdir):
        return ""
    srcs = _sources("host", (".cpp",))
    deps = srcs + _sources("host", (".hpp", ".h")) + [os.path.join(ROOT, "include", "lagassign.h")]
    if not srcs:
        return ""
    if not force and os.path.exists(HOST_LIB) and os.path.getmtime(HOST_LIB) >= max(_newest(deps), os.path.getmtime(LIB)):
        return HOST_LIB
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra",
           "-I", os.path.join(ROOT, "include"), "-o", HOST_LIB] + srcs + \
          ["-L", PKG_DIR, "-llagassign", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return HOST_LIB


def build_all(force: bool = False, verbose: bool = False):
    return build_cuda(force, verbose), build_host(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
