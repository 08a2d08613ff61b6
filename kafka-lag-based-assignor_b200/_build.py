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
    """One nvcc -c per .cu (in parallel: the sort templates are slow to compile), then link."""
    from concurrent.futures import ThreadPoolExecutor
    cus = _sources("", (".cu",))
    hdrs = _sources("", (".cuh",)) + [os.path.join(ROOT, "include", "lagassign.h")]
    objdir = os.path.join(PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    hdr_time = _newest(hdrs)
    nvcc = _nvcc()
    compile_flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                     "-Xcompiler", "-fPIC"] + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(cu):
        obj = os.path.join(objdir, os.path.basename(cu)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(cu), hdr_time):
            return obj, ""
        r = subprocess.run([nvcc] + compile_flags + ["-c", "-o", obj, cu], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {cu}:\n" + r.stdout + r.stderr)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=len(cus)) as ex:
        results = list(ex.map(compile_one, cus))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest(objs):
        r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
                            "-Xcompiler", "-fPIC", "-o", LIB] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    return LIB


def build_host(force: bool = False) -> str:
    hdir = os.path.join(CSRC, "host")
    if not os.path.isdir(hdir):
        return ""
    srcs = _sources("host", (".cpp",))
    deps = srcs + _sources("host", (".hpp", ".h")) + [os.path.join(ROOT, "include", "lagassign.h"),
                                                       os.path.join(ROOT, "include", "lagassign_host.h")]
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
