"""Build the C-ABI shared library nbldpc_b200/libnbldpc_b200.so for sm_100a with nvcc (in-tree)."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libnbldpc_b200.so")
SOURCES = ["decoder.cu", "encoder.cu", "code.cpp", "comm.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              # parity mode must not contract a*b+c (e.g. 0.25*old + 0.75*new, NBLDPC.cpp:1046)
              "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-ldl"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, "..", "include", "nbldpc_b200.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB


HOST = os.path.join(PKG, "host")
HOST_COMMON = ["link_chain.cpp", "codec.cpp", "gf_field.cpp", "sim_control.cpp"]
HOST_PROGRAMS = {"nbldpc_sim": "driver.cpp", "harness_dump": "harness_dump.cpp"}


def build_host(force=False):
    """C++ harness (the reference's CSimulation / CComm / CNBLDPC call surface on top of the C ABI)"""
    build()
    out = []
    for exe, main in HOST_PROGRAMS.items():
        path = os.path.join(HOST, exe)
        srcs = [os.path.join(HOST, f) for f in [main] + HOST_COMMON]
        deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")] + [LIB]
        if force or not os.path.exists(path) or any(os.path.getmtime(d) > os.path.getmtime(path) for d in deps):
            subprocess.check_call([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-Wall", "-ffp-contract=off", "-o", path] + srcs +
                                  ["-L" + PKG, "-lnbldpc_b200", "-Wl,-rpath,$ORIGIN/.."])
        out.append(path)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
