"""Build recipe of libtophat_b200.so and the host binaries (explicit nvcc / g++, in-tree outputs).

sm_100a only: `-gencode arch=compute_100a,code=sm_100a`.  The cudart is linked statically so the
.so travels to the GPU box without extra dependencies; NCCL is bound at run time (dlopen).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtophat_b200.so")
BIN_DIR = os.path.join(HERE, "bin")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
              "-I", os.path.join(ROOT, "include")]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _csrc_files():
    out = []
    for d, _, fs in os.walk(CSRC):
        out += [os.path.join(d, f) for f in fs if f.endswith((".cu", ".cuh", ".h", ".cpp", ".hpp"))]
    out.append(os.path.join(ROOT, "include", "tophat_b200.h"))
    return out


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, "thb_api.cu")]
    if force or _newer(LIB, _csrc_files()):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + srcs + ["-ldl"]
        subprocess.run(cmd, check=True)
    return LIB


def build_host_binaries(force: bool = False) -> list:
    """segment_juncs / long_spanning_reads drop-in executables (C++ hosts over the C ABI)."""
    host = os.path.join(CSRC, "host")
    outs = []
    if not os.path.isdir(host):
        return outs
    os.makedirs(BIN_DIR, exist_ok=True)
    common = sorted(os.path.join(host, f) for f in os.listdir(host) if f.endswith(".cpp") and not f.endswith("_main.cpp"))
    for prog in ("segment_juncs", "long_spanning_reads", "thb_host_selftest"):
        main = os.path.join(host, prog + "_main.cpp")
        if not os.path.exists(main):
            continue
        exe = os.path.join(BIN_DIR, prog)
        if force or _newer(exe, _csrc_files() + [LIB]):
            cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Wno-unused-function", "-Wno-misleading-indentation", "-I", os.path.join(ROOT, "include"), "-I", host,
                   "-o", exe, main] + common + ["-L", HERE, "-ltophat_b200", "-Wl,-rpath,$ORIGIN/..", "-lz", "-lpthread"]
            subprocess.run(cmd, check=True)
        outs.append(exe)
    return outs


def build_all(force: bool = False, verbose: bool = False):
    lib = build_library(force, verbose)
    bins = build_host_binaries(force)
    return lib, bins


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
