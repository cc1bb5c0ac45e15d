"""Build recipe of libtophat_b200.so and the host binaries (explicit nvcc / g++, in-tree outputs).

sm_100a only: `-gencode arch=compute_100a,code=sm_100a`.  The cudart is linked statically so the
.so travels to the GPU box without extra dependencies; NCCL is bound at run time (dlopen).
"""
from __future__ import annotations

import hashlib
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


def _digest(sources, extra: str = "") -> str:
    """Content hash of the sources (+ the command line): checkout times say nothing about what a binary was built from."""
    h = hashlib.sha256(extra.encode())
    for s in sorted(sources):
        h.update(os.path.relpath(s, ROOT).encode())
        with open(s, "rb") as f:
            h.update(hashlib.sha256(f.read()).digest())
    return h.hexdigest()


def _stale(target: str, digest: str) -> bool:
    stamp = target + ".srchash"
    if not os.path.exists(target) or not os.path.exists(stamp):
        return True
    with open(stamp) as f:
        return f.read().strip() != digest


def _stamp(target: str, digest: str) -> None:
    with open(target + ".srchash", "w") as f:
        f.write(digest + "\n")


def _csrc_files():
    out = []
    for d, _, fs in os.walk(CSRC):
        out += [os.path.join(d, f) for f in fs if f.endswith((".cu", ".cuh", ".h", ".cpp", ".hpp"))]
    out.append(os.path.join(ROOT, "include", "tophat_b200.h"))
    return out


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, "thb_api.cu")]
    lib_srcs = [f for f in _csrc_files() if os.sep + "host" + os.sep not in f]
    digest = _digest(lib_srcs, " ".join(NVCC_FLAGS))
    if force or _stale(LIB, digest):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + srcs + ["-ldl"]
        subprocess.run(cmd, check=True)
        _stamp(LIB, digest)
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
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Wno-unused-function", "-Wno-misleading-indentation", "-I", os.path.join(ROOT, "include"), "-I", host,
               "-o", exe, main] + common + ["-L", HERE, "-ltophat_b200", "-Wl,-rpath,$ORIGIN/..", "-lz", "-lpthread"]
        host_srcs = [main] + common + [f for f in _csrc_files() if f.endswith((".h", ".hpp"))]
        digest = _digest(host_srcs, " ".join(cmd))
        if force or _stale(exe, digest):
            subprocess.run(cmd, check=True)
            _stamp(exe, digest)
        outs.append(exe)
    return outs


def build_all(force: bool = False, verbose: bool = False):
    lib = build_library(force, verbose)
    bins = build_host_binaries(force)
    return lib, bins


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
