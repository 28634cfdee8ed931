"""Builds sdrainer_b200/libsdrgpu.so in-tree with nvcc for sm_100a (no torch dependency)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsdrgpu.so")
SOURCES = ["engine.cu", "goertzel.cu"]
HEADERS = ["fft_radix.cuh", "k1_spectral.cuh", "k1_large.cuh", "k1_mid4k.cuh", "k1_mid8k.cuh", "k1_warp.cuh", "k1_wide.cuh", "k2_post.cuh", "k3_goertzel.cuh"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; cannot build libsdrgpu.so")
    return p


def _deps():
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(HERE, "..", "include", "sdrgpu.h"))
    return deps


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


LOADED = False  # set by capi.lib(): this process has dlopen'ed LIB


def build(force: bool = False, verbose: bool = False) -> str:
    # A loaded library cannot be swapped: rebuilding it in place would give every later dlopen (libsdrhost.so's
    # dependency) a SECOND copy whose kernels never had their attributes set by the engine the first copy created.
    if not force and (LOADED or not is_stale()):
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return LIB


HOST_LIB = os.path.join(HERE, "libsdrhost.so")
HOST_SOURCES = [os.path.join(HERE, "host", "host_capi.cpp"), os.path.join(HERE, "host", "sdrhost.hpp"),
                os.path.join(HERE, "host", "realtime.hpp")]


def build_host(force: bool = False) -> str:
    """libsdrhost.so: the C++ host mirror of the Go interface (dsp/cw/rx) over the C ABI, linked to libsdrgpu.so.
    -ffp-contract=off: the decoder's float64 arithmetic follows Go/amd64 (no fused multiply-add)."""
    build(force=False)
    stale = (not os.path.exists(HOST_LIB)) or any(os.path.getmtime(d) > os.path.getmtime(HOST_LIB) for d in HOST_SOURCES + [LIB])
    if force or stale:
        subprocess.check_call(["g++", "-O2", "-mavx2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-o", HOST_LIB,
                               HOST_SOURCES[0], "-L" + HERE, "-lsdrgpu", "-lpthread", "-Wl,-rpath,$ORIGIN"])
    return HOST_LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
    print(build_host(force=True))
