"""Builds libw3codec.so (the CUDA kernels + C ABI) in-tree for sm_100a.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the
GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libw3codec.so")
CLI = os.path.join(HERE, "w3")
SOURCES = ["w3_kernels.cu", "w3_tierb.cu", "w3_geo.cu", "w3_cm.cu", "w3_capi.cu"]
DEPS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + [
    os.path.join("..", "..", "include", "w3_codec.h")
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall",
]


def _nvcc() -> str:
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found — the CUDA toolkit is required to build libw3codec.so")
    return path


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if force or _stale(LIB, DEPS):
        cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return LIB


def build_cli(force: bool = False) -> str:
    build_lib(force)
    src = os.path.join(CSRC, "w3_cli.cpp")
    if force or _stale(CLI, ["w3_cli.cpp", os.path.join("..", "..", "include", "w3_codec.h")]) or (
        os.path.getmtime(LIB) > os.path.getmtime(CLI)
    ):
        subprocess.check_call(
            ["g++", "-O2", "-std=c++17", "-Wall", "-o", CLI, src, "-L" + HERE, "-lw3codec",
             "-Wl,-rpath,$ORIGIN"]
        )
    return CLI


if __name__ == "__main__":
    print(build_lib(verbose=True))
    print(build_cli())
