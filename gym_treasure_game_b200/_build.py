"""Builds ``libtreasure_b200.so`` in-tree with nvcc for sm_100a (no JIT cache:
the built library travels with the source tree)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
SO = os.path.join(PKG, "libtreasure_b200.so")
SOURCES = ["tg_step.cu", "tg_render.cu", "tg_blend.cu", "tg_capi.cu"]
HEADERS = ["tg_types.h", "tg_device.cuh", "tg_launch.h", "tg_host_patch.h", os.path.join("..", "..", "include", "treasure_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fopenmp", "-shared", "-diag-suppress", "550", "-lgomp"]


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return SO
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", SO] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd)
    return SO
