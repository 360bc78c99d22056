"""Build recipes for the native libraries (all in-tree, so the .so files travel with the repo).

* ``libfmgpu.so``  — the product: CUDA kernels + C ABI (nvcc, sm_100a only).
* ``libfmhost.so`` — host-side index producer / workload generator (g++).
* ``oracle/liboracle.so`` — the CPU oracle (test infrastructure; built here, used only by tests,
  smoke() and bench.py's CPU baseline).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "index4j_b200")
CSRC = os.path.join(PKG, "csrc")

GPU_LIB = os.path.join(PKG, "libfmgpu.so")
HOST_LIB = os.path.join(PKG, "libfmhost.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math", "-Xcompiler", "-fPIC,-O3,-pthread", "-shared", "-Xptxas", "-v",
    "-I", os.path.join(ROOT, "include"),
]
CXX_FLAGS = ["-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-Wextra"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(d: str, exts=(".cu", ".cuh", ".h", ".hpp", ".cpp", ".inc")) -> list[str]:
    out = []
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return out


def _run(cmd: list[str], log_name: str | None = None) -> str:
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log_name:
        with open(os.path.join(PKG, log_name), "w") as fh:
            fh.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return p.stdout


def build_host(force: bool = False) -> str:
    srcs = _sources(os.path.join(CSRC, "host"))
    if force or not _newer(HOST_LIB, srcs):
        _run(["g++", *CXX_FLAGS, "-o", HOST_LIB, os.path.join(CSRC, "host", "fmhost.cpp")])
    return HOST_LIB


def build_gpu(force: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [s for s in _sources(CSRC) if os.sep + "host" + os.sep not in s] + _sources(os.path.join(ROOT, "include"))
    if force or not _newer(GPU_LIB, srcs):
        _run([nvcc, *NVCC_FLAGS, "-o", GPU_LIB, os.path.join(CSRC, "fmgpu.cu"), "-lcudart"], "nvcc_build.log")
    return GPU_LIB


def build_oracle(force: bool = False) -> str:
    src = os.path.join(ROOT, "oracle", "oracle.cpp")
    if force or not _newer(ORACLE_LIB, [src]):
        _run(["make", "-C", os.path.join(ROOT, "oracle"), "-B" if force else "-s"])
    return ORACLE_LIB


def build_all(force: bool = False) -> None:
    build_host(force)
    build_gpu(force)
    build_oracle(force)
