"""Build the in-tree native artefacts (sm_100a only; nvcc cross-compiles without a GPU).

    python -m steganosaurus_b200.build [--force]

Outputs (git-ignored, shipped to the GPU box by gpurun):
    steganosaurus_b200/libtfft_b200.so   CUDA kernels + the C ABI of include/tfft.h
    steganosaurus_b200/libtfft_host.so   host-side C++ (KDF, AEAD, turtlewalk, framing, PNG)
    steganosaurus_b200/turtlefft         drop-in CLI (embed / extract)
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_CUDA = os.path.join(PKG, "libtfft_b200.so")
LIB_HOST = os.path.join(PKG, "libtfft_host.so")
CLI = os.path.join(PKG, "turtlefft")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-march=x86-64-v3", "-Xcompiler", "-ffp-contract=off", "--expt-relaxed-constexpr",
]
CXX = shutil.which("g++") or "g++"
CXX_FLAGS = ["-std=c++17", "-O3", "-march=x86-64-v3", "-fPIC", "-Wall", "-Wextra"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd: list[str]) -> None:
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def build_cuda(force: bool = False) -> str:
    cu = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = cu + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(ROOT, "include", "tfft.h")]
    if force or _newer(LIB_CUDA, deps):
        _run([NVCC, *NVCC_FLAGS, "-shared", "-o", LIB_CUDA, *cu])
    return LIB_CUDA


def build_host(force: bool = False) -> str | None:
    hdir = os.path.join(CSRC, "host")
    cpp = sorted(glob.glob(os.path.join(hdir, "*.cpp")))
    lib_src = [c for c in cpp if not c.endswith("cli_main.cpp")]
    if not lib_src:
        return None
    deps = cpp + glob.glob(os.path.join(hdir, "*.h")) + [os.path.join(ROOT, "include", "tfft_host.h")]
    deps = [d for d in deps if os.path.exists(d)]
    if force or _newer(LIB_HOST, deps):
        _run([CXX, *CXX_FLAGS, "-shared", "-I", os.path.join(ROOT, "include"), "-o", LIB_HOST, *lib_src, "-lz", "-lpthread"])
    main = os.path.join(hdir, "cli_main.cpp")
    if os.path.exists(main) and (force or _newer(CLI, deps + [LIB_CUDA])):
        _run([CXX, *CXX_FLAGS, "-I", os.path.join(ROOT, "include"), "-o", CLI, main,
              "-L", PKG, "-ltfft_host", "-ltfft_b200", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"])
    return LIB_HOST


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_host(force)


if __name__ == "__main__":
    build_all("--force" in sys.argv)
