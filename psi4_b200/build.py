"""Compile libb200jk.so (sm_100a only) in-tree.  nvcc cross-compiles without a GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200jk.so")
SOURCES = ["engine.cu"]
HEADERS = ["dmma_gemm.cuh", "dmma_ws.cuh", "i8_kgemm.cuh", "i8_half.cuh", "fit_kernels.cuh", "fit_host.inl", "j_kernels.cuh", "aux_kernels.cuh", "peer_reduce.cuh", "peer_host.inl", "gemm_strided.cuh", "grad_host.inl", "power_host.inl", os.path.join("..", "..", "include", "b200jk.h")]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB


INTS_LIB = os.path.join(HERE, "libb200ints.so")


def build_ints(force: bool = False) -> str:
    """Host-side integral front end (plain C, gcc + OpenMP)."""
    src = os.path.join(CSRC, "ints.c")
    if not force and os.path.exists(INTS_LIB) and os.path.getmtime(INTS_LIB) >= os.path.getmtime(src):
        return INTS_LIB
    subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-std=gnu11", "-o",
                           INTS_LIB, src, "-lm"])
    return INTS_LIB


if __name__ == "__main__":
    print(build_ints(force="--force" in sys.argv))
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
