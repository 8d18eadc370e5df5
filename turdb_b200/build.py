"""In-tree build of libturdb_cuda.so (nvcc, sm_100a only).  `python -m turdb_b200.build`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libturdb_cuda.so")
SOURCES = ["turdb_cuda.cu"]
DEPS = ["turdb_cuda.cu", "common.cuh", "hnsw_search.cuh", "exact_search.cuh", "exact_abi.inl", "gather_probe.cuh", "hnsw_file.inl", "sql_topk.inl",
        os.path.join("..", "..", "include", "turdb_cuda.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libturdb_cuda.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd, cwd=HERE)
    return LIB


def build_variant(name: str, defines: list[str]) -> str:
    """A/B builds for measurement (tools/sweep.py --libs): same sources, extra -D switches, loaded through the
    TURDB_CUDA_LIB override of _lib.py.  Written next to the library as libturdb_cuda.<name>.so."""
    out = os.path.join(HERE, f"libturdb_cuda.{name}.so")
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", out, *[os.path.join(CSRC, s) for s in SOURCES]]
    subprocess.check_call(cmd, cwd=HERE)
    return out


if __name__ == "__main__":
    for arg in sys.argv[1:]:
        if arg.startswith("--variant="):  # --variant=name:DEF1=1,DEF2=0
            vname, _, defs = arg[len("--variant="):].partition(":")
            print(build_variant(vname, [d for d in defs.split(",") if d]))
            sys.exit(0)
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
