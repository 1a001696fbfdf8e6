"""Compiles libcandela_b200.so in-tree for sm_100a with nvcc (no JIT cache, no torch extension).

The flags are part of the parity contract: no FMA contraction, IEEE division and
square root, no flush-to-zero, and -ffp-contract=off for the little host-side float code.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libcandela_b200.so"
SOURCES = ["context.cu", "kernels_traverse.cu", "kernels_wavefront.cu", "kernels_raygen.cu", "kernels_hot.cu", "builder.cu", "builder_lbvh.cu", "ray_order.cu", "model_loader.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
]
# This image's g++ wrapper finds only a static libstdc++ (its libstdc++.so symlink dangles); a
# static copy inside a dlopen()ed library clashes with the process's own. Link the system one.
SYSTEM_STDCXX = "/usr/lib/x86_64-linux-gnu/libstdc++.so.6"


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "candela_b200.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", str(LIB), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    if Path(SYSTEM_STDCXX).exists():
        cmd += ["-Xlinker", SYSTEM_STDCXX]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
