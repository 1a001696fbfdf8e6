"""Compiles libcandela_b200.so in-tree for sm_100a with nvcc (no JIT cache, no torch extension).

The flags are part of the parity contract: no FMA contraction, IEEE division and
square root, no flush-to-zero, and -ffp-contract=off for the little host-side float code.

Each translation unit is compiled to candela_b200/_obj/<name>.o (in parallel, only when it or a
header is newer than its object) and the objects are linked into the shared library.  `last_report`
says what the last build() call did, so that the driver's build check can tell a compile from a reuse.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "_obj"
LIB = PKG / "libcandela_b200.so"
SOURCES = ["context.cu", "frame.cu", "kernels_traverse.cu", "kernels_wavefront.cu", "kernels_raygen.cu", "kernels_hot.cu", "builder.cu", "builder_lbvh.cu",
           "ray_order.cu", "model_loader.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
]
# This image's g++ wrapper finds only a static libstdc++ (its libstdc++.so symlink dangles); a
# static copy inside a dlopen()ed library clashes with the process's own. Link the system one.
SYSTEM_STDCXX = "/usr/lib/x86_64-linux-gnu/libstdc++.so.6"

last_report: dict = {}


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _headers():
    return list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "candela_b200.h", Path(__file__)]


def stale_sources(force: bool = False):
    newest_header = max(h.stat().st_mtime for h in _headers())
    out = []
    for s in SOURCES:
        o = OBJ / (Path(s).stem + ".o")
        if force or not o.exists() or o.stat().st_mtime < max((CSRC / s).stat().st_mtime, newest_header):
            out.append(s)
    return out


def is_stale() -> bool:
    if not LIB.exists() or stale_sources():
        return True
    t = LIB.stat().st_mtime
    return any((OBJ / (Path(s).stem + ".o")).stat().st_mtime > t for s in SOURCES)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Returns the library path.  last_report = {"compiled": [sources], "linked": bool, "reused": bool}."""
    global last_report
    todo = stale_sources(force)
    if not todo and not is_stale():
        last_report = {"compiled": [], "linked": False, "reused": True}
        return LIB
    OBJ.mkdir(exist_ok=True)
    nvcc = nvcc_path()

    def compile_one(src: str):
        cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", str(OBJ / (Path(src).stem + ".o")), str(CSRC / src)]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        return src, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4) or 1) as pool:
        results = list(pool.map(compile_one, todo))
    for src, proc in results:
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + proc.stdout + proc.stderr)
        if verbose:
            print(proc.stderr)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *[str(OBJ / (Path(s).stem + ".o")) for s in SOURCES]]
    if Path(SYSTEM_STDCXX).exists():
        cmd += ["-Xlinker", SYSTEM_STDCXX]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    last_report = {"compiled": todo, "linked": True, "reused": False}
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv), last_report)
