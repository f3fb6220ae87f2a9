"""In-tree build of the CUDA engine (libgpx.so) for sm_100a and of the CPU oracle.

nvcc cross-compiles without a GPU; the built .so files are git-ignored but travel to the
GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgpx.so")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libgpx_oracle.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-rdc=true",  # k_round launches k_round_slow from the device (tail launch): relocatable device code + cudadevrt
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-shared",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources() -> list[str]:
    out = [os.path.join(ROOT, "include", "gpx.h"), os.path.join(ROOT, "include", "gpx_wire.h")]
    for f in sorted(os.listdir(CSRC)):
        out.append(os.path.join(CSRC, f))
    return [s for s in out if os.path.exists(s)]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; the gpx engine has no CPU fallback")
    return p


def build_engine(force: bool = False, verbose: bool = False, out: str = LIB, defines: tuple = ()) -> str:
    srcs = _sources()
    if not force and _newer(out, srcs):
        return out
    units = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cpp"))]
    cmd = [nvcc_path(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           "-o", out, *units, "-ldl", "-lcudadevrt"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return out


def build_oracle(force: bool = False) -> str:
    src = os.path.join(ORACLE_DIR, "gpx_oracle.cpp")
    if not force and _newer(ORACLE_LIB, [src, os.path.join(ROOT, "include", "gpx.h")]):
        return ORACLE_LIB
    r = subprocess.run(["make", "-C", ORACLE_DIR, "-B" if force else "-s"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return ORACLE_LIB


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_engine(force=force, verbose="-v" in sys.argv))
    print(build_oracle(force=force))
