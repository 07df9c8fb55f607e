"""In-tree build of the sm_100a CUDA library behind include/mfsr.h.

`python -m multi_frame_super_resolution_b200.build` (or `__graft_entry__.build()`)
compiles every csrc/*.cu with nvcc for sm_100a into
multi_frame_super_resolution_b200/libmfsr_b200.so.  nvcc cross-compiles without a
GPU.  The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libmfsr_b200.so"
STAMP = PKG / "csrc" / ".build_stamp"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-I", str(ROOT / "include"),
] + os.environ.get("MFSR_NVCC_EXTRA", "").split()      # e.g. -DMFSR_LOOP_UNROLL=2 for same-box A/B builds (tools/ab_build.sh)


# Files whose fp32 results feed discrete decisions (quantisation, arg-min, rounding of tile
# shifts) are compiled without FMA contraction so they round exactly like strict IEEE code.
STRICT_FP = {"frontend.cu", "align.cu", "prealign.cu"}


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in list(_sources()) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "mfsr.h"]:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update((" ".join(FLAGS) + "|" + ",".join(sorted(STRICT_FP))).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    if not Path(NVCC).exists():
        if LIB.exists():      # GPU box without toolkit changes: use the prebuilt library
            return LIB
        raise RuntimeError(f"nvcc not found at {NVCC} and {LIB} is missing")
    objs = []
    procs = []
    for src in _sources():
        obj = src.with_suffix(".o")
        cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
        if src.name in STRICT_FP:
            cmd.insert(1, "-fmad=false")
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip() and (verbose or p.returncode != 0):
            print(f"--- {src.name}\n{out}", file=sys.stderr)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
    subprocess.check_call(cmd)
    STAMP.write_text(digest)
    return LIB


HOST_SRC = ROOT / "host" / "multi_frame_sr_b200.cpp"
HOST_BIN = ROOT / "host" / "multi_frame_sr_b200"


def build_host(force: bool = False) -> Path:
    """g++ build of the C++ host program (host/multi_frame_sr_b200.cpp): links the C ABI only, no CUDA headers."""
    lib = build()
    if not force and HOST_BIN.exists() and HOST_BIN.stat().st_mtime >= max(HOST_SRC.stat().st_mtime, lib.stat().st_mtime):
        return HOST_BIN
    cxx = os.environ.get("CXX", "g++")
    cuda_lib = str(Path(NVCC).resolve().parent.parent / "lib64")
    cmd = [cxx, "-O2", "-std=c++17", "-Wall", str(HOST_SRC), "-I", str(ROOT / "include"), "-L", str(PKG), "-lmfsr_b200",
           "-L", cuda_lib, "-lcudart", "-Wl,-rpath,$ORIGIN/../multi_frame_super_resolution_b200", f"-Wl,-rpath,{cuda_lib}",
           "-o", str(HOST_BIN)]
    subprocess.check_call(cmd)
    return HOST_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
