"""Builds lego_slam_b200/liblego_klt.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the resulting .so
travels to the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("LEGO_KLT_LIB") or os.path.join(HERE, "liblego_klt.so")  # override: tuning experiments only
SOURCES = ["lego_klt_capi.cu", "pyramid_sm100.cu", "klt_solver_exact.cu", "klt_solver_warp.cu", "klt_solver_patch.cu",
           "klt_solver_lane.cu", "klt_solver_lane_p8.cu", "klt_solver_lane_p11.cu", "klt_solver_lane_inv.cu",
           "triangulate_sm100.cu", "gftt_sm100.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "0",
    # parity: the reference CPU build has no fused multiply-add; fused ops are
    # written explicitly as fma() only where the product is exact
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-ldl",   # (-ldl: the header-only NVTX v3 looks its injection library up at run time)
]


def _nvcc() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "lego_klt.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Builds the library if it is missing or older than its sources.  Safe under `torchrun` (every rank calls it):
    an exclusive file lock serialises the ranks, the link goes to a temporary file and is renamed into place."""
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():      # another rank built it while we waited
                return LIB
            srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
            defs = os.environ.get("LEGO_KLT_NVCC_DEFS", "").split()
            tmp = f"{LIB}.tmp.{os.getpid()}"
            cmd = [_nvcc()] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + srcs
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building liblego_klt.so")
            os.replace(tmp, LIB)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
