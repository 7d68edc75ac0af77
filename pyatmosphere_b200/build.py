"""Builds libpyatm_b200.so (hand-written sm_100a CUDA behind the C ABI of include/pyatm_b200.h) in-tree.

    python -m pyatmosphere_b200.build [--force]

nvcc cross-compiles without a GPU; objects go to csrc/_build (git- and gpurun-ignored), the shared library
to pyatmosphere_b200/libpyatm_b200.so (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import concurrent.futures
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, os.environ.get("PYATM_OBJ_DIR", "_build"))
LIB = os.environ.get("PYATM_LIB", os.path.join(PKG, "libpyatm_b200.so"))      # override: experimental variants
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"] + os.environ.get("PYATM_NVCC_FLAGS", "").split()


def _deps_mtime():
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(CSRC, "*.inc")) + glob.glob(os.path.join(PKG, "..", "include", "*.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, obj, verbose):
    cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {os.path.basename(src)}:\n{r.stdout}\n{r.stderr}")
    if verbose and (r.stdout.strip() or r.stderr.strip()):
        print(r.stdout, r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdr_time = _deps_mtime()
    todo, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_time):
            todo.append((s, o))
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
            list(ex.map(lambda so: _compile(so[0], so[1], verbose), todo))
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-cudart", "static", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
