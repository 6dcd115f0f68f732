"""Build libmsda_b200.so (hand-written sm_100a CUDA + the C ABI in include/msda_b200.h) in-tree.

    python -m mdqe_cvpr2023_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared object lands next to this file so that it travels to
the GPU box with the source snapshot; it is git-ignored.
"""
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libmsda_b200.so")
SOURCES = ["msda_api.cu", "mask_gemm.cu", "consumers.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newest_source_mtime():
    newest = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(PKG_DIR), "include")):
        for name in os.listdir(root):
            if name.endswith((".cu", ".cuh", ".h")):
                newest = max(newest, os.path.getmtime(os.path.join(root, name)))
    return newest


def build(force=False, verbose=False):
    """Compile the library if it is missing or older than its sources.  Returns its path."""
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libmsda_b200.so:\n" + res.stdout[-4000:])
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
