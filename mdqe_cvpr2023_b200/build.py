"""Build libmsda_b200.so (hand-written sm_100a CUDA + the C ABI in include/msda_b200.h) in-tree.

    python -m mdqe_cvpr2023_b200.build [--force] [-v]

nvcc cross-compiles without a GPU.  Every translation unit is compiled to an object file on its own thread
(incrementally: only units older than their sources or than any header are rebuilt) and the objects are linked
into one shared object next to this file, so that it travels to the GPU box with the source snapshot; objects
and library are git-ignored.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "_obj")
LIB_PATH = os.path.join(PKG_DIR, "libmsda_b200.so")
SOURCES = ["msda_api.cu", "msda_launch_f32.cu", "msda_launch_bf16.cu", "msda_launch_bf16_loc32.cu", "msda_launch_f64.cu",
           "mask_gemm.cu", "consumers.cu", "allreduce.cu", "msda_packed.cu"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMPILE_FLAGS = ARCH_FLAGS + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]
NVCC_FLAGS = COMPILE_FLAGS + ["-shared"]          # one-shot form (tools/ build experiment libraries with it)


def _nvcc():
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _header_mtime():
    newest = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(PKG_DIR), "include")):
        for name in os.listdir(root):
            if name.endswith((".cuh", ".h")):
                newest = max(newest, os.path.getmtime(os.path.join(root, name)))
    return newest


def _compile(src, obj, extra, verbose):
    cmd = [_nvcc()] + COMPILE_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return src, res.returncode, res.stdout


def build(force=False, verbose=False, extra_flags=(), lib_path=LIB_PATH, obj_dir=OBJ_DIR):
    """Compile what is out of date and link.  Returns the library path."""
    os.makedirs(obj_dir, exist_ok=True)
    hdr = _header_mtime()
    jobs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src[:-3] + ".o")
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr, os.path.getmtime(os.path.join(CSRC, src)))
        if stale:
            jobs.append((src, obj))
    objs = [os.path.join(obj_dir, src[:-3] + ".o") for src in SOURCES]
    if not jobs and os.path.exists(lib_path) and os.path.getmtime(lib_path) >= max(os.path.getmtime(o) for o in objs):
        return lib_path
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as pool:
        results = list(pool.map(lambda j: _compile(j[0], j[1], extra_flags, verbose), jobs))
    failed = [(src, out) for src, rc, out in results if rc != 0]
    if verbose:
        for src, rc, out in results:
            sys.stderr.write(f"== {src}\n{out}")
    if failed:
        raise RuntimeError("nvcc failed building libmsda_b200.so:\n" + "\n".join(f"== {src}\n{out[-4000:]}" for src, out in failed))
    res = subprocess.run([_nvcc()] + ARCH_FLAGS + ["-shared", "-o", lib_path] + objs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("linking libmsda_b200.so failed:\n" + res.stdout[-4000:])
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
