"""Build the REFERENCE's own CUDA extension for sm_100a -- comparison column / second parity witness.

TEST + BENCH INFRASTRUCTURE ONLY; nothing under mdqe_cvpr2023_b200/ loads it.

The sources are compiled from where they lie (/root/reference/mdqe/models/ops/src); nothing is copied
into the repository.  They do not compile unmodified against torch 2.11: exactly two lines,
src/cuda/ms_deform_attn_cuda.cu:64 and :134, use the removed `value.type()` dispatch.  The recipe writes
a patched copy of that ONE file under /tmp (value.type() -> value.scalar_type() on those two lines),
compiles with nvcc for sm_100a, and leaves only the shared object in oracle/_ref/ (git-ignored, but it
travels to the GPU box).  The reference's setup.py is not used (it refuses to build without a visible GPU,
ops/setup.py:36-47).

    python oracle/build_ref_cuda.py        # needs /root/reference; run in the build container
"""
import os
import shutil
import subprocess
import sys
import sysconfig

REF_SRC = "/root/reference/mdqe/models/ops/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "msda_reference_cuda"
OUT = os.path.join(OUT_DIR, NAME + ".so")


def build(force=False):
    if os.path.exists(OUT) and not force:
        return OUT
    if not os.path.isdir(REF_SRC):
        raise RuntimeError(f"{REF_SRC} not present (only available in the build container)")
    import torch
    from torch.utils import cpp_extension as ce
    tmp = "/tmp/msda_ref_build"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    os.makedirs(OUT_DIR, exist_ok=True)
    cu = open(os.path.join(REF_SRC, "cuda", "ms_deform_attn_cuda.cu")).read()
    assert cu.count("AT_DISPATCH_FLOATING_TYPES(value.type()") == 2
    patched = os.path.join(tmp, "ms_deform_attn_cuda_patched.cu")
    open(patched, "w").write(cu.replace("AT_DISPATCH_FLOATING_TYPES(value.type()", "AT_DISPATCH_FLOATING_TYPES(value.scalar_type()"))
    inc = [f"-I{p}" for p in ce.include_paths(device_type="cuda")] + [f"-I{sysconfig.get_paths()['include']}", f"-I{REF_SRC}"]
    common = ["-O3", "-std=c++17", "-DWITH_CUDA", f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI)), "-w"]
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    for src, is_cu in ((patched, True), (os.path.join(REF_SRC, "vision.cpp"), False),
                       (os.path.join(REF_SRC, "cpu", "ms_deform_attn_cpu.cpp"), False)):
        obj = os.path.join(tmp, os.path.basename(src) + ".o")
        if is_cu:
            cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-c", src, "-o", obj,
                   "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                   "-D__CUDA_NO_HALF2_OPERATORS__"] + common + inc
        else:
            cmd = ["g++", "-fPIC", "-c", src, "-o", obj] + common + inc
        subprocess.run(cmd, check=True)
        objs.append(obj)
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    subprocess.run(["g++", "-shared", "-o", OUT] + objs + [f"-L{libdir}", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python",
                    "-lc10_cuda", "-ltorch_cuda", "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{libdir}"], check=True)
    shutil.rmtree(tmp, ignore_errors=True)
    return OUT


def load():
    """Import the built extension (needs torch imported first)."""
    import importlib.util
    import torch  # noqa: F401
    if not os.path.exists(OUT):
        return None
    spec = importlib.util.spec_from_file_location(NAME, OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
