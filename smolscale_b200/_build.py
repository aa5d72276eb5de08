"""Builds libsmolscale_cuda.so in-tree: gcc for the host C file, nvcc (sm_100a) for the CUDA TU.

The shared library is the product: a C-ABI drop-in for the reference's smolscale.h.  It links the
CUDA runtime statically, so it has no dependency besides libc / libpthread / the NVIDIA driver.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_PATH = os.path.join(HERE, "libsmolscale_cuda.so")
PNG_LIB_PATH = os.path.join(HERE, "libsmolpng.so")

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _cuda_home():
    for cand in (os.environ.get("CUDA_HOME"), os.environ.get("CUDA_PATH"), "/usr/local/cuda"):
        if cand and os.path.exists(os.path.join(cand, "bin", "nvcc")):
            return cand
    nvcc = shutil.which("nvcc")
    if nvcc:
        return os.path.dirname(os.path.dirname(os.path.realpath(nvcc)))
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, n) for n in
            ("smolscale-cuda.c", "smolscale-cuda-kernels.cu", "smolscale-cuda-private.h",
             "smolscale-cuda-luts.h")] + \
           [os.path.join(INCLUDE, n) for n in ("smolscale.h", "smolscale-cuda.h")]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    cuda = _cuda_home()
    nvcc = os.path.join(cuda, "bin", "nvcc")
    obj_c = os.path.join(CSRC, "smolscale-cuda.o")
    obj_cu = os.path.join(CSRC, "smolscale-cuda-kernels.o")
    cc = os.environ.get("CC", "gcc")
    cmds = [
        [cc, "-O2", "-g", "-Wall", "-Wextra", "-fPIC", "-fvisibility=hidden", "-pthread",
         "-I", INCLUDE, "-I", CSRC, "-I", os.path.join(cuda, "include"),
         "-c", os.path.join(CSRC, "smolscale-cuda.c"), "-o", obj_c],
        [nvcc, "-std=c++17", "-O3", "-lineinfo"] + NVCC_ARCH + os.environ.get("SMOL_NVCC_FLAGS", "").split() +
        ["-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v" if verbose else "-warn-spills",
         "-I", INCLUDE, "-I", CSRC,
         "-c", os.path.join(CSRC, "smolscale-cuda-kernels.cu"), "-o", obj_cu],
        [nvcc, "-shared", "-cudart", "static"] + NVCC_ARCH +
        ["-Xlinker", "--no-undefined", "-o", LIB_PATH, obj_c, obj_cu, "-lpthread", "-ldl", "-lrt"],
    ]
    for cmd in cmds:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            print(" ".join(cmd))
            print(r.stdout)
            print(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("build failed: " + " ".join(cmd))
    return LIB_PATH


def build_png(force=False):
    """libsmolpng.so: the PNG file I/O either side of the path (include/smol-png.h).  Plain C over
    zlib, no CUDA; kept out of libsmolscale_cuda.so so that library's dependencies do not grow."""
    srcs = [os.path.join(CSRC, "smol-png.c"), os.path.join(INCLUDE, "smol-png.h")]
    if not force and os.path.exists(PNG_LIB_PATH) and \
            all(os.path.getmtime(s) <= os.path.getmtime(PNG_LIB_PATH) for s in srcs):
        return PNG_LIB_PATH
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-g", "-Wall", "-Wextra", "-fPIC", "-fvisibility=hidden", "-shared",
           "-I", INCLUDE, "-o", PNG_LIB_PATH, srcs[0], "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(" ".join(cmd))
        print(r.stdout)
        print(r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return PNG_LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
    print(build_png(force=True))
