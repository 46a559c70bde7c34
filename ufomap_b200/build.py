"""Builds ufomap_b200/libufomap_b200.so (hand-written CUDA for sm_100a) in-tree.

    python -m ufomap_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU
box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libufomap_b200.so")
import glob

SOURCES = [os.path.join(CSRC, "ufo_map.cu")]
# every header the translation unit can include (ufo_export.cuh was missing here)
HEADERS = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [
    os.path.join(os.path.dirname(HERE), "include", "ufomap_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # the DDA must not be FMA-contracted (bit-exact parity with the reference)
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall",
    "-shared", "-cudart", "static",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS + [__file__])


def build(force=False, verbose=False, out=None, defines=()):
    """out/defines build an experimental variant (e.g. out='/tmp/v.so', defines=['UFO_QUEUE=8'])."""
    target = out or LIB
    if not out and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = ([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines]
           + ["-o", target] + SOURCES)
    subprocess.check_call(cmd)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=outs[0] if outs else None,
                defines=defs))
