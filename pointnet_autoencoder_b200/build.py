"""Compile the CUDA kernels + C ABI into pointnet_autoencoder_b200/libpnae.so (in-tree).

nvcc cross-compiles sm_100a without a GPU; the .so travels to the GPU box with the
gpurun snapshot (it is git-ignored, not gpurun-ignored).

    python -m pointnet_autoencoder_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libpnae.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=true",   # every rounding-relevant contraction in the kernels is an explicit intrinsic
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-I", INCLUDE,
] + os.environ.get("PNAE_NVCC_DEFS", "").split()   # tuning experiments only, e.g. PNAE_NVCC_DEFS="-DPNAE_NN_COLS=128"


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    # content only (file names relative, no absolute paths): the stamp must stay valid when the tree is copied to
    # another location, e.g. onto the GPU box
    h = hashlib.sha256()
    for p in _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(INCLUDE, "pnae.h"), __file__]:
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode()); h.update(f.read())
    h.update(" ".join(a for a in NVCC_FLAGS if a != INCLUDE).encode())
    return h.hexdigest()


def up_to_date():
    """the in-tree libpnae.so was built from exactly the sources, header and flags that are here now"""
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _digest()


def build(force=False, verbose=False):
    dig = _digest()
    if not force and up_to_date():
        return LIB
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(CSRC, os.path.basename(src)[:-3] + ".o")
        cmd = ["nvcc", *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
        if verbose and out:
            print(out)
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s%s" % (" ".join(cmd), r.stdout, r.stderr))
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
