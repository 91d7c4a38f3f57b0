"""Build recipe for the test oracle (TEST INFRASTRUCTURE ONLY).

  oracle/liboracle.so        gcc  oracle/oracle.c              -- the C restatement, always built
  oracle/_ref/libref_cpu.so  g++  ref_cpu_shim.cpp + the reference CPU loops cut from /root/reference
  oracle/_ref/libref_gpu.so  nvcc ref_gpu_shim.cu  + the reference .cu files compiled unmodified

The two `_ref` libraries are built only where /root/reference exists (the build
container); the GPU box receives the prebuilt files with the gpurun snapshot.
Reference sources are never copied into the repo: function bodies are cut into
a temporary directory outside it, compiled, and the directory is removed.

Run:  python -m oracle.build      (or  oracle.build.build_all())
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("PNAE_REFERENCE_ROOT", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

# (file relative to the reference root, first line, last line, text the first line must contain)
_CUTS = {
    "ref_nnsearch.inc": ("tf_ops/nn_distance/tf_nndistance.cpp", 21, 43, "static void nnsearch("),
    "ref_nngrad_body.inc": ("tf_ops/nn_distance/tf_nndistance.cpp", 126, 163, "for (int i=0;i<b*n*3;i++)"),
    "ref_approxmatch_cpu.inc": ("tf_ops/approxmatch/tf_approxmatch.cpp", 23, 140, "void approxmatch_cpu("),
}


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))


def _newer(target, *sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_oracle(force=False):
    """The C restatement -> oracle/liboracle.so"""
    src = os.path.join(HERE, "oracle.c")
    out = os.path.join(HERE, "liboracle.so")
    if not force and _newer(out, src):
        return out
    # -fopenmp only parallelises oracle_emd_fp64 (the fp64 ground truth); every fp32 restatement stays a scalar loop
    _run(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden", src, "-o", out, "-lm"])
    return out


def have_reference():
    return os.path.isfile(os.path.join(REF_ROOT, "tf_ops/nn_distance/tf_nndistance_g.cu"))


def build_ref_cpu(force=False):
    """Reference CPU loops, compiled from where they lie -> oracle/_ref/libref_cpu.so"""
    out = os.path.join(REF_OUT, "libref_cpu.so")
    shim = os.path.join(HERE, "ref_cpu_shim.cpp")
    if not have_reference():
        return out if os.path.exists(out) else None
    if not force and _newer(out, shim):
        return out
    os.makedirs(REF_OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="pnae_refcut_")
    try:
        for name, (rel, first, last, must) in _CUTS.items():
            with open(os.path.join(REF_ROOT, rel)) as f:
                lines = f.readlines()
            if must not in lines[first - 1]:
                raise RuntimeError("reference %s:%d does not start with %r" % (rel, first, must))
            with open(os.path.join(tmp, name), "w") as f:
                f.writelines(lines[first - 1:last])
        # -O2 and no -march: the flags a TF custom-op build of the reference would use
        _run(["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-I", tmp, shim, "-o", out])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def build_ref_gpu(force=False):
    """Reference CUDA kernels, compiled unmodified for sm_100a -> oracle/_ref/libref_gpu.so"""
    out = os.path.join(REF_OUT, "libref_gpu.so")
    shim = os.path.join(HERE, "ref_gpu_shim.cu")
    if not have_reference():
        return out if os.path.exists(out) else None
    if not force and _newer(out, shim):
        return out
    if shutil.which("nvcc") is None:
        return None
    os.makedirs(REF_OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="pnae_refgpu_")
    try:
        objs = []
        for rel, extra in (("tf_ops/nn_distance/tf_nndistance_g.cu", ["-DGOOGLE_CUDA=1"]),
                           ("tf_ops/approxmatch/tf_approxmatch_g.cu", [])):
            obj = os.path.join(tmp, os.path.basename(rel) + ".o")
            # the reference's own compile lines are `nvcc -O2 -c ... -x cu -Xcompiler -fPIC`
            # (tf_nndistance_compile.sh:1, tf_approxmatch_compile.sh:4); only the arch is added.
            _run(["nvcc", "-O2", *ARCH, *extra, "-x", "cu", "-Xcompiler", "-fPIC", "-c",
                  os.path.join(REF_ROOT, rel), "-o", obj])
            objs.append(obj)
        shim_o = os.path.join(tmp, "shim.o")
        _run(["nvcc", "-O2", *ARCH, "-Xcompiler", "-fPIC", "-c", shim, "-o", shim_o])
        _run(["nvcc", "-shared", *ARCH, "-o", out, shim_o, *objs])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def build_all(force=False):
    paths = {"oracle": build_oracle(force)}
    paths["ref_cpu"] = build_ref_cpu(force)
    paths["ref_gpu"] = build_ref_gpu(force)
    return paths


if __name__ == "__main__":
    for k, v in build_all(force="--force" in sys.argv).items():
        print("%-8s %s" % (k, v))
