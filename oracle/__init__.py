"""Test oracle for the reconstruction-loss ops (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package; the product package never does.

  oracle.cpu      numpy-facing wrappers over liboracle.so (the C restatement)
  oracle.ref_cpu  the reference's own CPU loops (oracle/_ref/libref_cpu.so), if built
  oracle.ref_gpu  the reference's own CUDA kernels (oracle/_ref/libref_gpu.so), if built
"""
from . import build  # noqa: F401
from .wrappers import cpu, ref_cpu, ref_gpu, RefGpu, RefCpu, Oracle, NUM_LEVELS, JSTART_GPU  # noqa: F401
