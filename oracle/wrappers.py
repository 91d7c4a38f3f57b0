"""ctypes wrappers over the oracle libraries (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

JSTART_GPU = 7          # tf_approxmatch_g.cu:21  (10 levels)
JSTART_CPU = 8          # tf_approxmatch.cpp:31   (11 levels)
NUM_LEVELS = JSTART_GPU + 3

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


def _check_pair(xyz1, xyz2):
    xyz1 = np.ascontiguousarray(xyz1, dtype=np.float32)
    xyz2 = np.ascontiguousarray(xyz2, dtype=np.float32)
    assert xyz1.ndim == 3 and xyz2.ndim == 3 and xyz1.shape[2] == 3 and xyz2.shape[2] == 3
    assert xyz1.shape[0] == xyz2.shape[0]
    return xyz1, xyz2, xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]


class Oracle:
    """The C restatement (oracle/oracle.c)."""

    def __init__(self):
        self._lib = None

    @property
    def lib(self):
        if self._lib is None:
            self._lib = C.CDLL(_build.build_oracle())
        return self._lib

    def nn_distance(self, xyz1, xyz2, contract=True):
        """-> dist1 (B,N) f32, idx1 (B,N) i32, dist2 (B,M) f32, idx2 (B,M) i32"""
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        d1 = np.empty((b, n), np.float32); i1 = np.empty((b, n), np.int32)
        d2 = np.empty((b, m), np.float32); i2 = np.empty((b, m), np.int32)
        self.lib.oracle_nn_distance(b, n, xyz1.ctypes.data_as(_f32p), m, xyz2.ctypes.data_as(_f32p),
                                    d1.ctypes.data_as(_f32p), i1.ctypes.data_as(_i32p),
                                    d2.ctypes.data_as(_f32p), i2.ctypes.data_as(_i32p), int(bool(contract)))
        return d1, i1, d2, i2

    def nn_distance_grad(self, xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        g1, g1p = _f(grad_dist1); g2, g2p = _f(grad_dist2)
        i1, i1p = _i(idx1); i2, i2p = _i(idx2)
        assert g1.shape == (b, n) and i1.shape == (b, n) and g2.shape == (b, m) and i2.shape == (b, m)
        o1 = np.empty((b, n, 3), np.float32); o2 = np.empty((b, m, 3), np.float32)
        self.lib.oracle_nn_distance_grad(b, n, xyz1.ctypes.data_as(_f32p), m, xyz2.ctypes.data_as(_f32p),
                                         g1p, i1p, g2p, i2p, o1.ctypes.data_as(_f32p), o2.ctypes.data_as(_f32p))
        return o1, o2

    def approx_match(self, xyz1, xyz2, dense=True, factors=False, jstart=JSTART_GPU):
        """GPU schedule.  -> match (B,M,N) and/or factors (B,nlev,N+M)"""
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        nlev = jstart + 3
        mt = np.empty((b, m, n), np.float32) if dense else None
        fc = np.empty((b, nlev, n + m), np.float32) if factors else None
        self.lib.oracle_approxmatch(b, n, m, xyz1.ctypes.data_as(_f32p), xyz2.ctypes.data_as(_f32p),
                                    mt.ctypes.data_as(_f32p) if dense else None,
                                    fc.ctypes.data_as(_f32p) if factors else None, jstart)
        if dense and factors:
            return mt, fc
        return mt if dense else fc

    def approx_match_order(self, xyz1, xyz2, chunk, jstart=JSTART_GPU):
        """factors of the same fp32 schedule with every per-point sum formed in blocks of `chunk` streamed points
        (0: one sequential sum): shows what a change of summation order alone does on a given cloud"""
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        fc = np.empty((b, jstart + 3, n + m), np.float32)
        self.lib.oracle_approxmatch_order(b, n, m, xyz1.ctypes.data_as(_f32p), xyz2.ctypes.data_as(_f32p),
                                          fc.ctypes.data_as(_f32p), jstart, int(chunk))
        return fc

    def match_from_factors(self, xyz1, xyz2, factors, jstart=JSTART_GPU):
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        fc, fcp = _f(factors)
        assert fc.shape == (b, jstart + 3, n + m)
        mt = np.empty((b, m, n), np.float32)
        self.lib.oracle_match_from_factors(b, n, m, xyz1.ctypes.data_as(_f32p), xyz2.ctypes.data_as(_f32p),
                                           fcp, jstart, mt.ctypes.data_as(_f32p))
        return mt

    def match_cost(self, xyz1, xyz2, match):
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        mt, mtp = _f(match)
        assert mt.shape == (b, m, n)
        cost = np.empty((b,), np.float32)
        self.lib.oracle_matchcost(b, n, m, xyz1.ctypes.data_as(_f32p), xyz2.ctypes.data_as(_f32p), mtp,
                                  cost.ctypes.data_as(_f32p))
        return cost

    def match_cost_grad(self, xyz1, xyz2, match):
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        mt, mtp = _f(match)
        assert mt.shape == (b, m, n)
        g1 = np.empty((b, n, 3), np.float32); g2 = np.empty((b, m, 3), np.float32)
        self.lib.oracle_matchcostgrad(b, n, m, xyz1.ctypes.data_as(_f32p), xyz2.ctypes.data_as(_f32p), mtp,
                                      g1.ctypes.data_as(_f32p), g2.ctypes.data_as(_f32p))
        return g1, g2

    def match_cost_factors(self, xyz1, xyz2, factors, jstart=JSTART_GPU):
        """-> cost (B,), grad1 (B,N,3), grad2 (B,M,3) straight from the factors"""
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        fc, fcp = _f(factors)
        assert fc.shape == (b, jstart + 3, n + m)
        cost = np.empty((b,), np.float32)
        g1 = np.empty((b, n, 3), np.float32); g2 = np.empty((b, m, 3), np.float32)
        self.lib.oracle_matchcost_factors(b, n, m, xyz1.ctypes.data_as(_f32p), xyz2.ctypes.data_as(_f32p),
                                          fcp, jstart, cost.ctypes.data_as(_f32p),
                                          g1.ctypes.data_as(_f32p), g2.ctypes.data_as(_f32p))
        return cost, g1, g2


    def emd_fp64(self, xyz1, xyz2, jstart=JSTART_GPU, dense=False):
        """fp64 ground truth of approx_match -> match_cost -> match_cost_grad: cost (B,), grad1, grad2 (float64)
        [, the dense match (B,M,N) float64 if dense]"""
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        cost = np.empty((b,), np.float64)
        g1 = np.empty((b, n, 3), np.float64); g2 = np.empty((b, m, 3), np.float64)
        nlev = jstart + 3
        fac = np.empty((b, nlev, n + m), np.float64) if dense else None
        dp = C.POINTER(C.c_double)
        self.lib.oracle_emd_fp64(b, n, m, xyz1.ctypes.data_as(_f32p), xyz2.ctypes.data_as(_f32p), jstart,
                                 cost.ctypes.data_as(dp), g1.ctypes.data_as(dp), g2.ctypes.data_as(dp),
                                 fac.ctypes.data_as(dp) if dense else None)
        if not dense:
            return cost, g1, g2
        levels = np.array([0.0 if j == -2 else -(4.0 ** j) for j in range(jstart, -3, -1)])
        mt = np.zeros((b, m, n), np.float64)
        for i in range(b):
            a = xyz1[i].astype(np.float64); c = xyz2[i].astype(np.float64)
            d = ((c[:, None, :] - a[None, :, :]) ** 2).sum(-1)            # (m, n)
            for t in range(nlev):
                mt[i] += np.exp(levels[t] * d) * fac[i, t, None, :n] * fac[i, t, n:, None]
        return cost, g1, g2, mt


class RefCpu:
    """The reference's own CPU loops (libref_cpu.so).  `match` crosses this
    boundary in the REFERENCE GPU layout (B,M,N); the (B,N,M) layout the CPU
    functions use internally is handled here by transposing."""

    def __init__(self):
        self._lib = None

    def available(self):
        p = os.path.join(_build.REF_OUT, "libref_cpu.so")
        return os.path.exists(p) or _build.have_reference()

    @property
    def lib(self):
        if self._lib is None:
            p = _build.build_ref_cpu()
            if p is None or not os.path.exists(p):
                raise RuntimeError("oracle/_ref/libref_cpu.so is not built and /root/reference is absent")
            self._lib = C.CDLL(p)
        return self._lib

    def nn_distance(self, xyz1, xyz2):
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        d1 = np.empty((b, n), np.float32); i1 = np.empty((b, n), np.int32)
        d2 = np.empty((b, m), np.float32); i2 = np.empty((b, m), np.int32)
        self.lib.ref_cpu_nn_distance(b, n, xyz1.ctypes.data_as(_f32p), m, xyz2.ctypes.data_as(_f32p),
                                     d1.ctypes.data_as(_f32p), i1.ctypes.data_as(_i32p),
                                     d2.ctypes.data_as(_f32p), i2.ctypes.data_as(_i32p))
        return d1, i1, d2, i2

    def nn_distance_grad(self, xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        g1, g1p = _f(grad_dist1); g2, g2p = _f(grad_dist2)
        i1, i1p = _i(idx1); i2, i2p = _i(idx2)
        o1 = np.empty((b, n, 3), np.float32); o2 = np.empty((b, m, 3), np.float32)
        self.lib.ref_cpu_nn_distance_grad(b, n, xyz1.ctypes.data_as(_f32p), m, xyz2.ctypes.data_as(_f32p),
                                          g1p, i1p, g2p, i2p, o1.ctypes.data_as(_f32p), o2.ctypes.data_as(_f32p))
        return o1, o2

    def approx_match(self, xyz1, xyz2):
        """11-level CPU schedule; returned transposed to (B,M,N)."""
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        mt = np.empty((b, n, m), np.float32)
        self.lib.ref_cpu_approxmatch(b, n, m, xyz1.ctypes.data_as(_f32p), xyz2.ctypes.data_as(_f32p),
                                     mt.ctypes.data_as(_f32p))
        return np.ascontiguousarray(mt.transpose(0, 2, 1))

    def match_cost(self, xyz1, xyz2, match):
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        mt = np.ascontiguousarray(np.asarray(match, np.float32).transpose(0, 2, 1))
        cost = np.empty((b,), np.float32)
        self.lib.ref_cpu_matchcost(b, n, m, xyz1.ctypes.data_as(_f32p), xyz2.ctypes.data_as(_f32p),
                                   mt.ctypes.data_as(_f32p), cost.ctypes.data_as(_f32p))
        return cost

    def match_cost_grad(self, xyz1, xyz2, match):
        xyz1, xyz2, b, n, m = _check_pair(xyz1, xyz2)
        mt = np.ascontiguousarray(np.asarray(match, np.float32).transpose(0, 2, 1))
        g1 = np.empty((b, n, 3), np.float32); g2 = np.empty((b, m, 3), np.float32)
        self.lib.ref_cpu_matchcostgrad(b, n, m, xyz1.ctypes.data_as(_f32p), xyz2.ctypes.data_as(_f32p),
                                       mt.ctypes.data_as(_f32p), g1.ctypes.data_as(_f32p), g2.ctypes.data_as(_f32p))
        return g1, g2

    # raw entry points for timing (no transposes in the timed region)
    def raw(self):
        return self.lib


class RefGpu:
    """The reference's own CUDA kernels (libref_gpu.so) over torch CUDA tensors."""

    def __init__(self):
        self._lib = None

    def available(self):
        return os.path.exists(os.path.join(_build.REF_OUT, "libref_gpu.so"))

    @property
    def lib(self):
        if self._lib is None:
            p = os.path.join(_build.REF_OUT, "libref_gpu.so")
            if not os.path.exists(p):
                p = _build.build_ref_gpu()
            if p is None or not os.path.exists(p):
                raise RuntimeError("oracle/_ref/libref_gpu.so is not built and /root/reference is absent")
            self._lib = C.CDLL(p)
        return self._lib

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError("reference %s launch failed: cuda error %d" % (what, rc))

    def nn_distance(self, xyz1, xyz2):
        import torch
        b, n, _ = xyz1.shape; m = xyz2.shape[1]
        xyz1 = xyz1.contiguous(); xyz2 = xyz2.contiguous()
        d1 = torch.empty((b, n), dtype=torch.float32, device=xyz1.device); i1 = torch.empty((b, n), dtype=torch.int32, device=xyz1.device)
        d2 = torch.empty((b, m), dtype=torch.float32, device=xyz1.device); i2 = torch.empty((b, m), dtype=torch.int32, device=xyz1.device)
        torch.cuda.synchronize()
        self._chk(self.lib.ref_gpu_nn_distance(b, n, self._p(xyz1), m, self._p(xyz2), self._p(d1), self._p(i1),
                                               self._p(d2), self._p(i2)), "nn_distance")
        self._chk(self.lib.ref_gpu_sync(), "sync")
        return d1, i1, d2, i2

    def nn_distance_grad(self, xyz1, xyz2, g1, i1, g2, i2):
        import torch
        b, n, _ = xyz1.shape; m = xyz2.shape[1]
        xyz1 = xyz1.contiguous(); xyz2 = xyz2.contiguous()
        g1 = g1.contiguous(); g2 = g2.contiguous(); i1 = i1.contiguous(); i2 = i2.contiguous()
        o1 = torch.empty((b, n, 3), dtype=torch.float32, device=xyz1.device)
        o2 = torch.empty((b, m, 3), dtype=torch.float32, device=xyz1.device)
        torch.cuda.synchronize()
        self._chk(self.lib.ref_gpu_nn_distance_grad(b, n, self._p(xyz1), m, self._p(xyz2), self._p(g1), self._p(i1),
                                                    self._p(g2), self._p(i2), self._p(o1), self._p(o2)), "nn_distance_grad")
        self._chk(self.lib.ref_gpu_sync(), "sync")
        return o1, o2

    def approx_match(self, xyz1, xyz2):
        import torch
        b, n, _ = xyz1.shape; m = xyz2.shape[1]
        assert b * n * m < 2 ** 31, "reference kernel indexes match with 32-bit ints (SURVEY section 7)"
        xyz1 = xyz1.contiguous(); xyz2 = xyz2.contiguous()
        match = torch.empty((b, m, n), dtype=torch.float32, device=xyz1.device)
        temp = torch.empty((max(b, 32), 2 * (n + m)), dtype=torch.float32, device=xyz1.device)
        torch.cuda.synchronize()
        self._chk(self.lib.ref_gpu_approxmatch(b, n, m, self._p(xyz1), self._p(xyz2), self._p(match), self._p(temp)), "approxmatch")
        self._chk(self.lib.ref_gpu_sync(), "sync")
        return match

    def match_cost(self, xyz1, xyz2, match):
        import torch
        b, n, _ = xyz1.shape; m = xyz2.shape[1]
        xyz1 = xyz1.contiguous(); xyz2 = xyz2.contiguous(); match = match.contiguous()
        cost = torch.empty((b,), dtype=torch.float32, device=xyz1.device)
        torch.cuda.synchronize()
        self._chk(self.lib.ref_gpu_matchcost(b, n, m, self._p(xyz1), self._p(xyz2), self._p(match), self._p(cost)), "matchcost")
        self._chk(self.lib.ref_gpu_sync(), "sync")
        return cost

    def match_cost_grad(self, xyz1, xyz2, match):
        import torch
        b, n, _ = xyz1.shape; m = xyz2.shape[1]
        xyz1 = xyz1.contiguous(); xyz2 = xyz2.contiguous(); match = match.contiguous()
        g1 = torch.empty((b, n, 3), dtype=torch.float32, device=xyz1.device)
        g2 = torch.empty((b, m, 3), dtype=torch.float32, device=xyz1.device)
        torch.cuda.synchronize()
        self._chk(self.lib.ref_gpu_matchcostgrad(b, n, m, self._p(xyz1), self._p(xyz2), self._p(match), self._p(g1), self._p(g2)), "matchcostgrad")
        self._chk(self.lib.ref_gpu_sync(), "sync")
        return g1, g2

    def levels(self):
        import torch
        out = torch.empty((10,), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        self._chk(self.lib.ref_gpu_levels(self._p(out)), "levels")
        self._chk(self.lib.ref_gpu_sync(), "sync")
        return out.cpu().numpy()


cpu = Oracle()
ref_cpu = RefCpu()
ref_gpu = RefGpu()
