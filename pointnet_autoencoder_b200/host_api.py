"""Host-buffer entry points: numpy / pinned host arrays in, numpy out.

This is the call a user without device tensors makes (and what bench.py's `e2e`
figure times): every step copies its inputs host->device, runs the kernels through
the C ABI and copies the results device->host.  Buffers (pinned staging, device
inputs/outputs) are allocated once per runner and reused.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .graphs import ChamferStep


def _pinned(shape, dtype):
    return torch.empty(shape, dtype=dtype, pin_memory=True)


class ChamferHostRunner:
    """NnDistance + NnDistanceGrad for fixed (B, N, M) with host buffers.

    step(xyz1, xyz2[, grad_dist1, grad_dist2]) -> dict of numpy views on pinned
    result buffers (valid until the next step).  Default upstream gradient is the
    Chamfer loss's constant 100/(B*N) (models/model.py:81-83)."""

    def __init__(self, b, n, m, device="cuda"):
        self.b, self.n, self.m = b, n, m
        self.device = torch.device(device)
        self.h_xyz1 = _pinned((b, n, 3), torch.float32); self.h_xyz2 = _pinned((b, m, 3), torch.float32)
        self.d_xyz1 = torch.empty((b, n, 3), dtype=torch.float32, device=self.device)
        self.d_xyz2 = torch.empty((b, m, 3), dtype=torch.float32, device=self.device)
        self.g1 = torch.full((b, n), 100.0 / (b * n), device=self.device)
        self.g2 = torch.full((b, m), 100.0 / (b * m), device=self.device)
        # all results live in ONE flat device buffer mirrored by ONE pinned host buffer: a step's
        # device->host traffic is a single copy
        spec = [("dist1", (b, n), torch.float32), ("idx1", (b, n), torch.int32), ("dist2", (b, m), torch.float32),
                ("idx2", (b, m), torch.int32), ("grad_xyz1", (b, n, 3), torch.float32), ("grad_xyz2", (b, m, 3), torch.float32)]
        offs, total = {}, 0
        for name, shape, dt in spec:
            offs[name] = total
            total += (int(np.prod(shape)) * 4 + 255) // 256 * 256
        self.d_flat = torch.empty((total,), dtype=torch.uint8, device=self.device)
        self.h_flat = torch.empty((total,), dtype=torch.uint8, pin_memory=True)
        carve = lambda flat, name, shape, dt: flat[offs[name]: offs[name] + int(np.prod(shape)) * 4].view(dt).view(shape)
        self.d_out = {name: carve(self.d_flat, name, shape, dt) for name, shape, dt in spec}
        self.h_out = {name: carve(self.h_flat, name, shape, dt) for name, shape, dt in spec}
        with torch.cuda.device(self.device):
            self.graph_step = ChamferStep(self.d_xyz1, self.d_xyz2, self.g1, self.g2, outputs=self.d_out)   # one launch per step
        self.h2d_bytes = 4 * 3 * b * (n + m)
        self.d2h_bytes = sum(int(np.prod(shape)) * 4 for _, shape, _ in spec)

    def _stage(self, src, pinned):
        if isinstance(src, torch.Tensor):
            if src.is_pinned():
                return src
            pinned.copy_(src)
            return pinned
        pinned.numpy()[...] = src
        return pinned

    def step(self, xyz1, xyz2, grad_dist1=None, grad_dist2=None):
        with torch.cuda.device(self.device):
            self.d_xyz1.copy_(self._stage(xyz1, self.h_xyz1), non_blocking=True)
            self.d_xyz2.copy_(self._stage(xyz2, self.h_xyz2), non_blocking=True)
            if grad_dist1 is not None:
                self.g1.copy_(torch.as_tensor(grad_dist1, dtype=torch.float32), non_blocking=True)
            if grad_dist2 is not None:
                self.g2.copy_(torch.as_tensor(grad_dist2, dtype=torch.float32), non_blocking=True)
            self.graph_step.run()
            self.h_flat.copy_(self.d_flat, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return {k: v.numpy() for k, v in self.h_out.items()}


class ChamferHostPipeline:
    """Streaming form of ChamferHostRunner over the C ABI's pnae_chamfer_host_pipeline_*: `depth` buffer sets and three
    streams so that the host->device copy of submission i+1, the kernels of submission i and the device->host copy of
    submission i-1 overlap (PCIe both directions + SMs busy at once).  One submit() is ONE call into libpnae.so (two input
    copies, one graph launch, one result copy, events); torch only owns the memory.

        pipe = ChamferHostPipeline(B, N, M, steps_per_submit=4)
        for xyz1, xyz2 in chunks_of_4_batches:     # pinned host tensors (or numpy arrays), shape (4, B, N, 3) / (4, B, M, 3)
            done = pipe.submit(xyz1, xyz2)         # -> results of the submission that just retired, or None
        for done in pipe.drain(): ...

    steps_per_submit: consecutive batches per submission (one CUDA graph, no gaps between its steps).  With 1, inputs are
    (B,N,3)/(B,M,3) and results (B,...) arrays; with k > 1 everything carries a leading axis of length k.
    results="all": every step returns dist1/idx1/dist2/idx2/grad_xyz1/grad_xyz2; results="grads": only the two
    gradient fields cross PCIe (what a training loop consumes).  Results are dicts of numpy views on pinned buffers,
    valid until the next submit()."""

    def __init__(self, b, n, m, device="cuda", depth=4, results="all", fused=True, grad_dist1=None, grad_dist2=None, steps_per_submit=1):
        import ctypes as C
        from . import _lib
        assert results in ("all", "grads") and steps_per_submit >= 1
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.depth = depth
        self.b, self.n, self.m, self.k = b, n, m, steps_per_submit
        k = steps_per_submit
        lib = _lib.load()
        spec = [("grad_xyz1", (b, n, 3), torch.float32), ("grad_xyz2", (b, m, 3), torch.float32), ("dist1", (b, n), torch.float32),
                ("idx1", (b, n), torch.int32), ("dist2", (b, m), torch.float32), ("idx2", (b, m), torch.int32)]
        offs, stride = [], 0
        for name, shape, dt in spec:
            offs.append(stride)
            stride += (int(np.prod(shape)) * 4 + 255) // 256 * 256
        grads_end = offs[2]
        self.names = [s_[0] for s_ in spec] if results == "all" else ["grad_xyz1", "grad_xyz2"]
        self.h2d_bytes = 4 * 3 * b * (n + m)                     # per step
        self.d2h_bytes = sum(int(np.prod(shape)) * 4 for name, shape, _ in spec if name in self.names)
        copy_bytes = stride if results == "all" else grads_end
        f32 = dict(dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self.g1 = torch.full((b, n), 100.0 / (b * n), **f32) if grad_dist1 is None else grad_dist1.to(**f32).contiguous()
            self.g2 = torch.full((b, m), 100.0 / (b * m), **f32) if grad_dist2 is None else grad_dist2.to(**f32).contiguous()
            self.d_xyz1 = [torch.empty((k, b, n, 3), **f32) for _ in range(depth)]
            self.d_xyz2 = [torch.empty((k, b, m, 3), **f32) for _ in range(depth)]
            self.d_out = [torch.empty((k * stride,), dtype=torch.uint8, device=self.device) for _ in range(depth)]
            self.h_out = [torch.empty((k * stride,), dtype=torch.uint8, pin_memory=True) for _ in range(depth)]
            self.h_in1 = [_pinned((k, b, n, 3), torch.float32) for _ in range(depth)]      # staging for non-pinned inputs
            self.h_in2 = [_pinned((k, b, m, 3), torch.float32) for _ in range(depth)]
            # room for three workspaces: with several fused steps per submission the library then builds the software-pipelined
            # graph (a step's sweep runs while the previous steps' finalizes resolve)
            wsb = 3 * ((lib.pnae_nn_distance_workspace_bytes(b, n, m) + 255) // 256 * 256)
            self.ws = torch.empty((max(wsb, 1),), dtype=torch.uint8, device=self.device)
            arr = lambda ts: (C.c_void_p * depth)(*[t.data_ptr() for t in ts])
            h = C.c_void_p()
            torch.cuda.synchronize(self.device)
            _lib.check(lib.pnae_chamfer_host_pipeline_create(depth, k, b, n, m, int(bool(fused)), arr(self.d_xyz1), arr(self.d_xyz2),
                                                             arr(self.d_out), arr(self.h_out), (C.c_size_t * 6)(*offs), stride, copy_bytes,
                                                             C.c_void_p(self.g1.data_ptr()), C.c_void_p(self.g2.data_ptr()),
                                                             C.c_void_p(self.ws.data_ptr()), wsb, C.byref(h)))
        self._h, self._lib, self._check, self._C = h, lib, _lib.check, C

        def carve(flat, i, shape, dt):
            blocks = flat.view(k, stride)[:, offs[i]: offs[i] + int(np.prod(shape)) * 4]      # (k, bytes) strided view
            a = blocks.numpy().view(np.float32 if dt == torch.float32 else np.int32).reshape((k,) + tuple(shape))
            return a[0] if k == 1 else a
        self._views = [{name: carve(self.h_out[j], i, shape, dt) for i, (name, shape, dt) in enumerate(spec) if name in self.names}
                       for j in range(depth)]
        self.kernels_per_step = 2 if fused else 3
        self.count = 0
        self._retired = C.c_int(-1)

    def _host_ptr(self, src, stage):
        """address of a host copy of `src` the async copy can read: pinned tensors as they are, anything else staged"""
        if isinstance(src, torch.Tensor):
            if src.is_pinned() and src.is_contiguous() and src.dtype == torch.float32 and src.numel() == stage.numel():
                return src.data_ptr()
            stage.copy_(src.reshape(stage.shape))
        else:
            stage.numpy()[...] = np.asarray(src, np.float32).reshape(tuple(stage.shape))
        return stage.data_ptr()

    def submit(self, xyz1, xyz2):
        j = self.count % self.depth
        C = self._C
        with torch.cuda.device(self.device):
            self._check(self._lib.pnae_chamfer_host_pipeline_submit(self._h, C.c_void_p(self._host_ptr(xyz1, self.h_in1[j])),
                                                                    C.c_void_p(self._host_ptr(xyz2, self.h_in2[j])), C.byref(self._retired)))
        self.count += 1
        r = self._retired.value
        return self._views[r] if r >= 0 else None

    def drain(self):
        C = self._C
        idx = (C.c_int * self.depth)(); cnt = C.c_int(0)
        with torch.cuda.device(self.device):
            self._check(self._lib.pnae_chamfer_host_pipeline_drain(self._h, idx, C.byref(cnt)))
        return [self._views[idx[i]] for i in range(cnt.value)]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.pnae_chamfer_host_pipeline_destroy(h)
            except Exception:
                pass
            self._h = None


def nn_distance_host(xyz1, xyz2, grad_dist1=None, grad_dist2=None):
    """One-shot convenience wrapper: numpy in -> dict of numpy arrays (copies)."""
    xyz1 = np.ascontiguousarray(xyz1, np.float32); xyz2 = np.ascontiguousarray(xyz2, np.float32)
    r = ChamferHostRunner(xyz1.shape[0], xyz1.shape[1], xyz2.shape[1])
    return {k: v.copy() for k, v in r.step(xyz1, xyz2, grad_dist1, grad_dist2).items()}


def emd_host(xyz1, xyz2):
    """approx_match + match_cost + gradient with host buffers: -> cost (B,), grad1, grad2 (numpy)."""
    x1 = torch.from_numpy(np.ascontiguousarray(xyz1, np.float32)).cuda()
    x2 = torch.from_numpy(np.ascontiguousarray(xyz2, np.float32)).cuda()
    fac = ops.approx_match_factors(x1, x2)
    cost, g1, g2 = ops.match_cost_factors(x1, x2, fac)
    return cost.cpu().numpy(), g1.cpu().numpy(), g2.cpu().numpy()
