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
    """Streaming form of ChamferHostRunner: `depth` buffer sets and three streams so that the
    host->device copy of step i+1, the kernels of step i and the device->host copy of step i-1
    overlap (PCIe both directions + SMs busy at once).  Every step still moves all of its inputs
    in and all of its results out.

        pipe = ChamferHostPipeline(B, N, M)
        for xyz1, xyz2 in batches:                 # pinned host tensors (or numpy arrays)
            done = pipe.submit(xyz1, xyz2)         # -> results of the step that just retired, or None
        for done in pipe.drain(): ...

    Results are dicts of numpy views on pinned buffers, valid until the next submit()."""

    def __init__(self, b, n, m, device="cuda", depth=4):
        self.device = torch.device(device)
        self.depth = depth
        with torch.cuda.device(self.device):
            self.s_in = torch.cuda.Stream(device=self.device)
            self.s_run = torch.cuda.Stream(device=self.device)
            self.s_out = torch.cuda.Stream(device=self.device)
            self.sets = []
            for _ in range(depth):
                r = ChamferHostRunner(b, n, m, self.device)          # its own inputs, outputs, graph, pinned results
                r.ev_in = torch.cuda.Event(); r.ev_run = torch.cuda.Event(); r.ev_out = torch.cuda.Event()
                r.busy = False
                self.sets.append(r)
        self.h2d_bytes = self.sets[0].h2d_bytes
        self.d2h_bytes = self.sets[0].d2h_bytes
        self.count = 0

    def _retire(self, r):
        r.ev_out.synchronize()
        r.busy = False
        return {k: v.numpy() for k, v in r.h_out.items()}

    def submit(self, xyz1, xyz2):
        r = self.sets[self.count % self.depth]
        assert not r.busy
        with torch.cuda.device(self.device):
            with torch.cuda.stream(self.s_in):
                r.d_xyz1.copy_(r._stage(xyz1, r.h_xyz1), non_blocking=True)
                r.d_xyz2.copy_(r._stage(xyz2, r.h_xyz2), non_blocking=True)
                r.ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(r.ev_in)
                r.graph_step.run()
                r.ev_run.record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(r.ev_run)
                r.h_flat.copy_(r.d_flat, non_blocking=True)
                r.ev_out.record(self.s_out)
        r.busy = True
        self.count += 1
        # free the buffer set the NEXT submit will use: its results stay valid until that submit
        nxt = self.sets[self.count % self.depth]
        return self._retire(nxt) if nxt.busy else None

    def drain(self):
        out = []
        for i in range(self.depth):
            r = self.sets[(self.count + i) % self.depth]
            if r.busy:
                out.append(self._retire(r))
        return out


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
