"""CUDA-graph replay of the launch-bound Chamfer step.

At B=32, N=M=2048 the forward + gradient is three kernels and about 56 us of GPU time;
launching them from Python one call at a time costs more host time than that.  A
`ChamferStep` fixes the buffers (inputs, outputs, workspace) once, has the C library capture
NnDistance + NnDistanceGrad into one CUDA graph (pnae_chamfer_graph_create) and replays it
with a single launch per step.  Results are identical to the eager calls (same kernels, same
arguments).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class ChamferStep:
    """fwd + grad for fixed input tensors `xyz1` (B,N,3), `xyz2` (B,M,3) and upstream gradients.

    After `run()` the results are in .dist1 .idx1 .dist2 .idx2 .grad_xyz1 .grad_xyz2 (static
    tensors, overwritten by every run).  To feed new data, copy into `.xyz1` / `.xyz2`."""

    def __init__(self, xyz1, xyz2, grad_dist1=None, grad_dist2=None, forward_only=False, share_buffers_with=None, outputs=None,
                 fused=False, pipelined=False):
        # pipelined (multi-step graphs): step s+1's sweep runs while step s's finalize / gradient resolve
        # (pnae_chamfer_graph_create_pipelined).  Steps cycle through three output sets and three workspaces; the public
        # attributes (.dist1 ...) are the set the LAST step wrote, `.other` holds the set of the step before it and
        # `.older` the one before that.
        # fused: NnDistance + NnDistanceGrad through pnae_nn_distance_fwd_grad (sweep + finalize, the finalize also
        # forms the gradients) instead of the three-kernel pnae_nn_distance_fwd -> pnae_nn_distance_bwd sequence
        self.fused = bool(fused) and not forward_only
        self.kernels_per_run = 2 if (self.fused or forward_only) else 3
        # xyz1 / xyz2 may be LISTS of equally shaped tensors: the graph then holds that many consecutive steps
        # (one per input pair, all writing the same output buffers) and one run() replays them back to back
        multi1 = list(xyz1) if isinstance(xyz1, (list, tuple)) else [xyz1]
        multi2 = list(xyz2) if isinstance(xyz2, (list, tuple)) else [xyz2]
        assert len(multi1) == len(multi2) and all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() for t in multi1 + multi2)
        self.steps = len(multi1)
        self._inputs = (multi1, multi2)
        xyz1, xyz2 = multi1[0], multi2[0]
        self.xyz1 = xyz1.contiguous()
        self.xyz2 = xyz2.contiguous()
        b, n, _ = self.xyz1.shape
        m = self.xyz2.shape[1]
        dev = self.xyz1.device
        self.device = dev
        f32 = dict(dtype=torch.float32, device=dev); i32 = dict(dtype=torch.int32, device=dev)
        self.g1 = grad_dist1 if grad_dist1 is not None else torch.full((b, n), 100.0 / (b * n), **f32)
        self.g2 = grad_dist2 if grad_dist2 is not None else torch.full((b, m), 100.0 / (b * m), **f32)
        lib = _lib.load()
        o = share_buffers_with          # another ChamferStep of the same shape: reuse its outputs and workspace
        if o is not None:               # (steps that run one after another on one stream, e.g. a ring of input batches)
            assert o.dist1.shape == (b, n) and o.dist2.shape == (b, m) and o.device == dev
            self.dist1, self.idx1, self.dist2, self.idx2 = o.dist1, o.idx1, o.dist2, o.idx2
            self.grad_xyz1, self.grad_xyz2, self.ws = o.grad_xyz1, o.grad_xyz2, o.ws
        elif outputs is not None:       # caller-provided output tensors (e.g. views into one flat buffer)
            self.dist1, self.idx1, self.dist2, self.idx2 = outputs["dist1"], outputs["idx1"], outputs["dist2"], outputs["idx2"]
            self.grad_xyz1, self.grad_xyz2 = outputs["grad_xyz1"], outputs["grad_xyz2"]
            assert all(t.is_contiguous() and t.device == dev for t in outputs.values())
        else:
            self.dist1 = torch.empty((b, n), **f32); self.idx1 = torch.empty((b, n), **i32)
            self.dist2 = torch.empty((b, m), **f32); self.idx2 = torch.empty((b, m), **i32)
            self.grad_xyz1 = torch.empty((b, n, 3), **f32); self.grad_xyz2 = torch.empty((b, m, 3), **f32)
        self.pipelined = bool(pipelined) and self.steps >= 2
        names = ("dist1", "idx1", "dist2", "idx2", "grad_xyz1", "grad_xyz2")
        if self.pipelined:
            # the pipelined form needs the whole batch in one sweep launch (plan[3] = elements per launch): batches too
            # large for that are captured strictly in order
            plan = (C.c_int * 9)()
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            _lib.check(lib.pnae_nn_distance_plan(b, n, m, sms, plan))
            self.pipelined = plan[3] >= b
        with torch.cuda.device(dev):
            wsb = lib.pnae_nn_distance_workspace_bytes(b, n, m)
            if o is None:
                self.ws = torch.empty((max(wsb, 1),), dtype=torch.uint8, device=dev)
            if self.pipelined:
                if o is not None and getattr(o, "other", None) is not None:
                    self.other, self.older = o.other, o.older
                else:
                    self.other = {k: torch.empty_like(getattr(self, k)) for k in names}
                    self.other["ws"] = torch.empty_like(self.ws)
                    self.older = {k: torch.empty_like(v) for k, v in self.other.items()}
            else:
                self.other = self.older = None
            torch.cuda.synchronize(dev)
            h = C.c_void_p()
            p = lambda t: C.c_void_p(t.data_ptr())
            arr1 = (C.c_void_p * self.steps)(*[t.data_ptr() for t in multi1])
            arr2 = (C.c_void_p * self.steps)(*[t.data_ptr() for t in multi2])
            if self.pipelined:
                last = (self.steps - 1) % 3          # the set the last step writes = the public attributes
                def pair(k):
                    mine = getattr(self, k) if k != "ws" else self.ws
                    sets = [None, None, None]
                    sets[last] = mine; sets[(last - 1) % 3] = self.other[k]; sets[(last - 2) % 3] = self.older[k]
                    return (C.c_void_p * 3)(*[t.data_ptr() for t in sets])
                none2 = None
                _lib.check(lib.pnae_chamfer_graph_create_pipelined(
                    int(self.fused), self.steps, 3, 3, b, n, arr1, m, arr2, pair("dist1"), pair("idx1"), pair("dist2"), pair("idx2"),
                    p(self.g1), p(self.g2), none2 if forward_only else pair("grad_xyz1"), none2 if forward_only else pair("grad_xyz2"),
                    pair("ws"), wsb, C.byref(h)))
                self._h = h
                self._lib = lib
                return
            create = lib.pnae_chamfer_graph_create_fused_multi if self.fused else lib.pnae_chamfer_graph_create_multi
            _lib.check(create(self.steps, b, n, arr1, m, arr2, p(self.dist1), p(self.idx1),
                              p(self.dist2), p(self.idx2), p(self.g1), p(self.g2),
                              None if forward_only else p(self.grad_xyz1),
                              None if forward_only else p(self.grad_xyz2), p(self.ws), wsb, C.byref(h)))
        self._h = h
        self._lib = lib

    def run(self):
        _lib.check(self._lib.pnae_graph_launch(self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return self

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.pnae_graph_destroy(h)
            except Exception:
                pass
            self._h = None
