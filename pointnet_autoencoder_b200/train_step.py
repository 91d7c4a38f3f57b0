"""One train.py-style training step (train.py:94-121,180-206) as a reusable object: forward (fused tcgen05 encoder,
library decoder), loss through tf_nndistance / tf_approxmatch, backward, ONE all-reduce of the flat gradient
bucket, Adam with train.py's LR / BN-decay schedules.

CUDA graphs cut the launch overhead of the ~150 small kernels of a step, but the collective stays OUT of the
capture: graph A = forward + backward into the gradient bucket, then an eager NCCL all-reduce, then graph B = the
optimizer update.  (Round 1 captured the all-reduce together with capturable Adam in one graph; that configuration
hung an 8-GPU box once and cannot be bisected cheaply, so the collective is simply never captured.)
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import input_pipeline, models, parallel, synthetic


class TrainStep:
    """step(i) runs training step i on this rank's shard and returns the (device) loss tensor.

    model_name: "upconv" (models/model_upconv.py) | "fc" (models/model.py) | "emd" (models/model_emd.py)
    input:      "host"   -- the batch is copied from pinned host memory every step (train.py feeds numpy batches)
                "device" -- the batch is built on the GPU by input_pipeline.DeviceDataset.batch (resample with
                            replacement + y rotation, part_dataset.py:21-39,118-121), never touching the host
    """

    def __init__(self, model_name="upconv", batch=32, device=None, fused_encoder=True, two_op_loss=False, use_graph=True,
                 tf32=False, input="host", num_point=2048, seed=0):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dev, self.batch, self.n = dev, batch, num_point
        self.use_graph = use_graph
        torch.manual_seed(seed)                                # identical initial replicas
        torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
        torch.backends.cudnn.allow_tf32 = bool(tf32)
        if model_name == "upconv":
            self.model = models.AutoEncoderUpconv(fused_encoder=fused_encoder).to(dev)
        else:
            self.model = models.AutoEncoderFC(num_point=num_point, fused_encoder=fused_encoder).to(dev)
        self.loss_fn = models.emd_loss if model_name == "emd" else (models.chamfer_loss if two_op_loss else models.chamfer_loss_fused)
        self.bucket = parallel.GradBucket(self.model.parameters())
        self.lr_t = torch.tensor(1e-3, device=dev)             # tensor LR: the schedule changes it without re-capturing
        self.opt = torch.optim.Adam(self.model.parameters(), lr=self.lr_t if use_graph else 1e-3, eps=1e-8, capturable=use_graph)
        self.gb = batch * self.world
        self.x = torch.empty((batch, num_point, 3), device=dev)
        self.loss_buf = torch.zeros((), device=dev)
        self.input = input
        # each replica's own shard of the synthetic "dataset"
        label, _ = synthetic.s_chair(batch * 4, num_point, first_id=self.rank * batch * 4)
        if input == "device":
            self.dataset = input_pipeline.DeviceDataset(list(label), npoints=num_point, device=dev)
            self.gen = torch.Generator(device=dev).manual_seed(1234 + self.rank)
        else:
            self.host = torch.from_numpy(label).pin_memory()
        self.fwd_bwd_graphs = {}                               # one captured forward+backward per BN-decay value
        self.opt_graph = None
        self.nparam = sum(p.numel() for p in self.model.parameters())

    # ---- the three phases
    def _forward_backward(self, bn_decay):
        pred, _ = self.model(self.x, bn_decay)
        loss, _ = self.loss_fn(pred, self.x)
        self.bucket.zero()
        loss.backward()
        self.loss_buf.copy_(loss.detach())

    def _all_reduce(self):
        self.bucket.all_reduce()                               # eager, never captured

    def _update(self):
        self.opt.step()

    def _load_batch(self, i):
        if self.input == "device":
            ids = torch.arange((i % 4) * self.batch, (i % 4 + 1) * self.batch, device=self.dev)
            self.x.copy_(self.dataset.batch(ids, generator=self.gen))
        else:
            self.x.copy_(self.host[(i % 4) * self.batch:(i % 4 + 1) * self.batch], non_blocking=True)

    def _capture(self, fn, warm=3):
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(warm):                              # warm-up outside capture (allocator, cuDNN autotune)
                fn()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    def step(self, i):
        self._load_batch(i)
        lr = models.get_learning_rate(i, self.gb)
        bn_decay = models.get_bn_decay(i, self.gb)
        if not self.use_graph:
            for g in self.opt.param_groups:
                g["lr"] = lr
            self._forward_backward(bn_decay)
            self._all_reduce()
            self._update()
            return self.loss_buf
        self.lr_t.fill_(lr)
        if bn_decay not in self.fwd_bwd_graphs:
            self.fwd_bwd_graphs[bn_decay] = self._capture(lambda: self._forward_backward(bn_decay))
            if self.opt_graph is None:
                # capturing the update needs one real (warm-up) update: run it on reduced gradients so replicas stay in
                # lock step, then put weights and optimizer state back so training starts from the initial weights
                self._all_reduce()
                saved = [p.detach().clone() for p in self.bucket.params]
                self.opt_graph = self._capture(self._update, warm=1)
                with torch.no_grad():
                    for p, q in zip(self.bucket.params, saved):
                        p.copy_(q)
                    for st in self.opt.state.values():
                        for v in st.values():
                            if torch.is_tensor(v):
                                v.zero_()
        self.fwd_bwd_graphs[bn_decay].replay()
        self._all_reduce()
        self.opt_graph.replay()
        return self.loss_buf

    def timed(self, steps, warmup):
        """(ms per step on this rank, last loss): CUDA events on the current stream around `steps` steps"""
        for i in range(warmup):
            self.step(i)
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize(self.dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            loss = self.step(warmup + i)
        e1.record()
        lv = float(loss)                                       # device -> host read of the step's result
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize(self.dev)
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, lv
