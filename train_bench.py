#!/usr/bin/env python
"""train_bench.py -- train.py-style autoencoder training step (BASELINE.json configs[3] and [0]).

    python train_bench.py [--model upconv|fc|emd] [--steps K] [--warmup W] [--batch 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
           --master-port P train_bench.py --model upconv          # global batch 8 x 32 = 256
    python train_bench.py --cpu-baseline                          # configs[0]: model_cpu-style step on the host

One step = what train.py:193-206 does per batch minus the file I/O: forward (fused tcgen05 encoder,
library decoder), loss through tf_nndistance / tf_approxmatch (the sm_100a kernels), backward, ONE
NCCL all-reduce of the flattened gradient bucket, Adam update with train.py's LR / BN-decay schedules.
Synthetic S-chair clouds (no dataset ships with the reference).  Prints one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def cpu_baseline(args):
    """configs[0]: models/model_cpu.py -- FC-decoder autoencoder with the pure-TF broadcast Chamfer
    (tf_ops/nn_distance/tf_nndistance_cpu.py:4-25), restated in PyTorch on the host CPU because
    TensorFlow is not installable here.  All host threads."""
    import torch
    import torch.nn.functional as F
    from pointnet_autoencoder_b200 import models, synthetic
    torch.manual_seed(0)
    b, n = args.batch, 2048
    model = models.AutoEncoderFC(num_point=n, fused_encoder=False)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, eps=1e-8)
    label, _ = synthetic.s_chair(b, n)
    x = torch.from_numpy(label)

    def step():
        pred, _ = model(x, 0.5)
        diff = pred[:, :, None, :] - x[:, None, :, :]                     # (B,N,M,3) broadcast, as the TF code does
        d = (diff * diff).sum(-1)
        loss = (d.min(2).values.mean() + d.min(1).values.mean()) * 100
        opt.zero_grad(); loss.backward(); opt.step()
        return float(loss)
    step()
    t0 = time.perf_counter()
    k = max(1, min(args.steps, 3))
    for _ in range(k):
        step()
    dt = (time.perf_counter() - t0) / k
    print(json.dumps({"metric": "ae_train_throughput", "value": b / dt, "unit": "samples/s", "impl": "cpu-restatement",
                      "config": {"workload": "model_cpu.py FC-decoder AE, broadcast Chamfer, B=%d N=%d on the host CPU" % (b, n),
                                 "threads": torch.get_num_threads()}, "ms_per_step": dt * 1e3}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="upconv", choices=["upconv", "fc", "emd"])
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (train.py default 32)")
    ap.add_argument("--unfused-encoder", action="store_true")
    ap.add_argument("--cpu-baseline", action="store_true")
    ap.add_argument("--two-op-loss", action="store_true", help="Chamfer loss through nn_distance + nn_distance_grad instead of the fused entry point")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying one CUDA graph per step")
    ap.add_argument("--tf32", action="store_true", help="let the LIBRARY GEMMs/convs (decoder, encoder layers 1-4) use TF32 tensor cores")
    ap.add_argument("--max-seconds", type=float, default=300.0, help="hard wall-clock limit: the process exits with status 3 instead of hanging a GPU box (0 = none)")
    args = ap.parse_args()
    if args.max_seconds > 0:
        import threading

        def _abort():
            sys.stderr.write("train_bench.py: exceeded --max-seconds %.0f, aborting\n" % args.max_seconds)
            sys.stderr.flush()
            os._exit(3)
        wd = threading.Timer(args.max_seconds, _abort)
        wd.daemon = True
        wd.start()
    if args.cpu_baseline:
        return cpu_baseline(args)

    import torch
    import torch.distributed as dist
    from pointnet_autoencoder_b200 import models, parallel, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)                                   # identical initial replicas
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
    torch.backends.cudnn.allow_tf32 = bool(args.tf32)
    n = 2048
    if args.model == "upconv":
        model = models.AutoEncoderUpconv(fused_encoder=not args.unfused_encoder).to(dev)
    else:
        model = models.AutoEncoderFC(num_point=n, fused_encoder=not args.unfused_encoder).to(dev)
    loss_fn = models.emd_loss if args.model == "emd" else (models.chamfer_loss if args.two_op_loss else models.chamfer_loss_fused)
    bucket = parallel.GradBucket(model.parameters())
    use_graph = not args.no_graph
    lr_t = torch.tensor(1e-3, device=dev)                  # tensor LR: the schedule can change it without re-capturing
    opt = torch.optim.Adam(model.parameters(), lr=lr_t if use_graph else 1e-3, eps=1e-8, capturable=use_graph)
    gb = args.batch * world
    # each replica's own shard of the synthetic "dataset": pinned host clouds, copied in every step
    label, _ = synthetic.s_chair(args.batch * 4, n, first_id=rank * args.batch * 4)
    host = torch.from_numpy(label).pin_memory()
    x = torch.empty((args.batch, n, 3), device=dev)

    loss_buf = torch.zeros((), device=dev)

    def compute(bn_decay):
        pred, _ = model(x, bn_decay)
        loss, pcloss = loss_fn(pred, x)
        bucket.zero()
        loss.backward()
        bucket.all_reduce()
        opt.step()
        loss_buf.copy_(loss.detach())

    graphs = {}                                            # one captured step per BN-decay value (it changes every ~200k samples)

    def step(i):
        x.copy_(host[(i % 4) * args.batch:(i % 4 + 1) * args.batch], non_blocking=True)
        lr = models.get_learning_rate(i, gb)
        bn_decay = models.get_bn_decay(i, gb)
        if not use_graph:
            for g in opt.param_groups:
                g["lr"] = lr
            compute(bn_decay)
            return loss_buf
        lr_t.fill_(lr)
        if bn_decay not in graphs:
            s_ = torch.cuda.Stream(device=dev)
            s_.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s_):
                for _ in range(3):                         # warm-up outside capture (allocator, NCCL, cuDNN autotune)
                    compute(bn_decay)
            torch.cuda.current_stream(dev).wait_stream(s_)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                compute(bn_decay)
            graphs[bn_decay] = g
        graphs[bn_decay].replay()
        return loss_buf

    for i in range(args.warmup):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(args.warmup + i)
    e1.record()
    lv = float(loss)                                       # device -> host read of the step's result
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        nparam = sum(p.numel() for p in model.parameters())
        print(json.dumps({"metric": "ae_train_throughput", "value": gb * args.steps / (ms * 1e-3), "unit": "samples/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                          "higher_is_better": True, "scaling": "weak", "dtype": "f32 (encoder conv5 operands bf16, fp32 accumulate)",
                          "data": "synthetic (S-chair)", "last_loss": lv,
                          "config": {"workload": "model_%s autoencoder train step, %s loss, N=2048" % (args.model, "EMD" if args.model == "emd" else "Chamfer"),
                                     "global_batch": gb, "per_gpu_batch": args.batch, "params": nparam,
                                     "grad_allreduce_bytes": 4 * nparam if world > 1 else 0,
                                     "parallelism": "dp%d, one NCCL all-reduce of the flat gradient bucket per step" % world,
                                     "fused_encoder": not args.unfused_encoder, "cuda_graph_step": use_graph, "library_tf32": bool(args.tf32)}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
