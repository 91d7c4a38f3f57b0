#!/usr/bin/env python
"""train_bench.py -- train.py-style autoencoder training step (BASELINE.json configs[3] and [0]).

    python train_bench.py [--model upconv|fc|emd] [--steps K] [--warmup W] [--batch 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
           --master-port P train_bench.py --model upconv          # global batch 8 x 32 = 256
    python train_bench.py --cpu-baseline                          # configs[0]: model_cpu-style step on the host

One step = what train.py:193-206 does per batch minus the file I/O: forward (fused tcgen05 encoder,
library decoder), loss through tf_nndistance / tf_approxmatch (the sm_100a kernels), backward, ONE
NCCL all-reduce of the flattened gradient bucket (eager, between the step's two CUDA graphs), Adam update with
train.py's LR / BN-decay schedules (pointnet_autoencoder_b200/train_step.py).
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
        d1, _, d2, _ = models.nn_distance_cpu(pred, x)                    # (B,N,M,3) broadcast, as the TF code does
        loss = (d1.mean() + d2.mean()) * 100
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
    ap.add_argument("--device-pipeline", action="store_true", help="build every batch on the GPU (input_pipeline.DeviceDataset: resample with replacement + y rotation) instead of copying pinned host batches")
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
    from pointnet_autoencoder_b200.train_step import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ts = TrainStep(args.model, args.batch, dev, fused_encoder=not args.unfused_encoder, two_op_loss=args.two_op_loss,
                   use_graph=not args.no_graph, tf32=args.tf32, input="device" if args.device_pipeline else "host")
    ms, lv = ts.timed(args.steps, args.warmup)
    if rank == 0:
        print(json.dumps({"metric": "ae_train_throughput", "value": ts.gb / (ms * 1e-3), "unit": "samples/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "weak", "dtype": "f32 (encoder conv5 operands bf16, fp32 accumulate)",
                          "data": "synthetic (S-chair)", "last_loss": lv,
                          "config": {"workload": "model_%s autoencoder train step, %s loss, N=2048" % (args.model, "EMD" if args.model == "emd" else "Chamfer"),
                                     "global_batch": ts.gb, "per_gpu_batch": args.batch, "params": ts.nparam,
                                     "grad_allreduce_bytes": 4 * ts.nparam if world > 1 else 0,
                                     "parallelism": "dp%d, one eager NCCL all-reduce of the flat gradient bucket per step (between the two CUDA graphs of a step)" % world,
                                     "fused_encoder": not args.unfused_encoder, "cuda_graph_step": not args.no_graph, "library_tf32": bool(args.tf32),
                                     "input": "device pipeline (resample + rotate on the GPU)" if args.device_pipeline else "pinned host batches"}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
