"""BASELINE configs[2] and [4] on G GPUs of one box: the EMD ops at B=32 N=2048 (strong scaling: 32/G elements per
GPU) and the Chamfer + EMD size sweep N=M in {2048, 4096, 8192, 16384} at B=64 (batch-sharded).

    python tools/scaling_sweep.py                                                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
           --master-port P tools/scaling_sweep.py                                      # G = 2, 4, 8

Every rank owns a contiguous batch slice (parallel.shard_bounds); there is no data-path collective.  Times are
CUDA-event times of `iters` back-to-back calls after a warm-up, max over ranks; one JSON line per size on rank 0.
NOT YET RUN on a GPU box (written after round 1's GPU budget was spent).  It aborts itself after
PNAE_MAX_SECONDS (default 600) so a hang cannot burn a multi-GPU box's budget.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pointnet_autoencoder_b200 import ops, parallel, synthetic


def timed(fn, iters, world, dev):
    fn(); fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def main():
    import threading
    limit = float(os.environ.get("PNAE_MAX_SECONDS", "600"))      # never hang a (multi-)GPU box: it is charged per GPU

    def _abort():
        sys.stderr.write("scaling_sweep.py: exceeded %.0f s, aborting\n" % limit)
        sys.stderr.flush()
        os._exit(3)
    wd = threading.Timer(limit, _abort)
    wd.daemon = True
    wd.start()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peak = sms * 128 * 2 * 1.965e9
    cases = [("emd_strong", 32, 2048)] + [("sweep", 64, n) for n in (2048, 4096, 8192, 16384)]
    for name, b, n in cases:
        lo, hi = parallel.shard_bounds(b, rank, world)
        if hi <= lo:
            raise SystemExit("batch %d does not cover %d ranks" % (b, world))
        # every rank generates only its own elements (same seeds as a single-GPU run of the whole batch would slice)
        xyz1, xyz2 = synthetic.s_randn(b, n, n, seed=n)
        x1 = torch.from_numpy(xyz1[lo:hi]).to(dev); x2 = torch.from_numpy(xyz2[lo:hi]).to(dev)
        del xyz1, xyz2
        out = {"case": name, "B": b, "N": n, "n_gpus": world, "elements_per_gpu": hi - lo}
        pairs = b * n * n
        if name == "sweep":
            d1, i1, d2, i2 = ops.nn_distance_fwd(x1, x2)
            g1 = torch.full_like(d1, 100.0 / (b * n)); g2 = torch.full_like(d2, 100.0 / (b * n))
            it = 20 if n <= 4096 else 5
            tf = timed(lambda: ops.nn_distance_fwd(x1, x2), it, world, dev)
            tb = timed(lambda: ops.nn_distance_bwd(x1, x2, g1, i1, g2, i2), it, world, dev)
            out.update(chamfer_fwd_ms=tf, chamfer_bwd_ms=tb, chamfer_gpairs_s=pairs / ((tf + tb) * 1e-3) / 1e9,
                       chamfer_fwd_frac_fp32_peak=16 * pairs / (tf * 1e-3) / (peak * world))
        fac = ops.approx_match_factors(x1, x2)
        it = 5 if n <= 4096 else 2
        ta = timed(lambda: ops.approx_match_factors(x1, x2), it, world, dev)
        tc = timed(lambda: ops.match_cost_factors(x1, x2, fac), it, world, dev)
        out.update(approx_match_ms=ta, match_cost_fwd_grad_ms=tc,
                   emd_frac_fp32_peak=423 * pairs / ((ta + tc) * 1e-3) / (peak * world))
        if rank == 0:
            print(json.dumps(out), flush=True)
        del x1, x2, fac
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
