"""Multi-process (world_size 2, gloo, CPU) checks of the batch-sharding plumbing: the N>1 path of
bench.py / the training harness minus the kernels.  The loss op on the shards is played by the CPU
oracle -- the sharding claim under test is "a batch slice of the op == the op on a batch slice"."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointnet_autoencoder_b200 import parallel, synthetic


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world)
        out[rank] = "ok"
    except BaseException as e:   # report instead of hanging the peer
        out[rank] = "%s: %s" % (type(e).__name__, e)
    finally:
        dist.destroy_process_group()


def run2(fn, world=2):
    mgr = mp.get_context("spawn").Manager()      # no fork() of this multi-threaded process
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, out), nprocs=world, join=True)
    assert all(out.get(r) == "ok" for r in range(world)), dict(out)


def test_shard_bounds_cover_the_batch():
    for b in (1, 2, 7, 32, 33):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(b, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _sharded_chamfer(rank, world):
    import oracle
    xyz1, xyz2 = synthetic.s_randn(5, 96, 64, seed=3)            # 5 elements over 2 ranks: uneven shards
    full = oracle.cpu.nn_distance(xyz1, xyz2)
    lo, hi = parallel.shard_bounds(5, rank, world)
    mine = oracle.cpu.nn_distance(xyz1[lo:hi], xyz2[lo:hi])
    for a, f in zip(mine, full):
        got = parallel.all_gather_batch(torch.from_numpy(a), 5)
        assert torch.equal(got, torch.from_numpy(f))


def test_sharded_op_equals_unsharded():
    run2(_sharded_chamfer)


def _grad_bucket(rank, world):
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    ref.load_state_dict(model.state_dict())
    x = torch.randn(8, 6)
    bucket = parallel.GradBucket(model.parameters())
    bucket.zero()
    xs = parallel.shard(x, rank, world)
    model(xs).square().mean().backward()            # mean over the LOCAL batch, like each replica's loss
    bucket.all_reduce()
    ref(x).square().mean().backward()               # mean over the global batch
    for p, q in zip(model.parameters(), ref.parameters()):
        assert p.grad.data_ptr() >= bucket.flat.data_ptr()            # still views into the bucket
        assert torch.allclose(p.grad, q.grad, atol=1e-6)


def test_grad_bucket_allreduce_matches_global_batch():
    run2(_grad_bucket)
