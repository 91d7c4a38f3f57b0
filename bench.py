#!/usr/bin/env python
"""bench.py -- the headline measurement (BASELINE.json configs[1]):
nn_distance forward + gradient at B=32, N=M=2048 per GPU, in G unordered pairs/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the Chamfer hot path (NnDistance + NnDistanceGrad through
the C ABI) over one batch of synthetic clouds; K steps form a timed window and the
median of several back-to-back windows is reported.  Prints ONE JSON line on rank 0.
See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, N, M = 32, 2048, 2048
METRIC = "nn_distance_fwd_grad_throughput"
UNIT = "Gpairs/s"            # unordered (xyz1 point, xyz2 point) pairs: B*N*M per step
FLOP_PER_PAIR = 16           # algorithmic: 2 directed evaluations x 8 FLOP (tf_nndistance_g.cu:25-28)
RING = 128                   # distinct batches cycled through: 201 MB of inputs, larger than the 126 MB L2


def make_inputs(b, n, m, ring, seed=100):
    """S-randn (mirrors tf_nndistance.py:45-49), `ring` independent batches."""
    rs = np.random.RandomState(seed)
    xyz1 = rs.randn(ring, b, n, 3).astype(np.float32)
    xyz2 = rs.randn(ring, b, m, 3).astype(np.float32)
    return xyz1, xyz2


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [s for (t, s) in self.samples if t0 - 0.05 <= t <= t1 + 0.05] or [s for (_, s) in self.samples]
        mhz, mx, reasons = [], None, set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                mhz.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz)}


# ---------------------------------------------------------------------------
# reference CPU arm / cpu_baseline  (the only places oracle/ is executed here)
# ---------------------------------------------------------------------------
def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _ref_cpu_step(lib, xyz1, xyz2, g1, g2, threads):
    """One fwd+grad pass of the reference's CPU loops over a (b,n,3)/(b,m,3) batch,
    batch elements split over `threads` host threads (the reference itself is single-threaded)."""
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.empty((b, n), np.float32); i1 = np.empty((b, n), np.int32)
    d2 = np.empty((b, m), np.float32); i2 = np.empty((b, m), np.int32)
    o1 = np.empty((b, n, 3), np.float32); o2 = np.empty((b, m, 3), np.float32)
    fp = C.POINTER(C.c_float); ip = C.POINTER(C.c_int)

    errors = []

    def work(lo, hi):
        try:
            _work(int(lo), int(hi))
        except BaseException as e:   # a silently dead thread would fake a fast baseline
            errors.append(e)

    def _work(lo, hi):
        if hi <= lo:
            return
        s = slice(lo, hi)
        args = (hi - lo, n, xyz1[s].ctypes.data_as(fp), m, xyz2[s].ctypes.data_as(fp))
        lib.ref_cpu_nn_distance(*args, d1[s].ctypes.data_as(fp), i1[s].ctypes.data_as(ip),
                                d2[s].ctypes.data_as(fp), i2[s].ctypes.data_as(ip))
        lib.ref_cpu_nn_distance_grad(*args, g1[s].ctypes.data_as(fp), i1[s].ctypes.data_as(ip),
                                     g2[s].ctypes.data_as(fp), i2[s].ctypes.data_as(ip),
                                     o1[s].ctypes.data_as(fp), o2[s].ctypes.data_as(fp))
    if threads <= 1:
        work(0, b)
    else:
        bounds = np.linspace(0, b, threads + 1).astype(int)
        ths = [threading.Thread(target=work, args=(bounds[t], bounds[t + 1])) for t in range(threads)]
        [t.start() for t in ths]
        [t.join() for t in ths]
    if errors:
        raise errors[0]
    return d1, i1, d2, i2, o1, o2


def _load_ref_cpu():
    """oracle/_ref/libref_cpu.so (the reference's own loops) if it is there, else the oracle port."""
    import oracle
    if oracle.ref_cpu.available():
        return oracle.ref_cpu.lib, "reference"
    lib = oracle.cpu.lib

    class Port:   # same entry-point names over the C restatement
        @staticmethod
        def ref_cpu_nn_distance(b, n, x1, m, x2, d1, i1, d2, i2):
            lib.oracle_nn_distance(b, n, x1, m, x2, d1, i1, d2, i2, 0)

        @staticmethod
        def ref_cpu_nn_distance_grad(*a):
            lib.oracle_nn_distance_grad(*a)
    return Port, "port"


def cpu_baseline(threads, reps, sample_b=None):
    lib, kind = _load_ref_cpu()
    b = sample_b or B
    xyz1, xyz2 = make_inputs(b, N, M, 1)
    g1 = np.full((b, N), 100.0 / (b * N), np.float32); g2 = np.full((b, M), 100.0 / (b * M), np.float32)
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        _ref_cpu_step(lib, xyz1[0], xyz2[0], g1, g2, threads)
        best = min(best, time.perf_counter() - t0)
    return {"value": b * N * M / best / 1e9, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "fwd+grad on %d of %d batch elements (N=M=%d), best of %d, %s" % (b, B, N, reps, _cpu_model()),
            "seconds_per_step_sample": best}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    lib, kind = _load_ref_cpu()
    if args.steps > 200:          # the GPU arm's default step count would take minutes on the host
        args.steps, args.warmup = 20, 2
    xyz1, xyz2 = make_inputs(B, N, M, 1)
    g1 = np.full((B, N), 100.0 / (B * N), np.float32); g2 = np.full((B, M), 100.0 / (B * M), np.float32)
    for _ in range(args.warmup):
        _ref_cpu_step(lib, xyz1[0], xyz2[0], g1, g2, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _ref_cpu_step(lib, xyz1[0], xyz2[0], g1, g2, threads)
    dt = time.perf_counter() - t0
    val = B * N * M * args.steps / dt / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (S-randn, seed 100)",
            "config": {"workload": "nn_distance fwd+grad B=%d N=M=%d (BASELINE.json configs[1]) on the host CPU" % (B, N),
                       "note": "reference CPU loops (tf_nndistance.cpp:21-43,126-163); the reference is single-threaded, "
                               "here the batch is split over all host threads"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": "full batch, every step, %s" % _cpu_model()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


# ---------------------------------------------------------------------------
# product arm
# ---------------------------------------------------------------------------
WINDOWS = 7                  # timed windows of K steps each; the reported step time is the median window


def _pin_to_own_cores(local, world):
    """Give every rank its own slice of the host cores: eight ranks submitting from the same cores starve each other
    (round 1's e2e figure scaled 0.29 at 8 GPUs with all ranks on the default affinity mask)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(world, 1))
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return mine
    except (AttributeError, OSError):
        return None


def _traffic_from_profiles():
    """DRAM bytes per launch of the sweep kernel from the committed ncu summary (profiles/): never a literal."""
    import glob
    import re
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_summary.txt")), reverse=True):
        rd = wr = None
        in_kernel = False
        with open(path) as f:
            for line in f:
                if line.startswith("kernel:"):
                    in_kernel = "nn_fwd_kernel" in line
                elif in_kernel:
                    m = re.match(r"\s*dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
                    if m:
                        v = float(m.group(2)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(3), 1)
                        if m.group(1) == "read":
                            rd = v
                        else:
                            wr = v
        if rd is not None and wr is not None:
            return int(rd + wr), os.path.relpath(path, ROOT)
    return None, None


def run_product(args):
    import torch
    import torch.distributed as dist
    from pointnet_autoencoder_b200 import _lib, ops
    from pointnet_autoencoder_b200 import host_api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product arm has no CPU fallback")
    cores = _pin_to_own_cores(local, world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    sms = C.c_int(); maj = C.c_int(); mnr = C.c_int()
    _lib.check(lib.pnae_device_info(C.byref(sms), C.byref(maj), C.byref(mnr)))

    # each rank owns its own B elements (batch sharding, no data-path collective): weak scaling
    h1, h2 = make_inputs(B, N, M, RING, seed=100 + rank)
    x1 = torch.from_numpy(h1).to(dev); x2 = torch.from_numpy(h2).to(dev)
    g1 = torch.full((B, N), 100.0 / (B * N), device=dev); g2 = torch.full((B, M), 100.0 / (B * M), device=dev)
    stream = torch.cuda.current_stream()

    # One CUDA graph per ring slot.  A step is NnDistance + NnDistanceGrad through the C ABI: the two-kernel form
    # (pnae_nn_distance_fwd_grad: sweep, then a finalize that also forms the gradients from the caller's grad_dist arrays)
    # is the benchmarked one; the three-kernel form (fwd, then the separate gradient op) is timed next to it.
    from pointnet_autoencoder_b200.graphs import ChamferStep
    SPG = args.steps_per_graph      # consecutive steps captured per graph (launch overhead amortised over SPG steps)
    assert RING % SPG == 0, "--steps-per-graph must divide the ring of %d batches" % RING
    NG = RING // SPG

    class Ring:
        """graphs of SPG consecutive steps over the ring of input batches, plus one shorter graph for the remainder, so
        that run(k) executes EXACTLY k steps whatever k is"""

        def __init__(self, **kw):
            self.kw = kw
            first = ChamferStep([x1[j] for j in range(SPG)], [x2[j] for j in range(SPG)], g1, g2, **kw)
            self.share = kw.pop("share_buffers_with", None) or first        # every graph writes the same outputs / workspace
            self.kw = dict(kw, share_buffers_with=self.share)
            self.full = [first] + [ChamferStep([x1[g * SPG + j] for j in range(SPG)], [x2[g * SPG + j] for j in range(SPG)], g1, g2, **self.kw)
                                   for g in range(1, NG)]
            self.rem = {}
            self.pos = 0

        def run(self, k):
            for _ in range(k // SPG):
                self.full[self.pos % NG].run(); self.pos += 1
            r = k % SPG
            if r:
                if r not in self.rem:
                    self.rem[r] = ChamferStep([x1[RING - 1 - j] for j in range(r)], [x2[RING - 1 - j] for j in range(r)], g1, g2, **self.kw)
                self.rem[r].run()

    # the benchmarked step: two kernels per step, software-pipelined inside each graph (step s+1's sweep runs while step
    # s's finalize + gradient resolve; alternating output sets and workspaces, pnae_chamfer_graph_create_pipelined)
    ring = Ring(fused=True, pipelined=True)
    ring_seq = Ring(fused=True, share_buffers_with=ring.share)     # the same steps strictly one after the other
    ring3 = Ring(share_buffers_with=ring.share)                    # the three-kernel form, sequential
    ring_fwd = Ring(forward_only=True, pipelined=True, share_buffers_with=ring.share)   # the dominant kernel pair alone (sweep + finalize), for the roofline figure
    ring_fwd_seq = Ring(forward_only=True, share_buffers_with=ring.share)
    for r_ in (ring, ring_seq, ring3, ring_fwd, ring_fwd_seq):     # build the remainder graphs outside the timed region
        r_.run(args.steps % SPG); r_.run(args.warmup % SPG)
    slots = ring.full

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_windows(which, windows):
        """`windows` back-to-back windows of EXACTLY args.steps steps each, CUDA events on the launching stream around
        every window; -> the events"""
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(windows + 1)]
        evs[0].record(stream)
        for w in range(windows):
            which.run(args.steps)
            evs[w + 1].record(stream)
        return evs

    ring.run(args.warmup)
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.15)
    barrier()
    t0 = time.perf_counter()
    ev_step = timed_windows(ring, WINDOWS)
    barrier()
    ev_three = timed_windows(ring3, WINDOWS)
    barrier()
    ev_fwd = timed_windows(ring_fwd, WINDOWS)        # forward alone, same ring, same clocks
    barrier()
    ev_seq = timed_windows(ring_seq, WINDOWS)
    barrier()
    ev_fwd_seq = timed_windows(ring_fwd_seq, WINDOWS)
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    med = lambda evs: float(np.median([evs[i].elapsed_time(evs[i + 1]) for i in range(len(evs) - 1)]))
    win_ms = [ev_step[i].elapsed_time(ev_step[i + 1]) for i in range(WINDOWS)]
    ms = float(np.median(win_ms))
    three_ms = med(ev_three) / args.steps
    fwd_ms = med(ev_fwd) / args.steps
    seq_ms = med(ev_seq) / args.steps
    fwd_seq_ms = med(ev_fwd_seq) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---- the last timed step's results against the oracle (one element; the checker, outside every timed region)
    check = None
    if rank == 0:
        import oracle
        last = (ring.pos - 1) % NG if ring.pos else 0
        slots[last].run(); torch.cuda.synchronize()
        e_in = last * SPG + SPG - 1
        od1, oi1, od2, oi2 = oracle.cpu.nn_distance(h1[e_in, :1], h2[e_in, :1])
        s0 = slots[0]
        ok = (np.array_equal(s0.dist1[:1].cpu().numpy(), od1) and np.array_equal(s0.idx1[:1].cpu().numpy(), oi1)
              and np.array_equal(s0.dist2[:1].cpu().numpy(), od2) and np.array_equal(s0.idx2[:1].cpu().numpy(), oi2))
        og1, og2 = oracle.cpu.nn_distance_grad(h1[e_in, :1], h2[e_in, :1], np.full((1, N), 100.0 / (B * N), np.float32), oi1,
                                               np.full((1, M), 100.0 / (B * M), np.float32), oi2)
        gerr = max(float(np.abs(s0.grad_xyz1[:1].cpu().numpy() - og1).max() / np.abs(og1).max()),
                   float(np.abs(s0.grad_xyz2[:1].cpu().numpy() - og2).max() / np.abs(og2).max()))
        check = {"dist_idx_bit_exact_vs_oracle": bool(ok), "grad_max_rel_err_vs_oracle": gerr, "element": "batch %d, element 0" % e_in}
        if not ok or gerr > 1e-4:
            raise SystemExit("bench.py: the timed step's results do not match the oracle: %r" % (check,))

    # ---- end to end through the public host-buffer API: pinned host in, results back on the host
    e2e_steps = 400
    p1 = torch.from_numpy(h1).pin_memory(); p2 = torch.from_numpy(h2).pin_memory()   # inputs start in pinned host memory

    E2E_WINDOWS = 5
    SUB = 4                     # batches per submission of the host pipeline (one CUDA graph, one copy each way; 8 and 16 measured the same)
    assert e2e_steps % SUB == 0 and RING % SUB == 0

    def e2e_run(results):
        runner = host_api.ChamferHostPipeline(B, N, M, dev, depth=4, results=results, steps_per_submit=SUB)
        chunk = lambda t, i: t[(i * SUB) % RING: (i * SUB) % RING + SUB]      # SUB consecutive batches: contiguous pinned memory
        for i in range(4):
            runner.submit(chunk(p1, i), chunk(p2, i))
        runner.drain()
        # E2E_WINDOWS back-to-back windows of e2e_steps steps each, the median reported (like `value`): one window is
        # 20 ms of host-driven submissions, and a single scheduling hiccup of the submitting thread shows in it
        dts = []
        for _ in range(E2E_WINDOWS):
            barrier()
            t_0 = time.perf_counter()
            got = 0
            for i in range(e2e_steps // SUB):
                got += runner.submit(chunk(p1, i), chunk(p2, i)) is not None
            got += len(runner.drain())                     # every step's results are back on the host when the clock stops
            dts.append(time.perf_counter() - t_0)
            assert got == e2e_steps // SUB
        dt = float(np.median(dts))
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt, runner
    e2e_s, runner = e2e_run("all")
    e2e_grads_s, runner_g = e2e_run("grads")

    def pcie_rate(nbytes, to_host):
        """GB/s of back-to-back pinned copies of one step's result (to_host) / input (to device) size on one stream:
        the link's own ceiling for the e2e figure"""
        dbuf = torch.empty((nbytes,), dtype=torch.uint8, device=dev); hbuf = torch.empty((nbytes,), dtype=torch.uint8, pin_memory=True)
        src, dst = (dbuf, hbuf) if to_host else (hbuf, dbuf)
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b_ = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100):
            dst.copy_(src, non_blocking=True)
        b_.record(); b_.synchronize()
        return nbytes * 100 / (a.elapsed_time(b_) * 1e-3) / 1e9
    pcie = None
    if rank == 0:
        try:
            pcie = {"d2h_gbs": pcie_rate(runner.d2h_bytes, True), "h2d_gbs": pcie_rate(runner.h2d_bytes // 2, False)}
        except Exception as exc:      # noqa: BLE001
            pcie = {"error": str(exc)}

    extra = {}

    def guarded(fn, *a):
        # an `extra` figure must never cost the bench line (every rank runs the same code, so a failure is the same
        # on all of them; the --max-seconds watchdog covers anything else)
        try:
            return fn(*a)
        except Exception as exc:      # noqa: BLE001
            return {"error": "%s: %s" % (type(exc).__name__, exc)}
    if not args.no_emd:
        extra["emd_strong_scaling"] = guarded(emd_numbers, dev, world, rank, sms.value)
    if not args.no_train:
        extra["ae_train"] = guarded(train_numbers, dev, world)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pairs = B * N * M
    value = pairs * args.steps * world / (ms * 1e-3) / 1e9
    ffma_tflops = measure_ffma_rate(lib, torch, dev)
    sm_max = clocks.get("sm_max_mhz") or 1965.0
    fp32_peak = sms.value * 128 * 2 * sm_max * 1e6 / 1e12          # TFLOP/s at the max SM clock
    achieved = FLOP_PER_PAIR * pairs / (fwd_ms * 1e-3) / 1e12
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except (OSError, ValueError):
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg_bytes = 4 * (3 * B * (N + M) + 2 * B * (N + M))               # xyz in, dist+idx out
    traffic, traffic_file = _traffic_from_profiles()
    step_ms = ms / args.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (S-randn, seed 100+rank; mirrors tf_nndistance.py:45-49)",
        "config": {"workload": "nn_distance fwd+grad B=%d N=M=%d per GPU (BASELINE.json configs[1])" % (B, N),
                   "pairs_per_step_per_gpu": pairs, "parallelism": "batch-sharded x%d, no data-path collective" % world,
                   "l2": "inputs cycle through a ring of %d distinct batches (%.0f MB) > 126 MB L2; outputs and workspace are reused" % (RING, RING * 12 * B * (N + M) / 1e6),
                   "launch": "CUDA graphs of %d consecutive steps (2 kernels per step: sweep, finalize+gradient; pnae_nn_distance_fwd_grad), one replay per %d steps, plus one shorter graph when K is not a multiple of %d; inside a graph the steps are software-pipelined: step s+1's sweep runs while step s's finalize resolves, on three rotating output sets and workspaces (every step's results are complete; extra.sequential_step_ms is the same work strictly in order)" % (SPG, SPG, SPG),
                   "timing": "%d back-to-back windows of %d steps, CUDA events on the launching stream; value = median window, max over ranks" % (WINDOWS, args.steps),
                   "upstream_grad": "100/(B*N) (models/model.py:81-83), passed as grad_dist arrays",
                   "host_cores_of_rank0": len(cores) if cores else None},
        "windows": WINDOWS, "window_ms": win_ms, "timed_steps_total": WINDOWS * args.steps,
        "roofline": {"bound": "fp32", "kernel": "nn_distance forward (nn_fwd_kernel sweep + nn_finalize_kernel)", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp32_peak, "traffic": traffic,
                     "peak_source": "%d SMs x 128 lanes x 2 x %.0f MHz (device max SM clock); FFMA microbench reaches 94%% of it (profiles/r1_microbench_b200.txt)" % (sms.value, sm_max),
                     "traffic_source": "dram__bytes_read+write of nn_fwd_kernel per launch, ncu --set full, read from %s; algorithmic bytes %d" % (traffic_file, alg_bytes),
                     # the same launch against the FFMA rate measured in this run (pnae_fp32_probe), None if the probe failed
                     "peak_ffma_measured": ffma_tflops, "frac_of_ffma_measured": (achieved / ffma_tflops) if ffma_tflops else None,
                     "algorithmic_flop_per_launch": FLOP_PER_PAIR * pairs, "kernel_ms": fwd_ms,
                     # the same two kernels with every step's finalize finished before the next sweep starts
                     "kernel_ms_sequential": fwd_seq_ms, "frac_sequential": FLOP_PER_PAIR * pairs / (fwd_seq_ms * 1e-3) / 1e12 / fp32_peak,
                     # the same FLOPs over the whole fwd+grad step (gradient FLOPs counted as 0)
                     "fwd_grad_step_frac": FLOP_PER_PAIR * pairs / (step_ms * 1e-3) / 1e12 / fp32_peak,
                     "fwd_grad_step_frac_sequential": FLOP_PER_PAIR * pairs / (seq_ms * 1e-3) / 1e12 / fp32_peak,
                     "fwd_grad_step_frac_three_kernel_form": FLOP_PER_PAIR * pairs / (three_ms * 1e-3) / 1e12 / fp32_peak,
                     "hbm": {"achieved": alg_bytes / (fwd_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (fwd_ms * 1e-3) / 1e9 / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        "e2e": {"value": pairs * e2e_steps * world / e2e_s / 1e9, "unit": UNIT,
                "h2d_bytes_per_step": runner.h2d_bytes, "d2h_bytes_per_step": runner.d2h_bytes, "steps": e2e_steps, "windows": E2E_WINDOWS,
                "api": "host_api.ChamferHostPipeline over pnae_chamfer_host_pipeline_submit (C): pinned host in -> H2D -> graph (2 kernels per step) -> D2H of dist/idx/grads; %d batches per submission, 4 buffer sets on 3 streams; every step's inputs go in and every step's results come back" % SUB,
                "batches_per_submission": SUB,
                "gradients_only": {"value": pairs * e2e_steps * world / e2e_grads_s / 1e9, "d2h_bytes_per_step": runner_g.d2h_bytes,
                                   "note": "same pipeline returning only grad_xyz1/grad_xyz2 (what a training loop consumes)"},
                # the link's own ceiling, measured in this run on rank 0 alone (copies of exactly these sizes, nothing else running)
                "pcie": pcie,
                "bound": (None if not pcie or "error" in pcie else
                          "device->host copy: %.2f MB per step at the measured %.1f GB/s is %.1f us of the %.1f us e2e step (kernels %.1f us)"
                          % (runner.d2h_bytes / 1e6, pcie["d2h_gbs"], runner.d2h_bytes / pcie["d2h_gbs"] / 1e3, e2e_s / e2e_steps * 1e6, step_ms * 1e3))},
        "gpu_launches": 2 * args.steps * WINDOWS,
        "clocks": clocks,
        "check": check,
    }
    extra["sequential_step_ms"] = seq_ms
    extra["sequential_step_gpairs"] = pairs * world / (seq_ms * 1e-3) / 1e9
    extra["three_kernel_step_ms"] = three_ms
    extra["three_kernel_step_gpairs"] = pairs * world / (three_ms * 1e-3) / 1e9
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(threads=1, reps=3)
    if world == 1 and not args.no_refgpu:
        extra["reference_gpu"] = guarded(reference_gpu_numbers, dev)
    if world == 1:
        extra["encoder_conv_pool"] = guarded(encoder_numbers, dev)
    line["extra"] = extra
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def measure_ffma_rate(lib, torch, dev):
    """TFLOP/s of a pure FFMA stream (8 independent chains per thread, 8 x 256 threads per SM), best of 5, timed with
    CUDA events on the launching stream; None if anything about the probe fails -- it must never cost the bench line."""
    try:
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        out = torch.empty((sms * 8 * 256,), dtype=torch.float32, device=dev)
        flop = C.c_longlong(0)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def _lib_check(rc):
            if rc:
                raise RuntimeError(lib.pnae_last_error().decode())

        _lib_check(lib.pnae_fp32_probe(256, C.c_void_p(out.data_ptr()), out.numel(), C.byref(flop), st))
        torch.cuda.synchronize()
        best = None
        for _ in range(5):
            a = torch.cuda.Event(enable_timing=True); b_ = torch.cuda.Event(enable_timing=True)
            a.record()
            _lib_check(lib.pnae_fp32_probe(16384, C.c_void_p(out.data_ptr()), out.numel(), C.byref(flop), st))
            b_.record(); b_.synchronize()
            t = a.elapsed_time(b_)
            best = t if best is None else min(best, t)
        return flop.value / (best * 1e-3) / 1e12
    except Exception as e:      # noqa: BLE001
        print("fp32 probe failed: %s" % e, file=sys.stderr)
        return None


def _event_times(torch, fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True); b_ = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b_.record(); b_.synchronize()
        ts.append(a.elapsed_time(b_))
    return ts


def emd_numbers(dev, world, rank, sms):
    """BASELINE.json configs[2]: approx_match + match_cost forward+gradient, B=32 in total, batch-sharded over the
    ranks (32/world elements per GPU: STRONG scaling), S-chair clouds, 100 iterations; max over ranks."""
    import torch
    import torch.distributed as dist
    from pointnet_autoencoder_b200 import ops, parallel, synthetic
    label, pred = synthetic.s_chair(B, N)
    lo, hi = parallel.shard_bounds(B, rank, world)
    x1 = torch.from_numpy(np.ascontiguousarray(label[lo:hi])).to(dev); x2 = torch.from_numpy(np.ascontiguousarray(pred[lo:hi])).to(dev)
    fac = ops.approx_match_factors(x1, x2)
    am = float(np.median(_event_times(torch, lambda: ops.approx_match_factors(x1, x2), 100)))
    mc = float(np.median(_event_times(torch, lambda: ops.match_cost_factors(x1, x2, fac), 100)))
    if world > 1:
        t = torch.tensor([am, mc], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        am, mc = float(t[0]), float(t[1])
    peak = sms * 128 * 2 * 1.965e9                    # FP32 FLOP/s of ONE GPU at the max SM clock
    mufu = sms * 16 * 1.965e9                         # MUFU.EX2 per second of one GPU
    pairs = B * N * N                                 # whole job
    return {"approx_match_ms": am, "match_cost_fwd_grad_ms": mc, "emd_fwd_grad_ms": am + mc,
            "elements_per_gpu": hi - lo, "total_elements": B, "iterations": 100, "data": "S-chair", "n_gpus": world,
            "emd_frac_of_fp32_peak_all_gpus": 423 * pairs / ((am + mc) * 1e-3) / (peak * world),
            "roofline": {"bound": "mufu", "kernel": "approx_match_kernel", "unit": "G ex2/s",
                         "achieved": 27 * pairs / (am * 1e-3) / 1e9, "peak": mufu * world / 1e9,
                         "frac": 27 * pairs / (am * 1e-3) / (mufu * world),
                         "note": "27 exponentials per pair (10 levels x 3 sweeps, C_j and A_j+1 fused, last level closed form) against 16 MUFU/clk/SM; floor %.3f ms" % (27 * pairs / (mufu * world) * 1e3)},
            "limit_at_small_shards": "19 grid-wide barriers per call are a fixed cost; at 4 elements per GPU the sweeps between them are ~8x shorter"}


def train_numbers(dev, world):
    """BASELINE.json configs[3]: model_upconv autoencoder training step, Chamfer loss, 32 clouds per GPU (global batch
    32 x world; 256 at 8 GPUs), fp32 library ops for the decoder; one NCCL all-reduce of the gradient bucket per step."""
    from pointnet_autoencoder_b200.train_step import TrainStep
    out = {}
    for name, kw in (("fp32", {}), ("tf32_library_ops", {"tf32": True})):
        ts = TrainStep("upconv", 32, dev, **kw)
        ms, loss = ts.timed(30, 5)
        out[name] = {"samples_per_s": ts.gb / (ms * 1e-3), "ms_per_step": ms, "global_batch": ts.gb, "last_loss": loss,
                     "grad_allreduce_bytes": 4 * ts.nparam if world > 1 else 0}
        del ts
    out["workload"] = "models/model_upconv.py autoencoder, Chamfer loss (fused op), N=2048, 32 clouds per GPU, Adam; 30 steps after 5"
    return out


def reference_gpu_numbers(dev):
    """The same-box incumbent: the reference's own CUDA kernels (tf_nndistance_g.cu, tf_approxmatch_g.cu compiled
    unmodified for sm_100a into oracle/_ref/libref_gpu.so) on the same inputs, outside every timed region of the
    product.  None if the library did not travel to this box."""
    import torch
    import oracle
    from pointnet_autoencoder_b200 import synthetic
    R = oracle.ref_gpu
    if not R.available():
        return None
    lib = R.lib; p = R._p
    h1, h2 = make_inputs(B, N, M, 1)
    x1 = torch.from_numpy(h1[0]).to(dev); x2 = torch.from_numpy(h2[0]).to(dev)
    g1 = torch.full((B, N), 100.0 / (B * N), device=dev); g2 = torch.full((B, M), 100.0 / (B * M), device=dev)
    d1 = torch.empty((B, N), device=dev); i1 = torch.empty((B, N), dtype=torch.int32, device=dev)
    d2 = torch.empty((B, M), device=dev); i2 = torch.empty((B, M), dtype=torch.int32, device=dev)
    o1 = torch.empty((B, N, 3), device=dev); o2 = torch.empty((B, M, 3), device=dev)
    torch.cuda.synchronize()

    def legacy_time(fn, iters, warm=2):
        # the reference launchers use the legacy default stream: bracket with events on that stream
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        legacy = torch.cuda.default_stream(dev)        # torch's default stream IS the legacy default stream
        e0.record(legacy)
        for _ in range(iters):
            fn()
        e1.record(legacy); e1.synchronize()
        return e0.elapsed_time(e1) / iters

    def chamfer():
        lib.ref_gpu_nn_distance(B, N, p(x1), M, p(x2), p(d1), p(i1), p(d2), p(i2))
        lib.ref_gpu_nn_distance_grad(B, N, p(x1), M, p(x2), p(g1), p(i1), p(g2), p(i2), p(o1), p(o2))
    step_ms = legacy_time(chamfer, 50)
    label, pred = synthetic.s_chair(B, N)
    c1 = torch.from_numpy(label).to(dev); c2 = torch.from_numpy(pred).to(dev)
    match = torch.empty((B, N, N), device=dev); temp = torch.empty((B, 4 * N), device=dev); cost = torch.empty((B,), device=dev)
    torch.cuda.synchronize()
    am = legacy_time(lambda: lib.ref_gpu_approxmatch(B, N, N, p(c1), p(c2), p(match), p(temp)), 3, warm=1)
    mc = legacy_time(lambda: (lib.ref_gpu_matchcost(B, N, N, p(c1), p(c2), p(match), p(cost)),
                              lib.ref_gpu_matchcostgrad(B, N, N, p(c1), p(c2), p(match), p(o1), p(o2))), 5, warm=1)
    return {"nn_distance_fwd_grad_ms": step_ms, "nn_distance_fwd_grad_gpairs": B * N * M / (step_ms * 1e-3) / 1e9,
            "approx_match_ms": am, "match_cost_fwd_grad_ms": mc,
            "what": "tf_nndistance_g.cu / tf_approxmatch_g.cu compiled unmodified for sm_100a, same B200, same inputs, legacy default stream"}


def _graph_replay_ms(torch, fn, reps, replays=20):
    """device time of one call of fn, from a CUDA graph of `reps` consecutive calls (no host launch overhead inside)"""
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / (replays * reps)


def encoder_numbers(dev):
    """The encoder at B=32, N=2048: the whole forward with training-mode BatchNorm (seven
    chained kernels) and its dominant layer alone, conv5 (128 -> 1024) + pooling statistics on tcgen05.
    Device time from CUDA-graph replays."""
    import torch
    from pointnet_autoencoder_b200 import ops
    from pointnet_autoencoder_b200.encoder import PointNetEncoder
    b, n, k, c = 32, 2048, 128, 1024
    x = torch.randn(b, n, k, device=dev).to(torch.bfloat16)
    wt = (torch.randn(c, k, device=dev) / k ** 0.5).to(torch.bfloat16)
    ms = _graph_replay_ms(torch, lambda: ops.encoder_conv_pool(x, wt), 10)
    ms_pdl = _graph_replay_ms(torch, lambda: ops.encoder_conv_pool(x, wt, overlap=True), 10)   # the way the encoder chain enqueues it
    flop = 2.0 * b * n * k * c
    out = {"ms": ms, "tflops": flop / (ms * 1e-3) / 1e12, "ms_as_dependent_launch": ms_pdl, "frac_of_measured_bf16_peak": None,
           "note": "conv5 + pooling statistics alone, CUDA-graph replay of 10 launches; bf16 operands, fp32 accumulate"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except (OSError, ValueError):
        peaks = {}
    bf16_peak = peaks.get("bf16_tflops_burst") or peaks.get("bf16_tflops") or 1660.6      # else the figure DESIGN.md quotes
    out["frac_of_measured_bf16_peak"] = out["tflops"] / bf16_peak
    out["bf16_peak_tflops"] = bf16_peak
    enc = PointNetEncoder(fused=True).to(dev).train()
    pc = torch.randn(b, n, 3, device=dev)
    with torch.no_grad():
        out["forward_ms"] = _graph_replay_ms(torch, lambda: enc(pc), 5)
    # with a backward to come conv5 also tracks the arg-extremum point of every (cloud, channel)
    out["forward_for_training_ms"] = _graph_replay_ms(torch, lambda: enc(pc), 5)
    out["forward_note"] = "PointNetEncoder forward, training-mode BatchNorm, B=32 N=2048: mlp_first, 3 x mlp_layer, mlp_apply_bf16, encoder_conv_pool, conv5_finish as programmatic dependent launches"
    return out


_REAL_STDOUT = None


def _emit(line):
    """the ONE JSON line, on the real stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # stdout carries exactly one JSON line: whatever libraries print on fd 1 meanwhile (NCCL's version banner at
    # N>1, for one) is sent to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-emd", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the autoencoder training-step figure (BASELINE configs[3])")
    ap.add_argument("--no-refgpu", action="store_true", help="skip timing the reference's own CUDA kernels")
    ap.add_argument("--steps-per-graph", type=int, default=32, help="consecutive steps captured in one CUDA graph (must divide the ring of 128 batches); a shorter graph covers the remainder, so exactly --steps steps are timed per window")
    ap.add_argument("--max-seconds", type=float, default=1200.0, help="hard wall-clock limit: exit with status 3 instead of hanging a GPU box (0 = none)")
    args = ap.parse_args()
    if args.max_seconds > 0:
        def _abort():
            sys.stderr.write("bench.py: exceeded --max-seconds %.0f, aborting\n" % args.max_seconds)
            sys.stderr.flush()
            os._exit(3)
        wd = threading.Timer(args.max_seconds, _abort)
        wd.daemon = True
        wd.start()
    if args.impl != "reference":
        args.steps = max(1, args.steps)
        args.warmup = max(3, args.warmup)          # never fewer than three warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
