"""The C-ABI library loads without a GPU and exports exactly what include/pnae.h
declares; the host-side checks (no compute) behave like the reference's OP_REQUIRES."""
import ctypes as C
import os
import re

import pytest
import torch

from pointnet_autoencoder_b200 import _lib, ops
from pointnet_autoencoder_b200.tf_ops.approxmatch import tf_approxmatch
from pointnet_autoencoder_b200.tf_ops.nn_distance import tf_nndistance

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    with open(os.path.join(ROOT, "include", "pnae.h")) as f:
        src = f.read()
    return sorted(set(re.findall(r"PNAE_API[^;(]*?\b(pnae_\w+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("pnae_nn_distance_fwd", "pnae_nn_distance_bwd", "pnae_approx_match", "pnae_match_cost_fwd",
                 "pnae_match_cost_bwd", "pnae_match_cost_factors", "pnae_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH) if os.path.exists(_lib.LIB_PATH) else _lib.load()
    for name in declared_functions():
        assert hasattr(lib, name), "libpnae.so lacks %s declared in include/pnae.h" % name


def test_python_binding_covers_the_header():
    assert sorted(_lib.SIGNATURES) == declared_functions()
    lib = _lib.load()
    assert lib.pnae_version() == 100


def test_argument_validation_without_compute():
    lib = _lib.load()
    # invalid sizes are rejected before anything touches the device
    rc = lib.pnae_nn_distance_fwd(1, 0, None, 4, None, None, None, None, None, None, 0, None)
    assert rc == _lib.PNAE_ERR_INVALID_ARG and b"n>=1" in lib.pnae_last_error()
    rc = lib.pnae_approx_match(1, 4, 4, None, None, None, None, None, 0, None)
    assert rc == _lib.PNAE_ERR_INVALID_ARG
    assert lib.pnae_approx_match_workspace_bytes(2, 10, 6) >= 2 * 16 * 4
    # b == 0 is a no-op, as in the reference (empty batch loops)
    dummy = C.c_void_p(16)
    assert lib.pnae_nn_distance_fwd(0, 4, dummy, 4, dummy, dummy, dummy, dummy, dummy, None, 0, None) == 0


def test_encoder_flags_and_pipelined_graph_arguments_are_checked_before_compute():
    """The round-2 entry points reject unknown flag bits and malformed pointer lists without touching a device."""
    lib = _lib.load()
    d = C.c_void_p(256)                      # never dereferenced: every call below fails its argument checks first
    bad = 8                                  # not PNAE_STATS_ZEROED (1) | PNAE_OVERLAP_PREVIOUS (2)
    assert ops.STATS_ZEROED == 1 and ops.OVERLAP_PREVIOUS == 2
    rc = lib.pnae_mlp_first(64, d, d, d, d, d, bad, None)
    assert rc == _lib.PNAE_ERR_INVALID_ARG and b"flag" in lib.pnae_last_error()
    rc = lib.pnae_mlp_layer(64, 64, 64, d, d, d, d, d, d, 1e-3, 0.9, 1, d, d, d, d, bad, None)
    assert rc == _lib.PNAE_ERR_INVALID_ARG and b"flag" in lib.pnae_last_error()
    rc = lib.pnae_mlp_layer(64, 64, 96, d, d, d, d, d, d, 1e-3, 0.9, 1, d, d, d, d, 0, None)
    assert rc == _lib.PNAE_ERR_INVALID_ARG and b"64 or 128" in lib.pnae_last_error()
    rc = lib.pnae_mlp_layer_xyz(64, d, d, d, d, d, d, d, d, 1e-3, 0.9, 1, 192, d, d, d, d, 0, None)
    assert rc == _lib.PNAE_ERR_INVALID_ARG and b"64 or 128" in lib.pnae_last_error()
    rc = lib.pnae_mlp_layer_xyz(64, d, None, d, d, d, d, d, d, 1e-3, 0.9, 1, 64, d, d, d, d, 0, None)     # training needs the moments
    assert rc == _lib.PNAE_ERR_INVALID_ARG
    rc = lib.pnae_xyz_moments(0, d, d, 0, None)
    assert rc == _lib.PNAE_ERR_INVALID_ARG
    rc = lib.pnae_mlp_apply_bf16(64, 128, d, d, d, d, d, d, 1e-3, 0.9, 1, d, 1, None)                       # STATS_ZEROED means nothing here
    assert rc == _lib.PNAE_ERR_INVALID_ARG and b"flag" in lib.pnae_last_error()
    rc = lib.pnae_encoder_conv_pool(2, 256, 128, 1024, d, d, d, d, d, d, None, None, 4, None)
    assert rc == _lib.PNAE_ERR_INVALID_ARG and b"flag" in lib.pnae_last_error()
    rc = lib.pnae_conv5_finish(2, 1024, 512.0, d, d, d, d, d, d, d, d, d, 1e-3, 0.9, 1, d, d, d, d, d, 5, None)
    assert rc == _lib.PNAE_ERR_INVALID_ARG and b"flag" in lib.pnae_last_error()
    # pipelined multi-step graph: at least two output sets, 2 <= workspaces <= sets, complete pointer lists
    h = C.c_void_p()
    one = (C.c_void_p * 1)(256)
    three = (C.c_void_p * 3)(256, 512, 768)
    hole = (C.c_void_p * 3)(256, None, 768)
    args = lambda nsets, nws, outs, ws: (1, 4, nsets, nws, 2, 64, three, 64, three, outs, outs, outs, outs, d, d, outs, outs, ws, 1 << 20, C.byref(h))
    assert lib.pnae_chamfer_graph_create_pipelined(*args(1, 1, one, one)) == _lib.PNAE_ERR_INVALID_ARG
    assert lib.pnae_chamfer_graph_create_pipelined(*args(2, 3, three, three)) == _lib.PNAE_ERR_INVALID_ARG      # more workspaces than sets
    assert lib.pnae_chamfer_graph_create_pipelined(*args(3, 3, hole, three)) == _lib.PNAE_ERR_INVALID_ARG
    assert b"output set" in lib.pnae_last_error()
    same = (C.c_void_p * 3)(256, 256, 256)
    assert lib.pnae_chamfer_graph_create_pipelined(*args(3, 3, three, same)) == _lib.PNAE_ERR_INVALID_ARG
    assert b"workspace" in lib.pnae_last_error()
    assert not h.value


def test_no_cpu_fallback():
    a = torch.zeros(2, 8, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        tf_nndistance.nn_distance(a, a)
    with pytest.raises(RuntimeError, match="no CPU path"):
        tf_approxmatch.approx_match(a, a)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.match_cost_dense_fwd(a, a, torch.zeros(2, 8, 8))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pointnet_autoencoder_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), fn
                assert "liboracle" not in src and "libref_" not in src, fn


def _simulate_row_slots(be, nrb, nch, warps):
    """The device's span / slot arithmetic of nn_fwd_kernel (csrc/nn_distance.cu), restated: every warp w owns
    units [w*U//W, (w+1)*U//W) (chunk fastest); leaving a row block it stores its partial row keys in slot
    min(w - owner(first unit of the block), its first chunk in the block).  Returns, per row block, the slots used."""
    U = be * nrb * nch
    owner = lambda t: ((t + 1) * warps - 1) // U
    used = {}
    for w in range(warps):
        u, uend = w * U // warps, (w + 1) * U // warps
        while u < uend:
            blk, ch = divmod(u, nch)
            seg = min(nch - ch, uend - u)
            first = blk * nch
            used.setdefault(blk, []).append(min(w - owner(first), ch))
            u += seg
    return used, owner


@pytest.mark.parametrize("b,n,m,sms", [(32, 2048, 2048, 148), (1, 1, 1, 148), (3, 777, 333, 148), (2, 16384, 1024, 148),
                                       (8, 4096, 4096, 148), (64, 64, 64, 148), (5, 300, 5000, 148), (1, 257, 33, 148),
                                       (7, 2048, 2048, 4), (33, 40, 50, 1), (2, 1, 700, 148), (128, 2048, 2048, 148)])
def test_chamfer_plan_slot_arithmetic(b, n, m, sms):
    """Host logic of the Chamfer forward without a GPU: the launcher's slot bound covers every slot the sweep's
    warps use, slots of one row block are distinct, and the slot-0 warp's pad range [used, nsl) is consistent."""
    lib = _lib.load()
    plan = (C.c_int * 9)()
    assert lib.pnae_nn_distance_plan(b, n, m, sms, plan) == 0
    nrb, nch, nslot, be, warps, nsl_full, nsl_last, rpb, cpc = list(plan)
    assert nrb == -(-n // rpb) and nch == -(-m // cpc) and warps == sms * 16 and 1 <= be <= min(b, 65535)
    for be_launch, nsl in {(be, nsl_full), (b % be or be, nsl_last)}:
        assert 1 <= nsl <= nslot <= nch
        used, owner = _simulate_row_slots(be_launch, nrb, nch, warps)
        assert sorted(used) == list(range(be_launch * nrb))                  # every row block is flushed by someone
        for blk, slots in used.items():
            first = blk * nch
            span_count = min(owner(first + nch - 1) - owner(first) + 1, nch)  # what the slot-0 warp computes as `used`
            assert sorted(slots) == list(range(len(slots))), (blk, slots)     # distinct, dense from 0
            assert len(slots) == span_count <= nsl, (blk, slots, span_count, nsl)


@pytest.mark.parametrize("b,n,m,sms", [(32, 2048, 2048, 148), (1, 1, 1, 148), (3, 777, 333, 148), (2, 4096, 1024, 148),
                                       (4, 2048, 2048, 148), (64, 64, 64, 148), (5, 300, 5000, 148), (1, 513, 129, 148),
                                       (7, 2048, 2048, 4), (33, 40, 50, 1), (8, 16384, 16384, 148)])
def test_approx_match_plan_slot_arithmetic(b, n, m, sms):
    """Host logic of approx_match without a GPU.  Every sweep splits its (element, own block, streamed chunk) tasks
    into contiguous spans, one per CTA of the cooperative grid; a CTA leaving an own block stores its partial sums
    in slot min(cta - owner(first task of the block), its first chunk in the block), and the NEXT sweep sums exactly
    slots_of() = min(owner(last) - owner(first) + 1, chunks) slots.  Both counts must agree and fit the workspace."""
    lib = _lib.load()
    plan = (C.c_int * 6)()
    assert lib.pnae_approx_match_plan(b, n, m, sms, plan) == 0
    grid, nslot, maxnm, own, ts, threads = list(plan)
    assert maxnm == max(n, m) and own == 2 * threads and own % ts == 0 and grid % sms == 0
    for nown, nstr in ((n, m), (m, n)):
        nob, nch = -(-nown // own), -(-nstr // ts)
        T = b * nob * nch
        owner = lambda t: ((t + 1) * grid - 1) // T
        used = {}
        for c in range(grid):
            t, tend = c * T // grid, (c + 1) * T // grid
            while t < tend:
                blk, ch = divmod(t, nch)
                used.setdefault(blk, []).append(min(c - owner(blk * nch), ch))
                t += min(nch - ch, tend - t)
        assert sorted(used) == list(range(b * nob))
        for blk, slots in used.items():
            first = blk * nch
            ns = min(owner(first + nch - 1) - owner(first) + 1, nch)
            assert sorted(slots) == list(range(len(slots))) and len(slots) == ns <= nslot, (blk, slots, ns, nslot)


def test_header_is_plain_c_and_links(tmp_path):
    """include/pnae.h must be usable from C (the boundary is a C ABI): compile a strict C99 translation unit that
    includes it, link it against libpnae.so and run a host-only call."""
    import shutil, subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    _lib.load()
    src = tmp_path / "use_pnae.c"
    src.write_text('#include "pnae.h"\n#include <stdio.h>\n'
                   'int main(void) { int plan[9]; int rc = pnae_nn_distance_plan(32, 2048, 2048, 148, plan);\n'
                   '  printf("%d %d %d %d\\n", rc, pnae_version(), plan[0], plan[1]); return rc; }\n')
    exe = tmp_path / "use_pnae"
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                        "-o", str(exe), "-L", libdir, "-lpnae", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.split() == ["0", "100", "8", "64"], (out.stdout, out.stderr)
