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
