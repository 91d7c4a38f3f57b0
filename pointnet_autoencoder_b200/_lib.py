"""ctypes binding of libpnae.so (the C ABI declared in include/pnae.h).

There is no fallback: if the library is missing and cannot be built, importing
any op raises.  The library is kept in-tree (pointnet_autoencoder_b200/libpnae.so).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpnae.so")

NUM_LEVELS = 10

PNAE_OK = 0
PNAE_ERR_INVALID_ARG = -1
PNAE_ERR_WORKSPACE = -2
PNAE_ERR_CUDA = -3
PNAE_ERR_UNSUPPORTED = -4

_vp = C.c_void_p
_i = C.c_int
_sz = C.c_size_t

# name -> (restype, argtypes); one entry per function declared in include/pnae.h
SIGNATURES = {
    "pnae_version": (_i, []),
    "pnae_last_error": (C.c_char_p, []),
    "pnae_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "pnae_fp32_probe": (_i, [_i, _vp, _sz, C.POINTER(C.c_longlong), _vp]),
    "pnae_nn_distance_workspace_bytes": (_sz, [_i, _i, _i]),
    "pnae_nn_distance_plan": (_i, [_i, _i, _i, _i, _vp]),
    "pnae_nn_distance_fwd": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pnae_nn_distance_bwd": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pnae_chamfer_loss_grad": (_i, [_i, _i, _vp, _i, _vp, C.c_float, C.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pnae_nn_distance_fwd_grad": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pnae_chamfer_graph_create_pipelined": (_i, [_i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, C.POINTER(_vp)]),
    "pnae_chamfer_graph_create_fused_multi": (_i, [_i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, C.POINTER(_vp)]),
    "pnae_chamfer_graph_create": (_i, [_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, C.POINTER(_vp)]),
    "pnae_chamfer_graph_create_multi": (_i, [_i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, C.POINTER(_vp)]),
    "pnae_chamfer_host_pipeline_create": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _sz, _vp, _vp, _vp, _sz, C.POINTER(_vp)]),
    "pnae_chamfer_host_pipeline_submit": (_i, [_vp, _vp, _vp, C.POINTER(_i)]),
    "pnae_chamfer_host_pipeline_drain": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "pnae_chamfer_host_pipeline_destroy": (_i, [_vp]),
    "pnae_graph_launch": (_i, [_vp, _vp]),
    "pnae_graph_destroy": (_i, [_vp]),
    "pnae_approx_match_workspace_bytes": (_sz, [_i, _i, _i]),
    "pnae_approx_match_plan": (_i, [_i, _i, _i, _i, _vp]),
    "pnae_approx_match": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pnae_match_from_factors": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pnae_match_cost_fwd": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pnae_match_cost_bwd": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pnae_match_cost_factors": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pnae_xyz_moments": (_i, [C.c_longlong, _vp, _vp, _i, _vp]),
    "pnae_mlp_layer_xyz": (_i, [C.c_longlong, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "pnae_mlp_first": (_i, [C.c_longlong, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "pnae_mlp_layer": (_i, [C.c_longlong, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "pnae_conv5_finish": (_i, [_i, _i, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "pnae_bn_fold": (_i, [_i, _vp, C.c_double, _vp, _vp, C.c_float, C.c_float, _i, _vp, _vp, _vp, _vp, _vp]),
    "pnae_mlp_apply_bf16": (_i, [C.c_longlong, _i, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, _i, _vp, _i, _vp]),
    "pnae_encoder_conv_pool": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
}


class PnaeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libpnae error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Load (building first if the .so is absent and nvcc exists).  Raises if neither works."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    override = os.environ.get("PNAE_LIB_OVERRIDE")          # tuning only (tools/build_variant.sh): time another build of the same sources
    if override:
        lib = C.CDLL(override)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib
    if not os.path.exists(LIB_PATH):
        if shutil.which("nvcc") is None:
            raise ImportError(
                "pointnet_autoencoder_b200: %s is missing and nvcc is not available to build it; "
                "run `python -m pointnet_autoencoder_b200.build`. There is no CPU fallback." % LIB_PATH)
        _build.build()
    elif not _build.up_to_date():
        # the .so does not match the sources next to it (an edit without a rebuild): never run a stale binary silently
        if shutil.which("nvcc") is not None and os.environ.get("PNAE_NO_REBUILD") is None:
            _build.build()
        else:
            raise ImportError("pointnet_autoencoder_b200: %s is older than its sources (csrc/*.cu, include/pnae.h); "
                              "run `python -m pointnet_autoencoder_b200.build`" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = header / library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != PNAE_OK:
        raise PnaeError(rc, load().pnae_last_error().decode("utf-8", "replace"))
