#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q -k "mlp_layer or chain or fused" 2>&1 | grep -v "^$" | tail -3
timeout 300 python tools/mlp_time.py 2>&1 | tail -5
PNAE_LIB_OVERRIDE=$PWD/tools/wip/variants/libpnae_mlptrace.so timeout 200 python tools/mlp_trace.py 2>&1 | tail -5
