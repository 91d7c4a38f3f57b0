#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -3
timeout 300 python tools/mlp_time.py 2>&1 | tail -5
