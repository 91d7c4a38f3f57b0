#!/bin/bash
for i in $(seq 1 14); do
  timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pipelined or wide_index" 2>&1 | grep -E "passed|failed|FAILED|assert " | head -6
done
