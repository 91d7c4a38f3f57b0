#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -k "emd or sweep_sizes or approx or match or smoke or full_size" > gpurun_out/i_pytest_emd.log 2>&1; tail -8 gpurun_out/i_pytest_emd.log
echo "--- pow4"; timeout 300 python tools/graph_time.py 32 2048 2048 --gen=chair --emd | tail -3
echo "--- nopow4"; PNAE_LIB_OVERRIDE=tools/wip/variants/libpnae_nopow4.so timeout 300 python tools/graph_time.py 32 2048 2048 --gen=chair --emd | tail -3
echo "--- pow4 B=4"; timeout 300 python tools/graph_time.py 4 2048 2048 --gen=chair --emd | tail -3
timeout 900 python tools/emd_truth_table.py chair 32 2048 > gpurun_out/i_truth_chair.txt 2>&1; tail -2 gpurun_out/i_truth_chair.txt
